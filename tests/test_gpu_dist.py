"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): a sequence chunk-sharded over 2
ranks with the NCCL all-reduce of the background record must give the same background fit,
the same hit list and the same posteriors as one GPU and as the oracle."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import json

    import torch.distributed as dist

    from cgb_b200 import PackedGenome, PSSMModel, SiteCollection, _cabi
    from cgb_b200 import dist as cdist
    from oracle import np_oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        with open(os.path.join(root, "tests", "golden", "lexa_sites.json")) as f:
            d = json.load(f)
        sites = [mo["sites"] for mo in d["motifs"]]
        counts = [len(s) for s in sites]
        M = PSSMModel([SiteCollection(s) for s in sites], [c / float(sum(counts)) for c in counts], device=dev)
        seq = O.synthetic_genome(n, 77, n_frac=0.001)
        starts, ends, cnt = cdist.chunk_bounds(n, M.length, world)
        g = PackedGenome.from_sequences([seq[starts[rank]:ends[rank]]], dev)
        bg = M.background_stats(g, n_bins=512)
        cdist.allreduce_background(bg)              # the one collective of the path
        M.fit_background(bg=bg)
        _, pos, strand, score = M.scan_hits(g, M.threshold() - 3.0)
        gp, gs, gv = cdist.gather_hits(pos, strand, score, starts[rank])
        # the motif-batch form of the same (bench.py's step): a shard with the halo of the longest
        # motif, every motif capped at the shard's own windows, ONE all-reduce for all records
        from cgb_b200 import motif_batch as mb
        import numpy as np
        rng = np.random.default_rng(9)
        site_lists = []
        for m in (8, 13, 22, 30):
            cols = rng.dirichlet([0.5] * 4, size=m)
            site_lists.append(["".join("ACGT"[rng.choice(4, p=cols[j])] for j in range(m)) for _ in range(20)])
        Ms = [PSSMModel([SiteCollection(s2)], [1.0], device=dev) for s2 in site_lists]
        m_max = 30
        own = (n + world - 1) // world
        lo = rank * own
        cap = min(own, n - lo)
        gb = PackedGenome.from_sequences([seq[lo:min(n, lo + cap + m_max - 1)]], dev)
        thr = [Mk.threshold() for Mk in Ms]
        res = mb.scan_batch(Ms, gb, thr, n_bins=128, win_cap=cap, to_host=False)
        counts = torch.tensor([int(h.numel()) for h in res["hits"]], dtype=torch.int64, device=dev)
        cdist.allreduce_batch(res["stats"], res["hist"])
        dist.all_reduce(counts)
        if rank == 0:
            q.put((bg["stats"].cpu().numpy(), bg["hist"].cpu().numpy(), gp, gs, gv,
                   res["stats"].cpu().numpy(), res["hist"].cpu().numpy(), counts.cpu().numpy(), site_lists, thr))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_chunk_sharded_two_gpus():
    import torch.multiprocessing as mp

    from cgb_b200 import _cabi
    from oracle import c_oracle, np_oracle as O
    import json
    n, world = 2_000_003, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    stats, hist, gp, gs, gv, bstats, bhist, bcounts, site_lists, bthr = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "lexa_sites.json")) as f:
        d = json.load(f)
    sites = [mo["sites"] for mo in d["motifs"]]
    counts = [len(s) for s in sites]
    om = O.OracleModel(sites, [c / float(sum(counts)) for c in counts])
    seq = O.synthetic_genome(n, 77, n_frac=0.001)
    _, _, b = c_oracle.score_seq(om, seq)
    assert stats[_cabi.BG_N] == len(b) and hist.sum() == len(b)
    assert abs(stats[_cabi.BG_MEAN] - b.mean()) < 1e-6 and abs(stats[_cabi.BG_STD] - b.std()) < 1e-6
    p, s, v = c_oracle.scan_hits(om, seq, om.threshold() - 3.0)
    assert np.array_equal(gp, p) and np.array_equal(gs, s) and np.array_equal(gv, v)
    for k, sites in enumerate(site_lists):
        omk = O.OracleModel([sites], [1.0])
        _, _, bk = c_oracle.score_seq(omk, seq)
        assert bstats[k, _cabi.BG_N] == len(bk) == bhist[k].sum()
        assert abs(bstats[k, _cabi.BG_MEAN] - bk.mean()) < 1e-6 and abs(bstats[k, _cabi.BG_STD] - bk.std()) < 1e-6
        assert bcounts[k] == len(c_oracle.scan_hits(omk, seq, bthr[k])[0])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_devices_from_one_process(lexa, lexa_weights):
    """One host process driving two GPUs through the C ABI (the per-device kernel attributes,
    multiprocessor counts and pipeline streams are keyed by device; every entry point runs on its
    handle's device whatever device is current): same numbers on both, equal to the oracle."""
    from cgb_b200 import PackedGenome, PSSMModel, SiteCollection, _cabi
    from cgb_b200 import motif_batch as mb
    from oracle import c_oracle, np_oracle as O
    sites = [mo["sites"] for mo in lexa["motifs"]]
    w = lexa_weights["site_count"]
    om = O.OracleModel(sites, w)
    seq = O.synthetic_genome(300_000, 5, n_frac=0.001)
    thr = om.threshold() - 2.0
    ep, es, ev = c_oracle.scan_hits(om, seq, thr)
    _, _, b = c_oracle.score_seq(om, seq)
    starts = np.arange(200, 290000, 9000)
    results = []
    torch.cuda.set_device(0)                      # stays current: device 1 is reached through its handles
    for d in (1, 0, 1):
        dev = torch.device("cuda", d)
        M = PSSMModel([SiteCollection(s) for s in sites], w, device=dev)
        g = PackedGenome.from_sequences([seq], dev)
        _, pos, strand, score = M.scan_hits(g, thr)
        M.fit_background(g)
        post = M.binding_probabilities(g, starts, starts + 350, 0.03, 1 / 350.0).cpu().numpy()
        p2, st2, h2 = M.regulation_pipeline(seq, starts, starts + 350, 0.03, 1 / 350.0)
        res = mb.scan_batch([M], g, [thr], n_bins=64)
        assert torch.cuda.current_device() == 0
        assert np.array_equal(pos, ep) and np.array_equal(strand, es) and np.array_equal(score, ev)
        assert abs(M._mu_bg - b.mean()) < 1e-6 and abs(M._std_bg - b.std()) < 1e-6
        assert int(h2.sum()) == len(b) and np.max(np.abs(p2 - post)) < 1e-9
        gp, gs, gv = mb.unpack_hits(res["hits"][0])
        assert len(gp) == len(ep)
        results.append((post, M._mu_bg, M._std_bg))
    for post, mu, sd in results[1:]:
        assert np.array_equal(post, results[0][0]) and mu == results[0][1] and sd == results[0][2]
