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
        if rank == 0:
            q.put((bg["stats"].cpu().numpy(), bg["hist"].cpu().numpy(), gp, gs, gv))
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
    stats, hist, gp, gs, gv = q.get()
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
