"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: chunk sharding with halo,
the single all-reduce of the background record, hit gathering, genome assignment.
The per-rank "scan" is played by the CPU oracle so the data path needs no GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cgb_b200 import _cabi
from cgb_b200 import dist as cdist
from oracle import c_oracle, np_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _lexa_model():
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "lexa_sites.json")) as f:
        d = json.load(f)
    sites = [mo["sites"] for mo in d["motifs"]]
    counts = [len(s) for s in sites]
    return O.OracleModel(sites, [c / float(sum(counts)) for c in counts])


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        om = _lexa_model()
        m = om.length
        seq = O.synthetic_genome(n, 9)
        starts, ends, counts = cdist.chunk_bounds(n, m, world)
        chunk = seq[starts[rank]:ends[rank]]
        # per-rank background record from the oracle's scores of this chunk
        _, _, b = c_oracle.score_seq(om, chunk)
        assert len(b) == counts[rank]
        lo, hi, nb = -44.0, 28.0, 72
        hist, _ = np.histogram(b, bins=np.linspace(lo, hi, nb + 1))
        stats = torch.zeros(_cabi.BG_COUNT, dtype=torch.float64)
        stats[0], stats[1], stats[2] = len(b), b.sum(), (b * b).sum()
        bg = {"stats": stats, "hist": torch.from_numpy(hist.astype(np.int64))}
        cdist.allreduce_background(bg)
        # per-rank hits, gathered in global coordinates
        thr = om.threshold() - 4.0
        p, s, v = c_oracle.scan_hits(om, chunk, thr)
        gp, gs, gv = cdist.gather_hits(p, s, v, starts[rank])
        if rank == 0:
            q.put((bg["stats"].numpy().copy(), bg["hist"].numpy().copy(), gp, gs, gv))
    finally:
        dist.destroy_process_group()


def test_chunk_sharded_background_and_hits_gloo():
    n, world = 60011, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    stats, hist, gp, gs, gv = res
    om = _lexa_model()
    seq = O.synthetic_genome(n, 9)
    _, _, b = c_oracle.score_seq(om, seq)
    assert stats[_cabi.BG_N] == len(b) == n - om.length + 1
    assert abs(stats[_cabi.BG_MEAN] - b.mean()) < 1e-10 and abs(stats[_cabi.BG_STD] - b.std()) < 1e-9
    exp_hist, _ = np.histogram(b, bins=np.linspace(-44.0, 28.0, 73))
    assert np.array_equal(hist, exp_hist)
    p, s, v = c_oracle.scan_hits(om, seq, om.threshold() - 4.0)
    assert len(p) > 10
    assert np.array_equal(gp, p) and np.array_equal(gs, s) and np.array_equal(gv, v)


def test_chunk_bounds_cover_every_window_once():
    for n, m, w in ((1000, 22, 8), (30, 22, 8), (21, 22, 4), (10 ** 10, 22, 8), (997, 8, 3)):
        starts, ends, counts = cdist.chunk_bounds(n, m, w)
        nwin = max(0, n - m + 1)
        assert counts.sum() == nwin
        assert (ends <= n).all() and (starts >= 0).all()
        cover = 0
        for r in range(w):
            if counts[r]:
                assert starts[r] == cover and ends[r] - starts[r] == counts[r] + m - 1
                cover += counts[r]
        assert cover == nwin


def test_assign_genomes_balanced_and_complete():
    rng = np.random.default_rng(3)
    sizes = rng.integers(2_000_000, 7_000_000, size=100)
    parts = cdist.assign_genomes(sizes, 8)
    assert sorted(i for p in parts for i in p) == list(range(100))
    loads = [sum(int(sizes[i]) for i in p) for p in parts]
    assert max(loads) - min(loads) <= sizes.max()
    assert cdist.assign_genomes([5, 1], 4) == [[0], [1], [], []]


def test_finalize_stats_host():
    x = np.random.default_rng(1).normal(-12, 6, size=1000)
    st = torch.zeros(_cabi.BG_COUNT, dtype=torch.float64)
    st[0], st[1], st[2] = len(x), x.sum(), (x * x).sum()
    cdist.finalize_stats_host(st)
    assert abs(float(st[_cabi.BG_MEAN]) - x.mean()) < 1e-12 and abs(float(st[_cabi.BG_STD]) - x.std()) < 1e-10


def _batch_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        K, nb = 5, 16
        x = [rng.normal(-10 - k, 3 + k, size=200 + 10 * rank) for k in range(K)]
        stats = torch.zeros(K, _cabi.BG_COUNT, dtype=torch.float64)
        hist = torch.zeros(K, nb, dtype=torch.int64)
        for k in range(K):
            stats[k, 0], stats[k, 1], stats[k, 2] = len(x[k]), x[k].sum(), (x[k] * x[k]).sum()
            hist[k] = torch.from_numpy(np.histogram(x[k], bins=np.linspace(-40, 20, nb + 1))[0])
        nbytes = cdist.allreduce_batch(stats, hist)          # ONE collective for the K records
        if rank == 0:
            q.put((stats.numpy().copy(), hist.numpy().copy(), nbytes))
    finally:
        dist.destroy_process_group()


def test_allreduce_batch_gloo():
    """The motif-batch collective (bench.py's step): K background records summed over 2 ranks in
    one all-reduce, mean / std recomputed per record."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_batch_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    stats, hist, nbytes = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    K, nb = 5, 16
    assert nbytes == K * (3 + nb) * 8
    for k in range(K):
        parts = []
        for r in range(world):
            rng = np.random.default_rng(100 + r)
            xs = [rng.normal(-10 - j, 3 + j, size=200 + 10 * r) for j in range(K)]
            parts.append(xs[k])
        x = np.concatenate(parts)
        assert stats[k, _cabi.BG_N] == len(x) and hist[k].sum() == np.histogram(x, bins=np.linspace(-40, 20, nb + 1))[0].sum()
        assert abs(stats[k, _cabi.BG_MEAN] - x.mean()) < 1e-10 and abs(stats[k, _cabi.BG_STD] - x.std()) < 1e-9


def test_bench_shards_cover_every_window_once():
    """bench.py's sharding: rank r owns the windows that start in its slice; with the cap every
    motif length counts each window of the sequence exactly once."""
    import bench
    for bp, world in ((1000, 3), (10 ** 9, 8), (997, 4), (40, 8)):
        m_max = 30
        for m in (8, 19, 30):
            total = 0
            prev_end = 0
            for r in range(world):
                lo, cap, hi = bench.shard_bounds(bp, m_max, world, r)
                assert lo == prev_end and hi <= bp and hi - lo <= cap + m_max - 1
                prev_end = lo + cap
                total += max(0, min(cap, bp - lo - m + 1))
            assert prev_end == bp and total == max(0, bp - m + 1)
    # the benchmark sequence is a function of the position only: any slice of it is the same bases
    a = bench.generate_bases(0, 3000).numpy()
    b = bench.generate_bases(1234, 2000).numpy()
    assert np.array_equal(a[1234:2000], b) and set(np.unique(a)) <= set(b"ACGT")
    assert bench.motif_site_lists(5) == bench.motif_site_lists(5)
    assert all(8 <= len(s[0]) <= 30 and len(s) == 20 for s in bench.motif_site_lists(40))


def test_plan_grid_and_motif_groups():
    """bench.py's grid for a motif set x one long sequence: S sequence shards x M motif groups with
    S * M = world; every (window, motif) pair belongs to exactly one rank; the groups are balanced
    by estimated cost and deterministic."""
    import bench
    assert cdist.plan_grid(1, 10 ** 9) == (1, 1)
    assert cdist.plan_grid(2, 10 ** 9) == (2, 1)
    assert cdist.plan_grid(4, 10 ** 9) == (2, 2)
    assert cdist.plan_grid(8, 10 ** 9) == (2, 4)
    assert cdist.plan_grid(8, 10 ** 10) == (8, 1)
    assert cdist.plan_grid(8, 10 ** 9, min_shard_bases=1) == (8, 1)
    assert cdist.plan_grid(6, 10 ** 9, min_shard_bases=4 * 10 ** 8) == (2, 3)
    assert cdist.plan_grid(8, 1000) == (1, 8)
    rng = np.random.default_rng(3)
    costs = rng.uniform(0.5, 4.0, size=137)
    for n_groups in (1, 2, 3, 4, 8):
        groups = cdist.plan_motif_shards(costs, n_groups)
        assert groups == cdist.plan_motif_shards(list(costs), n_groups)
        assert sorted(i for g in groups for i in g) == list(range(137))
        assert all(g == sorted(g) for g in groups)
        load = [costs[g].sum() for g in groups]
        assert max(load) - min(load) <= costs.max()          # LPT: within one job of each other
    # coverage of the (window, motif) grid: rank r scans shard r % S for group r // S
    bp, world, m_max = 10 ** 6, 8, 30
    m_len = rng.integers(8, 31, size=40)
    S, M = cdist.plan_grid(world, bp, min_shard_bases=250_000)
    assert (S, M) == (4, 2)
    groups = cdist.plan_motif_shards([float(m) for m in m_len], M)
    seen = np.zeros(len(m_len))
    for r in range(world):
        lo, cap, hi = bench.shard_bounds(bp, m_max, S, r % S)
        for k in groups[r // S]:
            seen[k] += max(0, min(cap, bp - lo - int(m_len[k]) + 1))
    assert np.array_equal(seen, bp - m_len + 1)


def _grid_models(K=5):
    rng = np.random.default_rng(77)
    lens = [8, 30, 13, 22, 17][:K]
    return [O.OracleModel([["".join("ACGT"[i] for i in rng.integers(0, 4, m)) for _ in range(10)]], [1.0]) for m in lens]


def _grid_worker(rank, world, port, n, q):
    """One rank of bench.py's step on the CPU: its sequence shard (window starts capped, halo kept) for
    its motif group, rows of the other motifs zero, ONE all-reduce of the whole record table."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        oms = _grid_models()
        K, nb, m_max = len(oms), 24, max(om.length for om in oms)
        seq = O.synthetic_genome(n, 12)
        S, M = cdist.plan_grid(world, n, min_shard_bases=n // 2)
        groups = cdist.plan_motif_shards([float(om.length) for om in oms], M)
        lo, cap, hi = bench.shard_bounds(n, m_max, S, rank % S)
        shard = seq[lo:hi]
        stats = torch.zeros(K, _cabi.BG_COUNT, dtype=torch.float64)
        hist = torch.zeros(K, nb, dtype=torch.int64)
        hits = {}
        for k in groups[rank // S]:
            om = oms[k]
            _, _, b = c_oracle.score_seq(om, shard)
            b = b[:max(0, min(cap, n - lo - om.length + 1))]          # the windows that START in this shard
            stats[k, 0], stats[k, 1], stats[k, 2] = len(b), b.sum(), (b * b).sum()
            hist[k] = torch.from_numpy(np.histogram(b, bins=np.linspace(-60, 40, nb + 1))[0])
            p, s, v = c_oracle.scan_hits(om, shard, 2.0)
            keep = p < cap
            hits[k] = (p[keep] + lo, s[keep], v[keep])
        cdist.allreduce_batch(stats, hist)
        q.put((rank, (S, M), groups, stats.numpy().copy(), hist.numpy().copy(), hits))
    finally:
        dist.destroy_process_group()


def test_motif_group_by_sequence_shard_grid_gloo():
    """World size 4 = 2 sequence shards x 2 motif groups (the N = 4 / N = 8 layout of bench.py) on the
    CPU: after the ONE all-reduce EVERY rank holds the complete record of EVERY motif, equal to the
    one-process scan; the ranks' site lists, put together, are the one-process site lists."""
    n, world = 30011, 4
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_grid_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get() for _ in range(world)]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    oms = _grid_models()
    seq = O.synthetic_genome(n, 12)
    assert all(g[1] == (2, 2) for g in got)
    groups = got[0][2]
    assert sorted(k for grp in groups for k in grp) == list(range(len(oms)))
    for k, om in enumerate(oms):
        _, _, b = c_oracle.score_seq(om, seq)
        eh = np.histogram(b, bins=np.linspace(-60, 40, 25))[0]
        for rank, _, _, stats, hist, _ in got:                       # every rank, every motif
            assert stats[k, _cabi.BG_N] == len(b) == n - om.length + 1
            assert abs(stats[k, _cabi.BG_MEAN] - b.mean()) < 1e-10 and abs(stats[k, _cabi.BG_STD] - b.std()) < 1e-9
            assert np.array_equal(hist[k], eh)
        parts = sorted((g[5][k] for g in got if k in g[5]), key=lambda t: (t[0][0] if len(t[0]) else 0))
        assert len(parts) == 2                                       # one per sequence shard
        p, s, v = c_oracle.scan_hits(om, seq, 2.0)
        assert np.array_equal(np.concatenate([x[0] for x in parts]), p)
        assert np.array_equal(np.concatenate([x[1] for x in parts]), s)
        assert np.array_equal(np.concatenate([x[2] for x in parts]), v)
