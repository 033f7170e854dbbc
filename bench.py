#!/usr/bin/env python
"""bench.py -- BASELINE.json headline metric: PSWM scan bp*motif/s (both strands) at 1/2/4/8 B200,
and promoter posteriors/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE configs[4], "config5" of SURVEY.md 8d): a motif set of 500 JASPAR-scale
PSWMs (length ~U{8..30}, 20 sites each drawn from a Dirichlet(0.5)-column PWM, pseudocount 1,
seed 5) over ONE synthetic 1 Gbp sequence (i.i.d. uniform ACGT, counter-based generator keyed by
the global position block, so every rank count sees the same sequence).  The sequence is
SHARDED over the N ranks (strong scaling): rank r owns the windows that start in its slice and
carries the (m_max - 1)-base halo; every motif is capped at the shard's own windows.

One step = for every motif ONE pass of the fused scan over the rank's shard -- background
histogram + moments of all windows (genome.py:215-226) AND the thresholded both-strand site list at
the motif's Patser threshold (genome.py:433-445; candidates re-scored exactly, sorted on the
device) -- followed by ONE NCCL all-reduce of {n, sum, sumsq, histogram} of all motifs and the
recomputation of mean / std on every rank.

value      bp*motif/s = (sum over motifs of windows) / step time, shard packed and resident in
           HBM, site lists produced (compacted, sorted) in HBM; CUDA events, max over ranks
e2e        the same from HOST memory: pinned ASCII shard -> H2D -> pack -> scan -> every site
           record and every background record back in host memory (+ the all-reduce)
roofline   the scan kernel over the whole motif set: 2m lookup-adds per bp*motif against
           SMs * 128 lanes * f_max (the binding bound), HBM bytes nested
posteriors BASELINE configs[1] (LexA 22 bp, one 5 Mbp genome per rank, 3000 promoters): background
           fit + posteriors, promoter posteriors/s; its own e2e through cgbk_regulation_pipeline
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "PSWM scan bp*motif/s (both strands): background record + thresholded sites per motif, sequence-sharded"
UNIT = "bp*motif/s"
MOTIF_SEED = 5
GEN_BLOCK = 1 << 26          # bases per generator block (seeded by the global block number)

# config 2 (the `posteriors` sub-object)
GENOME_BP = 5_000_000
N_GENES = 3000
UP, DOWN = 300, 50


def workload_name(n_motifs, bp):
    return ("config5: %d synthetic JASPAR-scale motifs (8-30 bp, seed %d) x one %d bp synthetic sequence, "
            "sequence-sharded over the ranks with an (m_max - 1)-base halo; per motif background histogram + "
            "moments and the Patser-threshold site list in one fused pass; one all-reduce per step"
            % (n_motifs, MOTIF_SEED, bp))


# ----------------------------------------------------------------------------- inputs
def motif_site_lists(n, seed=MOTIF_SEED):
    """length ~U{8..30}, 20 sites from a Dirichlet(0.5)-column PWM."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        m = int(rng.integers(8, 31))
        cols = rng.dirichlet([0.5] * 4, size=m)
        out.append(["".join("ACGT"[rng.choice(4, p=cols[j])] for j in range(m)) for _ in range(20)])
    return out


def shard_bounds(bp, m_max, world, rank):
    """Rank's slice of window START positions [lo, lo + cap) and the bases it must hold."""
    own = (bp + world - 1) // world
    lo = min(bp, rank * own)
    cap = max(0, min(own, bp - lo))
    hi = min(bp, lo + cap + m_max - 1)
    return lo, cap, hi


def generate_bases(lo, hi, dev=None):
    """ASCII bases [lo, hi) of THE benchmark sequence (a function of the position only)."""
    import torch
    if dev is None:
        gen_dev = torch.device("cpu")
    else:
        gen_dev = dev
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=gen_dev)
    out = torch.empty(hi - lo, dtype=torch.uint8, device=gen_dev)
    for b0 in range(lo // GEN_BLOCK * GEN_BLOCK, hi, GEN_BLOCK):
        gen = torch.Generator(device=gen_dev)
        gen.manual_seed(1000 + b0 // GEN_BLOCK)
        vals = lut[torch.randint(0, 4, (GEN_BLOCK,), device=gen_dev, generator=gen)]
        a, b = max(b0, lo), min(b0 + GEN_BLOCK, hi)
        out[a - lo:b - lo] = vals[a - b0:b - b0]
    return out


def lexa():
    with open(os.path.join(ROOT, "tests", "golden", "lexa_sites.json")) as f:
        d = json.load(f)
    site_lists = [mo["sites"] for mo in d["motifs"]]
    counts = [len(s) for s in site_lists]
    weights = [c / float(sum(counts)) for c in counts]
    return d, site_lists, weights


def synthetic_genome(n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)].copy()


def gene_layout(n, n_genes):
    step = n // n_genes
    start = np.arange(n_genes, dtype=np.int64) * step + step // 3
    end = np.minimum(n, start + min(1000, step // 2))
    strand = np.where(np.arange(n_genes) % 2 == 0, 1, -1)
    return start, end, strand


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def git_head():
    try:
        return subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT,
                                       stderr=subprocess.DEVNULL).decode().strip()
    except Exception:
        return None


def ncu_traffic(kernel_key):
    """DRAM traffic per launch of the dominant kernel from the committed ncu capture
    (profiles/r2_ncu_traffic.json, written from an `ncu --set full` run; null when absent)."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        d = json.load(f)
    e = d.get(kernel_key)
    if not e:
        return None, None
    return e.get("dram_bytes_per_base"), {k: e.get(k) for k in ("captured_at_commit", "workload", "source", "note")}


def spread(ms):
    ms = np.asarray(ms, dtype=np.float64)
    return {"min": float(ms.min()), "median": float(np.median(ms)), "max": float(ms.max())}


# ----------------------------------------------------------------------------- CPU arms
def cpu_motif_job(args):
    """Reference CPU path for a list of motifs over one sample of the sequence: per motif the
    background scores of every window (score_seq both strands -> mean / std / histogram) and the
    thresholded site loop on both strands (genome.py:433-445), oracle/py_port.py."""
    site_lists, thresholds, seq = args
    from oracle import np_oracle as O
    from oracle.py_port import PortModel
    units = 0
    hits = 0
    for sites, thr in zip(site_lists, thresholds):
        om = O.OracleModel([sites], [1.0])
        port = PortModel(om)
        bg = port.score_seq(seq)                                 # soft-max of every window
        np.mean(bg), np.std(bg)
        np.histogram(bg, bins=256)
        hits += len(port.region_hits(seq, 0, thr))
        units += len(bg)
    return units, hits


def cpu_sample_thresholds(site_lists):
    """Patser thresholds through the product's host C code (float64, no GPU): the CPU arm times
    the scan, not the once-per-model threshold DP."""
    from cgb_b200 import PSSMModel, SiteCollection, prepare_thresholds
    models = [PSSMModel([SiteCollection(s)], [1.0]) for s in site_lists]
    prepare_thresholds(models)
    return [M.threshold() for M in models]


_REF_SEQ = None


def _ref_worker(args):
    site_lists, thresholds = args
    return cpu_motif_job((site_lists, thresholds, _REF_SEQ))


def run_reference(args):
    """--impl reference: the reference's CPU path (Python port with the C inner scan,
    oracle/py_port.py; the reference tree itself cannot travel to the GPU box and is pinned to the
    port in the container tests) on all host cores, on a bounded sample of the same workload per
    step: all motifs x the first `--ref-sample-bp` bases of the same sequence."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import c_oracle
    c_oracle.build()
    cores = os.cpu_count() or 1
    site_lists = motif_site_lists(args.motifs)
    thresholds = cpu_sample_thresholds(site_lists)
    global _REF_SEQ
    _REF_SEQ = generate_bases(0, args.ref_sample_bp).numpy().tobytes()
    jobs = [(site_lists[c::cores], thresholds[c::cores]) for c in range(cores) if site_lists[c::cores]]
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(len(jobs)) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_worker, jobs)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    units = sum(r[0] for r in res)
    total = float(np.sum(times))
    value = units * len(times) / total
    sample = ("per step all %d motifs x the first %d bases of the benchmark sequence (%d bp*motif), "
              "%d processes" % (args.motifs, args.ref_sample_bp, units, len(jobs)))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "step_ms": spread(np.array(times) * 1e3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args.motifs, args.bp), "motifs": args.motifs, "sequence_bp": args.bp,
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": len(jobs), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_config2(seq_bytes, promoters, bg_slice, lexa_d, site_lists, weights):
    """One config-2 step of the reference's CPU path (oracle/py_port.py) on one core."""
    from oracle import np_oracle as O
    from oracle.py_port import PortModel, config2_step
    om = O.OracleModel(site_lists, weights)
    port = PortModel(om)
    self_scores = [port.score_seq(s) for s in om.sites]            # build_bayesian_estimator
    port._mu_m, port._std_m = np.mean(self_scores), np.std(self_scores)
    t0 = time.perf_counter()
    nwin, nprom, mu, sd, posts = config2_step(port, seq_bytes, promoters, lexa_d["prior_regulation_probability"],
                                              lexa_d["alpha"], bg_slice)
    return time.perf_counter() - t0, nwin, nprom


# ----------------------------------------------------------------------------- host binding
def bind_to_gpu_numa(dev_index):
    """Pin this process (and therefore the pages its pinned staging buffers are first touched on)
    to the CPUs of the GPU's NUMA node.  Returns a description of what was pinned."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(dev_index).pci_bus_id
        dom = torch.cuda.get_device_properties(dev_index).pci_domain_id
        devn = torch.cuda.get_device_properties(dev_index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, devn)
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return "no NUMA information for the GPU (single node)"
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return "NUMA node %d has no CPU this process may use" % node
        os.sched_setaffinity(0, allowed)
        return "process + pinned staging buffers on NUMA node %d (%d CPUs)" % (node, len(allowed))
    except Exception as exc:
        return "not pinned (%s)" % str(exc)[:60]


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from cgb_b200 import GeneTable, PackedGenome, PSSMModel, SiteCollection, _cabi, prepare_thresholds
    from cgb_b200 import dist as cdist
    from cgb_b200 import motif_batch as mb
    from tools.clock_sampler import ClockSampler

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor(x, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu().numpy()

    def sum_over_ranks(x):
        t = torch.tensor(x, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    # ---- the motif set; thresholds by the host DP
    site_lists = motif_site_lists(args.motifs)
    all_models = [PSSMModel([SiteCollection(s)], [1.0], device=dev) for s in site_lists]
    t0 = time.perf_counter()
    prepare_thresholds(all_models)
    all_thr = [M.threshold() for M in all_models]
    m_len = np.array([M.length for M in all_models])
    m_max = int(m_len.max())
    K = len(all_models)
    bp = args.bp

    # ---- the grid: S sequence shards x G motif groups over the ranks (cgb_b200/dist.py).  The
    # sequence is cut while the shards stay >= --min-shard-bp, the remaining ranks split the motif
    # set by estimated cost; rank r scans shard r % S for group r // S
    S, G_m = cdist.plan_grid(world, bp, args.min_shard_bp)
    seq_rank, grp_rank = rank % S, rank // S
    costs = [mb.estimated_cost_ms(M, bp // S) for M in all_models]
    groups = cdist.plan_motif_shards(costs, G_m)
    mine = groups[grp_rank]
    mine_t = torch.tensor(mine, dtype=torch.int64, device=dev)
    models = [all_models[i] for i in mine]
    thr = [all_thr[i] for i in mine]
    batch = mb.ModelBatch(models, dev)          # this rank's tables in one device block
    setup_s = time.perf_counter() - t0
    group_cost = [float(sum(costs[i] for i in gset)) for gset in groups]

    # ---- this rank's shard of THE sequence, packed and resident
    lo, cap, hi = shard_bounds(bp, m_max, S, seq_rank)
    ascii_dev = generate_bases(lo, hi, dev)
    g = PackedGenome([hi - lo], dev)
    g.pack_into(0, ascii_dev)
    torch.cuda.synchronize()
    pinned = None
    if not args.no_e2e:
        pinned = torch.empty(hi - lo, dtype=torch.uint8).pin_memory()
        pinned.copy_(ascii_dev)
    del ascii_dev
    torch.cuda.empty_cache()
    # windows this rank scores (zero for the motifs of other groups)
    owned = np.zeros(K, dtype=bool)
    owned[mine] = True
    win_rank = np.array([max(0, min(cap, bp - lo - m + 1)) for m in m_len], dtype=np.float64) * owned
    win_total = np.array([bp - m + 1 for m in m_len], dtype=np.float64)
    units_total = float(win_total.sum())
    exp_hits = np.array([mb.expected_hits(M, cap) * 1.25 + 4096 for M in models])
    hit_cap = int(min(exp_hits.sum() * 1.05, 1.5e9))

    kern_own = np.zeros(len(models))
    site_own = np.zeros(len(models))
    full_stats = torch.zeros((K, _cabi.BG_COUNT), dtype=torch.float64, device=dev)
    full_hist = torch.zeros((K, args.bins), dtype=torch.int64, device=dev) if args.bins else None

    def step():
        """One step: this rank's motifs over this rank's shard, then ONE all-reduce of the record
        table of ALL motifs (rows this rank did not scan are zero): afterwards every rank holds the
        complete background record of every motif."""
        res = mb.scan_batch(models, g, thr, n_bins=args.bins, win_cap=cap, to_host=False, recycle=True,
                            hit_cap=hit_cap, kernel_ms=kern_own if timing_kernels[0] else None,
                            site_ms=site_own if timing_kernels[0] else None)
        nbytes = 0
        if compute_done[0] is not None:
            compute_done[0].record()                 # this rank's own work of the step ends here
        if world > 1:
            full_stats.zero_()
            full_stats[mine_t] = res["stats"]
            if full_hist is not None:
                full_hist.zero_()
                full_hist[mine_t] = res["hist"]
            nbytes = cdist.allreduce_batch(full_stats, full_hist)
            out = {"stats": full_stats, "hist": full_hist}
        else:
            out = {"stats": res["stats"], "hist": res["hist"]}
        out["hits"], out["launches"] = res["hits"], res["launches"]
        return out, nbytes

    timing_kernels = [False]
    compute_done = [None]
    # ---- the in-run checks after one step, then the warm-up steps (all outside the timed region).
    # The order matters: the sharded-vs-one-GPU check drops rank 0's scratch buffers and allocator
    # cache; with the warm-up BEFORE it the first two timed steps of every rank paid for that
    # (761 and 625 ms instead of 553 at N = 2).
    res, ar_bytes = step()
    torch.cuda.synchronize()
    stats = res["stats"].cpu().numpy()
    hist = res["hist"].cpu().numpy() if res["hist"] is not None else None
    hits_rank = np.zeros(K)
    hits_rank[mine] = np.array(res["hits"], dtype=np.float64)
    hits_all = sum_over_ranks(hits_rank)
    checks = {"n_equals_bp_minus_m_plus_1": bool(np.array_equal(stats[:, _cabi.BG_N], win_total)),
              "hist_sums_to_n": bool(hist is None or np.array_equal(hist.sum(axis=1), stats[:, _cabi.BG_N])),
              "hits_total": float(hits_all.sum())}
    assert checks["n_equals_bp_minus_m_plus_1"], "window counts after the all-reduce are wrong"
    assert checks["hist_sums_to_n"], "histograms do not sum to the window counts"
    if world > 1 and not args.no_check:
        # rank 0 scans the first motifs over the WHOLE sequence on its own: the sharded, all-reduced
        # records and the summed site counts must match the one-GPU ones
        P = min(args.check_motifs, K)
        ok = 1.0
        if rank == 0:
            whole = PackedGenome([bp], dev)
            whole.pack_into(0, generate_bases(0, bp, dev))
            one = mb.scan_batch(all_models[:P], whole, all_thr[:P], n_bins=args.bins, to_host=False, recycle=True,
                                hit_cap=int(min(sum(mb.expected_hits(M, bp) for M in all_models[:P]) * 1.3 + 1e6, 1.5e9)))
            s1 = one["stats"].cpu().numpy()
            same_n = np.array_equal(s1[:, _cabi.BG_N], stats[:P, _cabi.BG_N])
            same_hist = hist is None or np.array_equal(one["hist"].cpu().numpy(), hist[:P])
            same_hits = np.array_equal(np.array(one["hits"], dtype=np.float64), hits_all[:P])
            dmean = float(np.max(np.abs(s1[:, _cabi.BG_MEAN] - stats[:P, _cabi.BG_MEAN])))
            dstd = float(np.max(np.abs(s1[:, _cabi.BG_STD] - stats[:P, _cabi.BG_STD])))
            checks["sharded_vs_one_gpu"] = {"motifs": P, "same_n": bool(same_n), "same_histogram": bool(same_hist),
                                            "same_hit_counts": bool(same_hits), "max_abs_dmean": dmean,
                                            "max_abs_dstd": dstd}
            ok = float(same_n and same_hist and same_hits and dmean < 1e-9 and dstd < 1e-9)
            del whole, one
            mb.release_buffers(dev)
            torch.cuda.empty_cache()
        ok = float(max_over_ranks([1.0 - ok])[0] == 0.0)
        assert ok == 1.0, "sharded run differs from the one-GPU run: %s" % checks.get("sharded_vs_one_gpu")

    for _ in range(args.warmup):
        res, ar_bytes = step()
    torch.cuda.synchronize()

    # ---- timed: K steps, device-resident
    sampler = ClockSampler(dev, period_s=0.005)
    if rank == 0:
        sampler.start()
    # every motif's scan kernel is bracketed by its own CUDA event pair on the launching stream
    # (two event records per ~ms kernel): the roofline's kernel time comes from the timed steps
    _cabi.profile_enable(2)
    timing_kernels[0] = True
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    barrier()
    t_wall0 = time.perf_counter()
    ev_own = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for i in range(args.steps):
        ev[i][0].record()
        compute_done[0] = ev_own[i]
        res, ar_bytes = step()
        ev[i][1].record()
        launches += res["launches"] + (1 if world > 1 else 0)
    compute_done[0] = None
    barrier()
    wall = time.perf_counter() - t_wall0
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    # per rank: its own work of each step (start of the step -> before the all-reduce); the rest of a
    # step is the collective and the wait for the slowest rank
    own_ms = np.array([a.elapsed_time(b) for (a, _), b in zip(ev, ev_own)])
    own_by_rank = np.zeros((world, args.steps))
    own_by_rank[rank] = own_ms
    own_by_rank = sum_over_ranks(own_by_rank.ravel().tolist()).reshape(world, args.steps)
    _cabi.profile_enable(False)
    timing_kernels[0] = False
    kern_ms = np.full(K, np.nan)
    site_ms = np.full(K, np.nan)
    kern_ms[mine] = kern_own / args.steps
    site_ms[mine] = site_own / args.steps

    # all-reduce alone (the collective of the step, same buffer)
    ar_us = None
    if world > 1:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        cdist.allreduce_batch(res["stats"].clone(), res["hist"].clone() if res["hist"] is not None else None)
        barrier()
        st2, h2 = res["stats"].clone(), (res["hist"].clone() if res["hist"] is not None else None)
        a.record()
        for _ in range(reps):
            cdist.allreduce_batch(st2, h2)
        b.record()
        torch.cuda.synchronize()
        ar_us = float(max_over_ranks([a.elapsed_time(b) / reps * 1e3])[0])

    # ---- end to end from host memory
    e2e = None
    if not args.no_e2e:
        sink_count = [0]

        def sink(k, rec):
            sink_count[0] += len(rec)

        def e2e_step():
            gg = PackedGenome.from_sequences(None, dev, pinned=[pinned])        # H2D + pack
            r = mb.scan_batch(models, gg, thr, n_bins=args.bins, win_cap=cap, to_host=True, sink=sink,
                              split=not args.e2e_records8)
            if world > 1:
                own = np.concatenate([r["stats"][:, :3], r["hist"].astype(np.float64)], axis=1) \
                    if r["hist"] is not None else r["stats"][:, :3].copy()
                rec = np.zeros((K, own.shape[1]))
                rec[mine] = own
                t = torch.from_numpy(rec).to(dev)
                dist.all_reduce(t)
                rec = t.cpu().numpy()
            return r

        e2e_steps = min(args.steps, args.e2e_steps)
        for _ in range(min(2, args.warmup)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        sink_count[0] = 0
        for _ in range(e2e_steps):
            r = e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_ms = float(max_over_ranks([e2e_s * 1e3 / e2e_steps])[0])
        site_bytes = 8 if args.e2e_records8 else 6
        block_bytes = 0 if args.e2e_records8 else 4 * len(models) * (int(_cabi.lib().cgbk_split_sites_blocks(g.cap_bases)) + 1)
        d2h = sink_count[0] / e2e_steps * site_bytes + block_bytes + \
            len(models) * (8 * _cabi.BG_COUNT + 8 * args.bins + 8 * _cabi.CTR_COUNT)
        d2h_all = float(sum_over_ranks([d2h])[0])
        h2d_all = float(sum_over_ranks([float(hi - lo)])[0])
        e2e = {"value": units_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_all),
               "d2h_bytes_per_step": int(d2h_all), "ms_per_step": e2e_ms, "steps": e2e_steps,
               "call": "PackedGenome.from_sequences(pinned ASCII) + motif_batch.scan_batch(to_host=True%s) "
                       "-> cgbk_pack_ascii / cgbk_scan_batch%s" % (("", "") if args.e2e_records8 else
                                                                   (", split=True", " / cgbk_split_sites")),
               "d2h": ("every site record (8 B, sorted per motif)" if args.e2e_records8 else
                       "every site (6 B: uint16 offset in its 4 096-base block + strand, float32 score; plus one int32 "
                       "per block and motif -- the 8-byte records are rebuilt exactly on the host, "
                       "motif_batch.SplitSites)") +
                      " and every background record lands in pinned host memory; the sites of one sub-batch cross "
                      "PCIe while the next sub-batch is scanned",
               "sites_per_step": int(sink_count[0] / e2e_steps),
               "host_binding": numa}
        e2e["d2h_gb_per_s_achieved"] = round(d2h_all / (e2e_ms * 1e-3) / 1e9, 1)
        # what the box can move device -> host with every rank copying at once (the end-to-end leg is
        # bounded by it: every site record crosses PCIe into ONE host's memory): 2 x 1 GiB per rank
        probe_d = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        probe_h = pinned[:1 << 30] if pinned.numel() >= (1 << 30) else torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
        probe_h.copy_(probe_d, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            probe_h.copy_(probe_d, non_blocking=True)
        torch.cuda.synchronize()
        rate = 2 * float(1 << 30) / (time.perf_counter() - t0) / 1e9
        e2e["d2h_gb_per_s_platform"] = {"aggregate_all_ranks_copying": round(float(sum_over_ranks([rate])[0]), 1),
                                        "slowest_rank": round(float(-max_over_ranks([-rate])[0]), 1),
                                        "how": "2 x 1 GiB device -> pinned host per rank, all ranks at once"}
        del probe_d, probe_h
        del pinned
    clocks = sampler.stop() if rank == 0 else None

    total_ms = float(max_over_ranks([float(step_ms.sum())])[0])
    step_ms_max = max_over_ranks(step_ms.tolist())
    kern_rank = float(np.nansum(kern_ms))
    la_rank = float(np.nansum(2.0 * m_len * win_rank * (~np.isnan(kern_ms))))

    # ---- config 2 (posteriors/s), every rank its own genome, no collective
    post = None
    if not args.no_posteriors:
        mb.release_buffers(dev)
        torch.cuda.empty_cache()
        post = posteriors_leg(args, dev, rank, world, barrier, max_over_ranks)

    if rank == 0:
        hbm_peak, sm_max, peak_src = peaks()
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        issue_peak = sms * 128 * sm_max * 1e6
        value = units_total * args.steps / (total_ms * 1e-3)
        by_g = {}
        for G in sorted(set(((m_len + 3) // 4).tolist())):
            sel = ((m_len + 3) // 4 == G) & ~np.isnan(kern_ms)
            if sel.any():
                t = float(kern_ms[sel].sum()) * 1e-3
                by_g["G%d" % G] = {"motifs": int(sel.sum()), "kernel_ms_mean": float(kern_ms[sel].mean()),
                                   "bp_motif_per_s": float(win_rank[sel].sum() / t),
                                   "issue_frac": float((2.0 * m_len[sel] * win_rank[sel]).sum() / t / issue_peak)}
        # DRAM bytes of one scan launch: measured per base under ncu (1 Gbp launches), times this
        # rank's shard (the kernel streams the shard once, whatever its size)
        per_base, traffic_src = ncu_traffic("fused_scan_1gbp")
        traffic = per_base * float(hi - lo) if per_base is not None else None
        roofline = {
            "bound": "issue", "kernel": "fast_scan_kernel<G, hist+moments+candidates> (fused scan), all %d motifs" % K,
            "achieved": la_rank / (kern_rank * 1e-3), "peak": issue_peak, "unit": "lookup-add/s",
            "frac": la_rank / (kern_rank * 1e-3) / issue_peak,
            "definition": "algorithmic 2*m lookup-adds per bp*motif (SURVEY 8d) x windows of this rank's shard, over the "
                          "CUDA-event time of the scan kernels alone (one event pair per motif, on the launching stream); "
                          "peak = SMs x 128 lanes x f_max; north_star's scan roofline is the slower of this and HBM, "
                          "this one binds",
            "kernel_ms_per_step": kern_rank, "kernel_share_of_step": kern_rank / (float(step_ms.mean())),
            "site_pass_ms_per_step": float(np.nansum(site_ms)),
            "site_pass": "per motif: prefix sums of the scan's per-tile candidate counts, exact float64 re-score of "
                         "the flagged windows, ordered compaction into the site list (CUDA events around the pass)",
            "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": 0.25 * float(hi - lo),
            "hbm": {"achieved": float(np.nansum(0.25 * win_rank * (~np.isnan(kern_ms)))) / (kern_rank * 1e-3) / 1e9,
                    "peak": hbm_peak, "unit": "GB/s",
                    "frac": float(np.nansum(0.25 * win_rank * (~np.isnan(kern_ms)))) / (kern_rank * 1e-3) / 1e9 / hbm_peak,
                    "algorithmic_bytes_per_step": float(0.25 * win_rank.sum())},
            "by_group_count": by_g, "peak_source": peak_src,
        }
        cpu_baseline = None
        if world == 1 and not args.no_cpu:
            from oracle import c_oracle
            c_oracle.build()
            n_s = args.cpu_sample_bp
            seq_s = generate_bases(0, n_s).numpy().tobytes()
            t0 = time.perf_counter()
            units_cpu, hits_cpu = cpu_motif_job((site_lists, thr, seq_s))
            t_cpu = time.perf_counter() - t0
            cpu_baseline = {"value": units_cpu / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": "all %d motifs x the first %d bases of the benchmark sequence once "
                                      "(%d bp*motif, %.1f s on one core)" % (K, n_s, units_cpu, t_cpu)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "step_ms": spread(step_ms_max),
            "own_work_ms_by_rank": [round(float(x), 2) for x in own_by_rank.mean(axis=1)],
            "step_ms_each": [round(float(x), 2) for x in step_ms_max],
            "own_work_ms_each": [[round(float(x), 1) for x in row] for row in own_by_rank],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32 fixed point (2^-21..2^-28 per motif) scan, f64 exact re-score of site candidates, f64 moments",
            "data": "synthetic",
            "config": {"workload": workload_name(K, bp), "motifs": K, "sequence_bp": bp,
                       "motif_lengths": "8-30 (mean %.1f)" % float(m_len.mean()), "hist_bins": args.bins,
                       "windows_x_motifs_per_step": units_total, "shard_bp_per_rank": int(hi - lo),
                       "halo_bases": m_max - 1,
                       "grid": {"sequence_shards": S, "motif_groups": G_m, "min_shard_bp": args.min_shard_bp,
                                "motifs_per_group": [len(x) for x in groups],
                                "estimated_ms_per_group": [round(c, 1) for c in group_cost],
                                "rule": "the sequence is cut while shards stay >= min_shard_bp; the ranks left over "
                                        "split the motif set by estimated cost (longest first); rank r scans shard "
                                        "r % S for group r // S"},
                       "collective": "one all-reduce(SUM) of float64 [%d][3 + %d] per step (%d bytes): the record table "
                                     "of all motifs, rows a rank did not scan are zero"
                                     % (K, args.bins, int(ar_bytes)) if world > 1 else "none (one rank)",
                       "allreduce_us": ar_us,
                       "l2": "inputs larger than L2 (packed shard %.0f MB per pass, %d passes per step)"
                             % ((hi - lo) / 4e6, K),
                       "timing": "per-step CUDA events on the launching stream, barrier + synchronize on both sides, "
                                 "max over ranks",
                       "hit_buffer_records": hit_cap,
                       "sites_per_step": checks["hits_total"], "model_setup_s": setup_s},
            "checks": checks, "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": int(launches), "posteriors": post,
            "clocks": clocks, "wall_s_timed_loop": wall, "commit": git_head(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    del batch


def posteriors_leg(args, dev, rank, world, barrier, max_over_ranks):
    """BASELINE configs[1]: LexA 22-bp PSWM, one synthetic 5 Mbp genome per rank, 3000 promoters x
    350 bp: genome-wide background histogram + moments, then the posterior of every promoter."""
    import torch

    from cgb_b200 import GeneTable, PackedGenome, PSSMModel, SiteCollection, _cabi

    d, site_lists, weights = lexa()
    prior, alpha = d["prior_regulation_probability"], d["alpha"]
    M = PSSMModel([SiteCollection(s) for s in site_lists], weights, device=dev)
    host = synthetic_genome(GENOME_BP, 2 + rank)
    pinned = torch.from_numpy(host).pin_memory()
    gs, ge, gstrand = gene_layout(GENOME_BP, N_GENES)
    gt = GeneTable(gs, ge, gstrand, GENOME_BP)
    ps, pe = gt.promoter_region_location(UP, DOWN)
    g = PackedGenome.from_sequences(None, dev, pinned=[pinned])
    desc = M.promoter_descriptors(g, ps, pe)
    n_windows = g.total_windows(M.length)
    bg = M.background_stats(g, n_bins=256, keep_scores=True)
    M.fit_background(bg=bg, sync=True)
    post = torch.empty(N_GENES, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        M.background_stats(g, out=bg, keep_scores=True)
        M.fit_background(bg=bg, sync=False)
        M.binding_probabilities(g, None, None, prior, alpha, out=post, region_tensors=desc)

    # per-kernel time of the stats scan (eager launches, CUDA events inside the library)
    _cabi.profile_enable(1)
    kern = []
    for i in range(8):
        step()
        torch.cuda.synchronize()
        if i >= 3:
            kern.append(_cabi.profile_last_ms())
        flush.zero_()
    _cabi.profile_enable(False)
    # the step replayed from a CUDA graph (one memset + two kernels): `inner` replays per timed
    # step, L2 flushed before each timed step
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        step()
    cold, warm = [], []
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(3 + args.post_steps):
        flush.zero_()
        a.record(); graph.replay(); b.record()
        torch.cuda.synchronize()
        if i >= 3:
            cold.append(a.elapsed_time(b))
    inner = 200
    for i in range(2 + 5):
        a.record()
        for _ in range(inner):
            graph.replay()
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            warm.append(a.elapsed_time(b) / inner)
    barrier()
    # end to end through the host-buffer call
    for _ in range(5):
        M.regulation_pipeline(pinned, ps, pe, prior, alpha, n_bins=256)
    torch.cuda.synchronize()
    e2e_ms = []
    for _ in range(args.post_steps):
        t0 = time.perf_counter()
        M.regulation_pipeline(pinned, ps, pe, prior, alpha, n_bins=256)
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
    barrier()
    # the same from the host-PACKED form of the genome (cgbk_pack_host once, outside the loop: the
    # form a replicon is kept / stored in): bases2 plane only for this upper-case ACGT genome
    from cgb_b200 import HostPackedGenome
    t0 = time.perf_counter()
    hp = HostPackedGenome.from_ascii(pinned)
    pack_host_ms = (time.perf_counter() - t0) * 1e3
    for _ in range(5):
        pp, _, _ = M.regulation_pipeline(hp, ps, pe, prior, alpha, n_bins=256)
    torch.cuda.synchronize()
    e2e_packed_ms = []
    for _ in range(args.post_steps):
        t0 = time.perf_counter()
        pp, sp_, hp_ = M.regulation_pipeline(hp, ps, pe, prior, alpha, n_bins=256)
        e2e_packed_ms.append((time.perf_counter() - t0) * 1e3)
    pa, sa_, ha_ = M.regulation_pipeline(pinned, ps, pe, prior, alpha, n_bins=256)
    assert np.array_equal(pp, pa) and np.array_equal(hp_, ha_), "packed and ASCII pipelines disagree"
    barrier()
    e2e_packed_med = float(max_over_ranks([float(np.median(e2e_packed_ms))])[0])
    cold_med = float(max_over_ranks([float(np.median(cold))])[0])
    warm_med = float(max_over_ranks([float(np.median(warm))])[0])
    e2e_med = float(max_over_ranks([float(np.median(e2e_ms))])[0])
    out = None
    if rank == 0:
        hbm_peak, sm_max, _ = peaks()
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        k_ms = float(np.mean(kern))
        out = {
            "workload": "config2: LexA 22-bp PSWM, one synthetic 5 Mbp genome per rank, 3000 promoters x 350 bp "
                        "(genome-wide background histogram + moments, then posteriors); no collective",
            "promoter_posteriors_per_s": world * N_GENES / (cold_med * 1e-3),
            "bp_motif_per_s": world * n_windows / (cold_med * 1e-3),
            "step_ms_l2_flushed": spread(cold), "step_ms_l2_warm_200_replays": spread(warm),
            "timing": "CUDA-graph replay of the step (1 memset + 2 kernels), CUDA events; max over ranks of the median",
            "stats_kernel_us": k_ms * 1e3,
            "stats_kernel_issue_frac": n_windows * 2 * M.length / (k_ms * 1e-3) / (sms * 128 * sm_max * 1e6),
            "e2e": {"ms_per_step": e2e_med, "step_ms": spread(e2e_ms),
                    "promoter_posteriors_per_s": world * N_GENES / (e2e_med * 1e-3),
                    "bp_motif_per_s": world * n_windows / (e2e_med * 1e-3),
                    "h2d_bytes_per_step": int(pinned.numel()) + 12 * N_GENES + 16,
                    "d2h_bytes_per_step": N_GENES * 8 + 64 + 256 * 8,
                    "call": "PSSMModel.regulation_pipeline -> cgbk_regulation_pipeline (host buffers)"},
            "e2e_packed": {"ms_per_step": e2e_packed_med, "step_ms": spread(e2e_packed_ms),
                           "promoter_posteriors_per_s": world * N_GENES / (e2e_packed_med * 1e-3),
                           "bp_motif_per_s": world * n_windows / (e2e_packed_med * 1e-3),
                           "h2d_bytes_per_step": int(hp.upload_bytes) + 12 * N_GENES + 16,
                           "d2h_bytes_per_step": N_GENES * 8 + 64 + 256 * 8,
                           "host_pack_ms_once": pack_host_ms,
                           "call": "PSSMModel.regulation_pipeline(HostPackedGenome) -> cgbk_regulation_pipeline_packed "
                                   "(genome kept packed on the host: 0.25 B/base uploaded, no pack kernel); "
                                   "results identical to the ASCII call (asserted)"},
        }
        if world == 1 and not args.no_cpu:
            from oracle import c_oracle
            c_oracle.build()
            s2, e2 = np.clip(ps, 0, GENOME_BP), np.clip(pe, 0, GENOME_BP)
            # a tenth of the step (500 kbp of background windows + 300 promoters) on one core
            proms = [(int(x), int(y)) for x, y in zip(s2[:N_GENES // 10], e2[:N_GENES // 10])]
            t_cpu, nwin_cpu, nprom_cpu = cpu_config2(host.tobytes(), proms, (0, GENOME_BP // 10), d, site_lists, weights)
            out["cpu_baseline"] = {"promoter_posteriors_per_s": nprom_cpu / t_cpu, "bp_motif_per_s": nwin_cpu / t_cpu,
                                   "cores": 1, "kind": "port",
                                   "sample": "1/10 of the step (%d windows + %d promoters, %.1f s)"
                                             % (nwin_cpu, nprom_cpu, t_cpu)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--motifs", type=int, default=500)
    ap.add_argument("--bp", type=int, default=1_000_000_000)
    ap.add_argument("--bins", type=int, default=256)
    ap.add_argument("--e2e-steps", type=int, default=3, help="end-to-end steps (each moves every site record to the host)")
    ap.add_argument("--post-steps", type=int, default=30, help="timed steps of the posteriors leg")
    ap.add_argument("--check-motifs", type=int, default=8, help="motifs of the sharded-vs-one-GPU check (N > 1)")
    ap.add_argument("--min-shard-bp", type=int, default=500_000_000,
                    help="smallest sequence shard; ranks beyond bp / this split the motif set instead")
    ap.add_argument("--cpu-sample-bp", type=int, default=20_000)
    ap.add_argument("--ref-sample-bp", type=int, default=100_000)
    ap.add_argument("--e2e-records8", action="store_true",
                    help="end-to-end leg with 8-byte site records instead of the 6-byte split form")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg")
    ap.add_argument("--no-posteriors", action="store_true", help="skip the config-2 leg")
    ap.add_argument("--no-check", action="store_true", help="skip the sharded-vs-one-GPU check")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
