#!/usr/bin/env python
"""bench.py — BASELINE.json headline metric on BASELINE configs[1].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload ("config2"): LexA 22-bp PSWM (site-count-weighted mix of the six
test_input.json collections), a synthetic 5 Mbp genome (i.i.d. uniform ACGT,
PCG64 seed 2 + rank), 3 000 genes laid out every ~1 667 bp on alternating strands,
promoters = Gene.promoter_region(300, 50).  One step = the hot path over one
genome: genome-wide both-strand scan -> soft-max background histogram + moments ->
mean/std -> posterior of regulation of all 3 000 promoters.

value      bp*motif/s with the packed genome already resident in HBM (per-step CUDA
           events; L2 flushed between steps, flush outside the timed span)
e2e        the same through the public API from pinned HOST bytes (one C call): H2D of the
           ASCII genome + pack + scan + posteriors, results landing in pinned host memory
roofline   the tiled stats-scan kernel (dominant): algorithmic bytes and lookup-adds
           per launch over its CUDA-event duration (cgbk_profile_*)
at_scale   (N = 1) the two tiled kernels on a 1 Gbp synthetic sequence, i.e. at the
           size where launch ramp and tile quantisation no longer matter
N > 1      one genome per rank (BASELINE configs[2] style sharding, no collective)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GENOME_BP = 5_000_000
N_GENES = 3000
UP, DOWN = 300, 50
M_LEN = 22
METRIC = "PSWM scan bp*motif/s (both strands): background fit + promoter posteriors"
UNIT = "bp*motif/s"
WORKLOAD = ("config2: LexA 22-bp PSWM, synthetic 5 Mbp genome per GPU, 3000 promoters x 350 bp "
            "(background histogram+moments, then posteriors)")


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


# ----------------------------------------------------------------------------- CPU arms
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


_REF_SEQ = None     # genome bytes, inherited by the forked workers (not pickled per step)


def _ref_worker(args):
    promoters, bg_slice = args
    d, site_lists, weights = lexa()
    t, nwin, nprom = cpu_config2(_REF_SEQ, promoters, bg_slice, d, site_lists, weights)
    return nwin, nprom


def run_reference(args):
    """--impl reference: the reference's CPU path (Python port, oracle/py_port.py; the
    reference tree itself cannot travel to the GPU box) on all host cores, on a bounded
    sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import c_oracle
    from cgb_b200.gene_regions import GeneTable, python_slice_bounds
    c_oracle.build()
    cores = os.cpu_count() or 1
    global _REF_SEQ
    _REF_SEQ = synthetic_genome(GENOME_BP, 2).tobytes()
    gs, ge, gstrand = gene_layout(GENOME_BP, N_GENES)
    gt = GeneTable(gs, ge, gstrand, GENOME_BP)
    ps, pe = python_slice_bounds(*gt.promoter_region_location(UP, DOWN), GENOME_BP)
    # per step: a 1/10 sample of config 2 (500 kbp of background windows + 300 promoters),
    # split evenly over the cores
    sample_bp, sample_prom = GENOME_BP // 10, N_GENES // 10
    per_bp = sample_bp // cores
    jobs = []
    for c in range(cores):
        lo = c * per_bp
        proms = [(int(ps[i]), int(pe[i])) for i in range(c, sample_prom, cores)]
        jobs.append((proms, (lo, lo + per_bp + M_LEN - 1)))
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_worker, jobs)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    nwin = sum(r[0] for r in res)
    total = float(np.sum(times))
    value = nwin * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "sample": "per step 1/10 of config 2: %d background windows + %d promoters" % (nwin, sample_prom)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "1/10 of config 2 per step, %d processes" % cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def at_scale(M, dev, bp, reps=3):
    """The two tiled kernels on one `bp`-base synthetic sequence (kernel CUDA-event time)."""
    import torch

    from cgb_b200 import PackedGenome, _cabi
    gen = torch.Generator(device=dev)
    gen.manual_seed(4)
    g = PackedGenome([bp], dev)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    ascii_all = torch.empty(bp, dtype=torch.uint8, device=dev)
    chunk = 1 << 28
    for off in range(0, bp, chunk):
        n = min(chunk, bp - off)
        ascii_all[off:off + n] = lut[torch.randint(0, 4, (n,), device=dev, generator=gen)]
    g.pack_into(0, ascii_all)
    torch.cuda.synchronize()
    del ascii_all
    nwin = g.total_windows(M.length)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    hbm_peak, sm_max, _ = peaks()
    issue_peak = sms * 128 * sm_max * 1e6
    out = {"bp": bp, "motif_len": M.length}
    thr = M.threshold()
    for mode in ("hit_scan", "stats_scan"):
        times = []
        for rep in range(reps + 2):
            if mode == "hit_scan":
                M.scan_hits(g, thr)
            else:
                M.background_stats(g)
                torch.cuda.synchronize()
            if rep >= 2:
                times.append(_cabi.profile_last_ms())
        ms = float(np.mean(times))
        rate = nwin / (ms * 1e-3)
        out[mode] = {"kernel_ms": ms, "bp_motif_per_s": rate,
                     "issue_frac": rate * 2 * M.length / issue_peak,
                     "hbm_frac": nwin * 0.25 / (ms * 1e-3) / 1e9 / hbm_peak}
        if mode == "hit_scan":
            out[mode].update(M.last_scan_counters)
    # what actually bounds both kernels (one `ncu --set full` capture of this leg,
    # profiles/r1_ncu_scans_1gbp.csv: l1tex__data_pipe_lsu_wavefronts_mem_shared, % of peak)
    out["smem_wavefront_pct_of_peak_ncu"] = {"hit_scan": 89.4, "stats_scan": 88.5}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    from cgb_b200 import GeneTable, PackedGenome, PSSMModel, SiteCollection, _cabi
    from tools.clock_sampler import ClockSampler

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

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
    n_post_windows = int(desc[1].sum().item())
    bg = M.background_stats(g, n_bins=256, keep_scores=True)
    M.fit_background(bg=bg, sync=True)            # also scores the model's own sites once (mu_m, std_m)
    post = torch.empty(N_GENES, dtype=torch.float64, device=dev)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    _cabi.profile_enable(True)

    def step():
        M.background_stats(g, out=bg, keep_scores=True)
        M.fit_background(bg=bg, sync=False)
        M.binding_probabilities(g, None, None, prior, alpha, out=post, region_tensors=desc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    post_host = torch.empty(N_GENES, dtype=torch.float64).pin_memory()
    stats_host = torch.empty(_cabi.BG_COUNT, dtype=torch.float64).pin_memory()

    def e2e_step():
        # the reference-facing host-buffer call (cgbk_regulation_pipeline through
        # PSSMModel.regulation_pipeline): pinned host ASCII + promoter slices in, host
        # posteriors / background moments / histogram out
        return M.regulation_pipeline(pinned, ps, pe, prior, alpha, n_bins=256)

    def e2e_step_multi_call():
        # the same through the individual batched calls (pack, stats, posteriors)
        gg = PackedGenome.from_sequences(None, dev, pinned=[pinned])
        b = M.background_stats(gg, n_bins=256, keep_scores=True)
        M.fit_background(bg=b, sync=False)
        p = M.binding_probabilities(gg, ps, pe, prior, alpha, out=post)
        post_host.copy_(p, non_blocking=True)
        stats_host.copy_(b["stats"], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return post_host, stats_host

    for _ in range(args.warmup):
        step()
        flush.zero_()
    for _ in range(max(3, args.warmup)):
        e2e_step()
    barrier()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = []
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        ev[i][0].record()
        step()
        ev[i][1].record()
        kern_ms.append(_cabi.profile_last_ms())     # syncs on the scan kernel's stop event
        flush.zero_()                               # L2 flush, outside the timed span
    barrier()
    wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    eager_total_ms = float(np.sum(step_ms))
    total_ms = eager_total_ms
    launches_per_step = 2                           # stats scan (exact tiles + moment reduction inside), posteriors
    info = _cabi.last_launch_info()

    # ---- the same K steps replayed from a CUDA graph (the step is launch-bound: one memset and
    # two kernels, ~37 us in total).  `value` is taken from these replays; the eager loop above
    # provides the per-kernel CUDA-event time for the roofline.
    graph_mode = "eager"
    if not args.no_graph:
        try:
            _cabi.profile_enable(False)             # event records are not wanted inside the graph
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                step()
            for _ in range(3):
                graph.replay()
                flush.zero_()
            barrier()
            for i in range(args.steps):
                ev[i][0].record()
                graph.replay()
                ev[i][1].record()
                flush.zero_()
            barrier()
            total_ms = float(np.sum([a.elapsed_time(b) for a, b in ev]))
            graph_mode = "cuda_graph"
        except Exception as exc:                    # keep the eager number
            graph_mode = "eager (graph capture failed: %s)" % str(exc)[:80]
            total_ms = eager_total_ms
        finally:
            _cabi.profile_enable(True)

    # ---- end to end through the public API, from pinned host memory (no profiling events)
    _cabi.profile_enable(False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    for _ in range(3):
        e2e_step_multi_call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step_multi_call()
    torch.cuda.synchronize()
    e2e_multi_s = time.perf_counter() - t0
    _cabi.profile_enable(True)
    clocks = sampler.stop() if rank == 0 else None

    tt = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(tt[0].item()), float(tt[1].item())

    if rank == 0:
        units = n_windows * world * args.steps          # bp*motif (1 motif), all ranks
        value = units / (total_ms * 1e-3)
        e2e_value = units / (e2e_ms * 1e-3)
        hbm_peak, sm_max, peak_src = peaks()
        k_ms = float(np.mean(kern_ms))
        alg_bytes = n_windows * 0.25                    # 0.25 B per bp per pass
        lookup_adds = n_windows * 2 * M_LEN             # 2m lookup-adds per bp*motif
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        issue_peak = sms * 128 * sm_max * 1e6
        roofline = {
            "bound": "hbm", "kernel": "fast_scan_kernel<G=6, stats> on one 5 Mbp genome",
            "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload, one
            # `ncu --set full` capture (profiles/r1_ncu_stats_bench.csv): 1.956 MB read + 10.8 KB written per launch
            "traffic": 1966592, "algorithmic_bytes": alg_bytes,
            "peak_source": peak_src, "kernel_ms": k_ms,
            "issue": {"achieved": lookup_adds / (k_ms * 1e-3), "peak": issue_peak, "unit": "lookup-add/s",
                      "frac": lookup_adds / (k_ms * 1e-3) / issue_peak,
                      "note": "north_star scan roofline = slower of HBM bytes and 2m lookup-adds at "
                              "SMs*128 lanes*f_max; the issue bound binds"},
        }
        cpu_baseline = None
        scale = None
        if world == 1 and not args.no_cpu:
            from oracle import c_oracle
            c_oracle.build()
            s2, e2 = np.clip(ps, 0, GENOME_BP), np.clip(pe, 0, GENOME_BP)
            t_cpu, nwin_cpu, nprom_cpu = cpu_config2(host.tobytes(), [(int(a), int(b)) for a, b in zip(s2, e2)],
                                                     None, d, site_lists, weights)
            cpu_baseline = {"value": nwin_cpu / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": "the full config-2 step once (%d windows + %d promoters, %.1f s)"
                                      % (nwin_cpu, nprom_cpu, t_cpu)}
        if world == 1 and not args.no_scale:
            del flush
            torch.cuda.empty_cache()
            scale = at_scale(M, dev, args.scale_bp)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32 fixed point (2^-24 for this motif) + f64 moments/posterior",
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "genome_bp": GENOME_BP, "promoters": N_GENES, "posterior_windows": n_post_windows,
                       "l2": "flushed between steps (512 MiB memset, outside the timed span)",
                       "timing": "per-step CUDA events, max over ranks; value: %s; roofline kernel time: eager steps"
                                 % graph_mode,
                       "eager_ms_per_step": eager_total_ms / args.steps, "tiles": info["tiles"],
                       "lane_bases": info["lane_bases"], "sharding": "one genome per rank, no collective"},
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pinned.numel()) + 12 * N_GENES + 16,
                    "d2h_bytes_per_step": N_GENES * 8 + 64 + 256 * 8, "ms_per_step": e2e_ms / args.steps,
                    "call": "PSSMModel.regulation_pipeline -> cgbk_regulation_pipeline (host buffers)",
                    "d2h": "posteriors + stats + histogram land in the pinned host buffer inside the call "
                           "(stored over PCIe by the finalising kernels; a D2H copy for pageable buffers)",
                    "multi_call_ms_per_step": 1e3 * e2e_multi_s / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            # SURVEY.md 8d's second figure: promoters / time of (background fit + posterior kernels)
            "promoter_posteriors_per_s": world * N_GENES / (total_ms * 1e-3 / args.steps),
            "clocks": clocks, "wall_s_timed_loop": wall, "at_scale": scale,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-scale", action="store_true", help="skip the at_scale leg")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--scale-bp", type=int, default=1_000_000_000)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
