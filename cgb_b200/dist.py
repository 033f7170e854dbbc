"""Multi-GPU sharding of the hot path: one process per GPU (torch.distributed).

The reference has no parallelism at all; the data-parallel structure
used here:

* whole genomes (BASELINE configs[2]): assigned to ranks largest-first, no collective
  at all — every genome has its own background fit;
* chunks of one long sequence (BASELINE configs[3]): contiguous chunks with an
  (m - 1)-base halo on the right so that every window is scored by exactly one rank,
  and ONE all-reduce (SUM) of the background record {n, sum, sumsq, histogram}; the
  mean / std are then recomputed on every rank (``cgbk_background_finalize``).

Hit lists are per-rank results; positions are shifted back to global coordinates
with the chunk offset.  Nothing here touches CUDA directly: the all-reduce runs on
whatever backend the process group has (NCCL over NVLink on the B200 box, gloo in the
CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi


def assign_genomes(sizes, world_size):
    """Greedy largest-first assignment of genomes to ranks; returns a list of index
    lists, one per rank (deterministic)."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(sizes[i])
    return [sorted(x) for x in out]


def chunk_bounds(n_bases, m, world_size):
    """Chunks [start, end) of a sequence of n_bases for each rank: windows are split
    evenly, each chunk carries the m-1 bases its last windows need (halo).  Returns
    (starts, ends, window_counts); chunk r owns windows [starts[r], starts[r]+count[r])."""
    nwin = max(0, n_bases - m + 1)
    base = nwin // world_size
    extra = nwin % world_size
    counts = np.array([base + (1 if r < extra else 0) for r in range(world_size)], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
    ends = np.where(counts > 0, starts + counts + m - 1, starts)
    return starts, ends, counts


def plan_motif_shards(costs, n_groups):
    """Motif indices for each of ``n_groups`` ranks: longest-processing-time-first over the
    estimated costs (deterministic), every motif exactly once, each list ascending."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * n_groups
    out = [[] for _ in range(n_groups)]
    for i in order:
        r = min(range(n_groups), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += float(costs[i])
    return [sorted(x) for x in out]


def plan_grid(world_size, n_bases, min_shard_bases=500_000_000):
    """How a motif set x one long sequence (BASELINE configs[4]) is laid over ``world_size`` ranks:
    ``S`` sequence shards x ``M`` motif groups, S * M = world_size.  The sequence is cut as far as
    the shards stay at or above ``min_shard_bases`` (a scan launch has a fixed cost of tens of
    microseconds: shards much smaller than that stop scaling), the ranks left over split the motif
    set (SURVEY 8e names both partitions).  Rank r scans sequence shard r % S for motif group
    r // S; ONE all-reduce over [all motifs][record] -- rows a rank did not scan stay zero --
    assembles every motif's record on every rank.  Returns (S, M)."""
    S = 1
    for d in range(1, world_size + 1):
        if world_size % d == 0 and (n_bases + d - 1) // d >= min_shard_bases:
            S = d
    return S, world_size // S


def allreduce_background(bg, group=None):
    """Sum the raw background record of every rank in place and recompute mean / std.

    ``bg``: result of ``PSSMModel.background_stats`` (device tensors, or CPU tensors in
    the gloo tests).  One collective: float64 buffer [n, sum, sumsq | histogram]
    (histogram counts are exact in float64 up to 2**53)."""
    stats, hist = bg["stats"], bg["hist"]
    buf = torch.cat([stats[:3], hist.to(torch.float64)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    stats[:3] = buf[:3]
    hist.copy_(buf[3:].round().to(hist.dtype))
    if stats.is_cuda:
        _cabi.check(_cabi.lib().cgbk_background_finalize(
            stats.data_ptr(), torch.cuda.current_stream(stats.device).cuda_stream))
    else:
        finalize_stats_host(stats)
    return bg


def allreduce_batch(stats, hist, group=None):
    """The same for a motif batch: ``stats`` float64 [K][8] and ``hist`` int64 [K][n_bins] (or
    None) of ``motif_batch.scan_batch(..., to_host=False)`` are summed over the ranks with ONE
    collective -- a float64 buffer [K][3 + n_bins] -- and mean / std of all K records recomputed
    (``cgbk_background_finalize_batch``).  In place; returns the buffer size in bytes."""
    cols = [stats[:, :3]]
    if hist is not None:
        cols.append(hist.to(torch.float64))
    buf = torch.cat(cols, dim=1).contiguous()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    stats[:, :3] = buf[:, :3]
    if hist is not None:
        hist.copy_(buf[:, 3:].round().to(hist.dtype))
    if stats.is_cuda:
        if not stats.is_contiguous():
            raise ValueError("stats must be contiguous")
        _cabi.check(_cabi.lib().cgbk_background_finalize_batch(
            stats.data_ptr(), int(stats.shape[0]), torch.cuda.current_stream(stats.device).cuda_stream))
    else:
        for k in range(stats.shape[0]):
            finalize_stats_host(stats[k])
    return buf.numel() * 8


def finalize_stats_host(stats):
    """mean / std (ddof = 0) from n, sum, sumsq — the host mirror of
    cgbk_background_finalize, used on CPU tensors."""
    n, s1, s2 = float(stats[_cabi.BG_N]), float(stats[_cabi.BG_SUM]), float(stats[_cabi.BG_SUMSQ])
    if n > 0:
        mean = s1 / n
        var = max(s2 / n - mean * mean, 0.0)
        stats[_cabi.BG_MEAN] = mean
        stats[_cabi.BG_STD] = var ** 0.5
    else:
        stats[_cabi.BG_MEAN] = float("nan")
        stats[_cabi.BG_STD] = float("nan")
    return stats


def gather_hits(pos, strand, score, chunk_start, group=None):
    """All ranks' hit arrays (numpy) concatenated in global coordinates on every rank,
    sorted by (pos, +strand first)."""
    obj = (np.asarray(pos) + int(chunk_start), np.asarray(strand), np.asarray(score))
    world = dist.get_world_size(group)
    gathered = [None] * world
    dist.all_gather_object(gathered, obj, group=group)
    p = np.concatenate([g[0] for g in gathered])
    s = np.concatenate([g[1] for g in gathered])
    v = np.concatenate([g[2] for g in gathered])
    order = np.lexsort((-s, p))
    return p[order], s[order], v[order]
