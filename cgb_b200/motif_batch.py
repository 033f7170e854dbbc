"""Many PSWMs at once: batched model tables and the batched (fused) scan.

The reference builds one ``PSSMModel`` per genome inside its ``go()`` loop
(cgb/__init__.py:673-686, pssm_model.py:31-35) and scores one motif per call
(pssm_model.py:103-133).  For a motif set (BASELINE configs[4]: 500 JASPAR-scale motifs over a
Gbp sequence) or a genome batch with several PSWMs each (configs[2]) that is one device
allocation, one synchronous upload and a handful of launches *per model*.  Here

* :class:`ModelBatch` puts the device tables of K models into ONE block with ONE upload
  (``cgbk_model_batch_create``);
* :func:`scan_batch` runs K models over a packed genome in ONE library call: per model a
  single pass that yields the background record (moments, optional histogram;
  genome.py:215-226) AND the thresholded sites of both strands (genome.py:433-445), the
  latter as compact 8-byte records in one shared buffer.  No allocation and no host
  synchronisation per model; the host sees one counters block per sub-batch.

A rank's shard of a longer sequence passes ``win_cap`` = the number of windows it owns: the
shard carries the halo of the longest motif, every motif stops at the shard's own windows
(SURVEY.md §8e).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi

HIT_DTYPE = np.dtype([("gpos", "<i8"), ("strand", "<i4"), ("score", "<f4")])


def unpack_hits(rec):
    """Compact records (uint64 ndarray, ``cgbk_hit8``) -> (gpos int64, strand int32, score float32)."""
    rec = np.asarray(rec, dtype=np.uint64)
    gpos = (rec >> np.uint64(33)).astype(np.int64)
    strand = (1 - 2 * ((rec >> np.uint64(32)) & np.uint64(1)).astype(np.int32)).astype(np.int32)
    score = (rec & np.uint64(0xffffffff)).astype(np.uint32).view(np.float32)
    return gpos, strand, score


class SplitSites(object):
    """One model's ordered site list in the 6-byte split form (``cgbk_split_sites``): ``local``
    uint16 = (offset inside the 4 096-base block) << 1 | minus strand, ``score`` float32,
    ``block_start`` int32 [blocks + 1].  Views of pinned staging memory when handed to a ``sink``
    (valid until it returns); :meth:`copy` detaches."""
    BLOCK_SHIFT = 12

    def __init__(self, local, score, block_start):
        self.local, self.score, self.block_start = local, score, block_start

    def __len__(self):
        return len(self.local)

    def copy(self):
        return SplitSites(self.local.copy(), self.score.copy(), self.block_start.copy())

    def positions(self):
        """Offsets into the packed buffer (int64), ascending."""
        counts = np.diff(self.block_start.astype(np.int64))
        base = np.repeat(np.arange(len(counts), dtype=np.int64) << self.BLOCK_SHIFT, counts)
        return base + (self.local >> np.uint16(1)).astype(np.int64)

    def strands(self):
        return (1 - 2 * (self.local & np.uint16(1)).astype(np.int32)).astype(np.int32)

    def records(self):
        """The same sites as ``cgbk_hit8`` records (uint64), bit for bit what the 8-byte path delivers."""
        return (self.positions().astype(np.uint64) << np.uint64(33)) | \
            ((self.local & np.uint16(1)).astype(np.uint64) << np.uint64(32)) | \
            self.score.view(np.uint32).astype(np.uint64)


class ModelBatch:
    """Device tables of many :class:`PSSMModel` in one block.  The models' handles are borrowed
    from the batch, which they keep alive; models that already have tables are left alone."""

    def __init__(self, models, device=None):
        models = list(models)
        todo = [M for M in models if M._handle is None]
        self.models = models
        self._batch = C.c_void_p()
        if not todo:
            return
        dev = torch.device(device) if device is not None else todo[0]._dev()
        rows = []
        for M in todo:
            if sorted(M.alphabet) != ["A", "C", "G", "T"]:
                raise ValueError("PSSM has wrong alphabet: %s - Use only with DNA motifs" % M.alphabet)
            M._device = dev
            rows.append(np.asarray(M.pssm.rows(), dtype=np.float64))
        flat = np.ascontiguousarray(np.concatenate([r.reshape(-1) for r in rows]))
        ms = np.ascontiguousarray(np.array([len(r) for r in rows], dtype=np.int32))
        lib = _cabi.lib()
        stream = torch.cuda.current_stream(dev).cuda_stream
        _cabi.check(lib.cgbk_model_batch_create(
            flat.ctypes.data_as(C.POINTER(C.c_double)), ms.ctypes.data_as(C.POINTER(C.c_int)), len(todo),
            dev.index or 0, stream, C.byref(self._batch)))
        for k, M in enumerate(todo):
            M._handle = C.c_void_p(lib.cgbk_model_batch_get(self._batch, k))
            M._batch_owner = self          # the tables live as long as any model refers to them

    def __del__(self):
        try:
            if getattr(self, "_batch", None) is not None and self._batch.value:
                _cabi.lib().cgbk_model_batch_destroy(self._batch)
                self._batch = C.c_void_p()
        except Exception:
            pass


class _ScanBuffers:
    """Device and pinned host buffers of :func:`scan_batch`, kept on the genome and reused between
    calls.  Two hit buffers: while the host receives the sites of one sub-batch, the next is scanned."""

    def __init__(self, dev, n_models, n_bins, hit_cap, work_bytes, host_hits, split_blocks=0):
        self.n_models, self.n_bins, self.hit_cap, self.work_bytes = n_models, n_bins, hit_cap, work_bytes
        self.split_blocks = split_blocks
        # stats | histograms | counters back to back: the library clears them with one memset
        n_rec = n_models * _cabi.BG_COUNT + n_models * n_bins + (n_models + 1) * _cabi.CTR_COUNT
        self.record = torch.empty(n_rec, dtype=torch.int64, device=dev)
        self.host_record = torch.empty(n_rec, dtype=torch.int64).pin_memory()
        n_buf = 2 if host_hits else 1
        self.hits = [torch.empty(max(hit_cap, 1), dtype=torch.int64, device=dev) for _ in range(n_buf)]
        if host_hits and split_blocks:
            # split transfer: the 8-byte records never leave the device; 2 + 4 bytes per site and one
            # int32 per 4 096-base block and model go to the host instead
            nb1 = split_blocks + 1
            self.host_hits = [True] * n_buf
            self.d_local = [torch.empty(max(hit_cap, 1), dtype=torch.int16, device=dev) for _ in range(n_buf)]
            self.d_score = [torch.empty(max(hit_cap, 1), dtype=torch.float32, device=dev) for _ in range(n_buf)]
            self.d_blocks = [torch.empty(n_models * nb1, dtype=torch.int32, device=dev) for _ in range(n_buf)]
            self.h_local = [torch.empty(max(hit_cap, 1), dtype=torch.int16).pin_memory() for _ in range(n_buf)]
            self.h_score = [torch.empty(max(hit_cap, 1), dtype=torch.float32).pin_memory() for _ in range(n_buf)]
            self.h_blocks = [torch.empty(n_models * nb1, dtype=torch.int32).pin_memory() for _ in range(n_buf)]
            self.d_cum = [torch.empty(n_models + 1, dtype=torch.int64, device=dev) for _ in range(n_buf)]
            self.h_cum = [torch.empty(n_models + 1, dtype=torch.int64).pin_memory() for _ in range(n_buf)]
        else:
            self.host_hits = [torch.empty(max(hit_cap, 1), dtype=torch.int64).pin_memory() for _ in range(n_buf)] \
                if host_hits else []
        self.copied = [None] * n_buf                  # event: the D2H out of hits[p] has finished
        self.copy_stream = torch.cuda.Stream(device=dev, priority=-1) if host_hits else None
        self.work = torch.empty(work_bytes, dtype=torch.uint8, device=dev)


# one set of scratch buffers per device, shared by all genomes scanned there (calls on one device
# are expected to come from one thread at a time; the buffers are stream-ordered on the current
# stream and host-synchronised before they are handed out again)
_BUFFER_CACHE = {}


def _buffers(genome, n_models, n_bins, hit_cap, work_bytes, host_hits, split_blocks=0):
    key = genome.device.index or 0
    b = _BUFFER_CACHE.get(key)
    if (b is None or b.n_models < n_models or b.n_bins != n_bins or b.hit_cap < hit_cap or b.work_bytes < work_bytes
            or (host_hits and not b.host_hits) or (host_hits and b.split_blocks != split_blocks)):
        _BUFFER_CACHE.pop(key, None)
        b = None                                          # free the old buffers first
        b = _ScanBuffers(genome.device, n_models, n_bins, hit_cap, work_bytes, host_hits, split_blocks)
        _BUFFER_CACHE[key] = b
    return b


def release_buffers(device=None):
    """Drop the cached scratch buffers (all devices, or one)."""
    if device is None:
        _BUFFER_CACHE.clear()
    else:
        _BUFFER_CACHE.pop(torch.device(device).index or 0, None)


def expected_hits(model, n_windows):
    """Sites expected at the Patser operating point (FPR = 2^-IC per strand, pssm_model.py:62-70)
    under the uniform background, both strands."""
    return 2.0 * n_windows * 2.0 ** (-max(float(model.IC), 0.0))


# measured on B200, 1 Gbp, fused scan (profiles/r2_bench_line_n1.json, r2_site_pass_times.log):
# scan kernel ms per Gbp by group count G = ceil(m / 4), site pass ms = fixed + per site
_SCAN_MS_PER_GBP = {1: 1.1, 2: 1.2, 3: 1.32, 4: 1.40, 5: 1.61, 6: 1.79, 7: 2.18, 8: 2.42}
_SITE_MS_FIXED, _SITE_MS_PER_SITE = 0.095, 1.7e-8


def estimated_cost_ms(model, n_windows):
    """Device time of one motif's fused scan + site pass over ``n_windows`` windows (for balancing
    motif groups over ranks; a model of the measurements above, not a promise)."""
    g = (int(model.length) + 3) // 4
    return _SCAN_MS_PER_GBP[min(max(g, 1), 8)] * n_windows * 1e-9 + _SITE_MS_FIXED + \
        _SITE_MS_PER_SITE * expected_hits(model, n_windows)


def interleave_order(cost_ms, sites):
    """Processing order for a batch whose sites go to the host while the next models are scanned:
    models whose site lists are long for their scan time (PCIe-heavy) and models that are mostly scan
    time are interleaved so that, all along the batch, sites sent and device time spent keep the
    batch's overall ratio -- the copies of one sub-batch then take about as long as the scan of the
    next, instead of a run of dense motifs waiting on PCIe and a run of sparse ones leaving it idle.
    Deterministic; returns a permutation (list of indices)."""
    n = len(cost_ms)
    tot_c, tot_s = float(sum(cost_ms)), float(sum(sites))
    if n < 3 or tot_c <= 0 or tot_s <= 0:
        return list(range(n))
    ratio = tot_s / tot_c
    heavy = sorted((i for i in range(n) if sites[i] > ratio * cost_ms[i]), key=lambda i: (-float(sites[i]), i))
    light = sorted((i for i in range(n) if not sites[i] > ratio * cost_ms[i]), key=lambda i: (float(sites[i]), i))
    order, cs, cc = [], 0.0, 0.0
    while heavy or light:
        if heavy and (not light or cs <= ratio * cc):
            i = heavy.pop(0)
        else:
            i = light.pop(0)
        order.append(i)
        cs += float(sites[i])
        cc += float(cost_ms[i])
    return order


def scan_batch(models, genome, thresholds="patser", n_bins=256, win_cap=None, revseq_order=True,
               hit_cap=None, to_host=True, sink=None, recycle=False, kernel_ms=None, site_ms=None, split=False,
               interleave=None):
    """Background record and thresholded sites of every model over ``genome`` (a PackedGenome);
    see :func:`_scan_batch_in_order` for the arguments.  ``interleave`` (default: on for batches of 8
    or more models whose sites go to the host): the models are PROCESSED in :func:`interleave_order`
    so that the site copies overlap the scans all along the batch; everything is returned (and
    handed to the ``sink``) under the caller's indices."""
    models = list(models)
    K = len(models)
    want_hits = thresholds is not None
    if interleave is None:
        interleave = want_hits and to_host and K >= 8
    if not (interleave and want_hits and K > 2):
        return _scan_batch_in_order(models, genome, thresholds, n_bins, win_cap, revseq_order, hit_cap, to_host, sink,
                                    recycle, kernel_ms, site_ms, split)
    from .pssm_model import prepare_thresholds
    if isinstance(thresholds, str):
        if thresholds != "patser":
            raise ValueError
        prepare_thresholds(models)
        thresholds = [M.threshold() for M in models]
    if len(thresholds) != K:
        raise ValueError("one threshold per model")
    nwin = genome.total_windows(min(M.length for M in models))
    if win_cap is not None:
        nwin = min(nwin, int(win_cap) * genome.n_seq)
    order = interleave_order([estimated_cost_ms(M, nwin) for M in models], [expected_hits(M, nwin) for M in models])
    km = np.zeros(K) if kernel_ms is not None else None
    sm = np.zeros(K) if site_ms is not None else None
    res = _scan_batch_in_order([models[i] for i in order], genome, [thresholds[i] for i in order], n_bins, win_cap,
                               revseq_order, hit_cap, to_host, (lambda k, rec: sink(order[k], rec)) if sink else None,
                               recycle, km, sm, split)
    inv = np.empty(K, dtype=np.int64)
    inv[order] = np.arange(K)
    if kernel_ms is not None:
        kernel_ms += km[inv]
    if site_ms is not None:
        site_ms += sm[inv]
    out = dict(res)
    for key in ("stats", "hist", "counters"):
        if res[key] is not None:
            out[key] = res[key][inv] if isinstance(res[key], np.ndarray) else res[key][torch.from_numpy(inv).to(res[key].device)]
    if res["hits"] is not None:
        out["hits"] = [res["hits"][int(j)] for j in inv]
    out["order"] = order
    return out


def _scan_batch_in_order(models, genome, thresholds="patser", n_bins=256, win_cap=None, revseq_order=True,
                         hit_cap=None, to_host=True, sink=None, recycle=False, kernel_ms=None, site_ms=None,
                         split=False):
    """Background record and thresholded sites of every model over ``genome`` (a PackedGenome).

    ``thresholds``: ``"patser"`` (each model's own, pssm_model.py:72-77), a sequence of floats, or
    ``None`` for background records only.  ``n_bins`` = 0 runs without histograms.
    Returns a dict:

      stats   float64 [K][8] (n, sum, sumsq, mean, std), hist int64 [K][n_bins] or None
      hits    list of K compact uint64 arrays (see :func:`unpack_hits`), each ordered by
              (position, + strand first) -- the scan delivers them in that order; positions are
              offsets into the packed buffer.  ``None`` without thresholds
      counters  int64 [K][8];  launches: kernels launched

    ``to_host`` (default): numpy arrays.  The models go through the device in sub-batches sized
    to the hit buffer; each sub-batch's sites cross PCIe on a copy stream (pinned staging) while
    the next sub-batch is scanned.  ``sink(k, records)``: called with model k's records as a
    view of the pinned staging buffer instead of a fresh array (valid until it returns).
    ``split`` (with ``to_host``): the sites travel in the 6-byte split form (:class:`SplitSites`:
    uint16 block-local offset + strand, float32 score, one int32 per 4 096-base block) instead of
    8-byte records -- a quarter fewer bytes over PCIe, which is what bounds a large scan end to end;
    ``hits`` / the ``sink`` then get :class:`SplitSites` objects (``.records()`` rebuilds the 8-byte
    records exactly).
    ``to_host=False``: device tensors; the sites stay in the device hit buffer, so the whole
    batch must fit ``hit_cap`` (ValueError otherwise) -- unless ``recycle``: then the buffer is
    reused from sub-batch to sub-batch and ``hits`` holds the per-model COUNTS only (a consumer
    on the device would take each sub-batch's sorted lists before the next overwrites them; the
    benchmark's device-resident leg).
    ``kernel_ms``: float array [K]; with ``_cabi.profile_enable(2)`` every model's scan-kernel time
    (CUDA events around the launch) is ADDED to it; ``site_ms`` likewise the time of every model's
    site pass (candidate re-score + ordered emission)."""
    from .pssm_model import prepare_thresholds
    models = list(models)
    K = len(models)
    if K == 0:
        raise ValueError("no models")
    dev = genome.device
    ModelBatch(models, dev)                               # tables of the models that have none yet
    lib = _cabi.lib()
    want_hits = thresholds is not None
    if want_hits and isinstance(thresholds, str):
        if thresholds != "patser":
            raise ValueError
        prepare_thresholds(models)
        thresholds = [M.threshold() for M in models]
    thr = np.ascontiguousarray(thresholds, dtype=np.float64) if want_hits else None
    if want_hits and len(thr) != K:
        raise ValueError("one threshold per model")
    nwin_all = genome.total_windows(min(M.length for M in models))
    if win_cap is not None:
        nwin_all = min(nwin_all, int(win_cap) * genome.n_seq)
    # expected sites per model; explicit thresholds below the Patser point are only guessed (the
    # overflow path below re-runs with the real count)
    exp = np.array([expected_hits(M, nwin_all) * 1.25 + 4096 for M in models]) if want_hits else np.zeros(K)
    if hit_cap is None:
        hit_cap = 0
        if want_hits:
            want = exp.sum() if (not to_host and not recycle) else max(exp.max() * 1.5, 1 << 20)
            hit_cap = int(max(want, 1 << 20))
            if hit_cap * 16 > (1 << 30):
                # (cudaMemGetInfo costs 0.1 - 2 ms: asked only when the request is large)
                free_bytes, _ = torch.cuda.mem_get_info(dev)
                hit_cap = int(min(hit_cap, free_bytes // 4 // 16))
    # candidates are re-scored model by model: room for ONE model's worth (its sites plus the few
    # windows within the fixed-point error of the threshold)
    cand_cap = int(min(max(exp.max() * 1.25, 1 << 16), (1 << 30) - 1)) if want_hits else 0
    # scratch of the scan: moment partials, candidate flag plane, per-tile counts, candidate records
    work_bytes = int(lib.cgbk_scan_batch_workspace_bytes(genome.seqset, -1 if win_cap is None else int(win_cap),
                                                         cand_cap))
    handles_all = [M._model() for M in models]
    stream = torch.cuda.current_stream(dev)
    flags = _cabi.RC_REVSEQ_ORDER if revseq_order else _cabi.RC_MATRIX_ORDER
    BG, CT = _cabi.BG_COUNT, _cabi.CTR_COUNT
    if to_host:
        stats_out = np.zeros((K, BG))
        hist_out = np.zeros((K, n_bins), dtype=np.int64) if n_bins else None
    else:
        stats_out = torch.empty((K, BG), dtype=torch.float64, device=dev)
        hist_out = torch.empty((K, n_bins), dtype=torch.int64, device=dev) if n_bins else None
    ctr_out = np.zeros((K, CT), dtype=np.int64)
    hits_out = [None] * K if want_hits else None
    launches = 0
    host_hits = want_hits and to_host
    pending = None        # (first model, cum offsets, buffer index, buffers) of the sub-batch whose sites are in flight

    split_blocks = int(lib.cgbk_split_sites_blocks(genome.cap_bases)) if (split and want_hits and to_host) else 0

    def deliver(first, cum, p, buf):
        buf.copied[p].synchronize()
        if split_blocks:
            nb1 = split_blocks + 1
            loc, sc = buf.h_local[p].numpy().view(np.uint16), buf.h_score[p].numpy()
            blk = buf.h_blocks[p].numpy()
            for k in range(len(cum) - 1):
                seg = SplitSites(loc[cum[k]:cum[k + 1]], sc[cum[k]:cum[k + 1]], blk[k * nb1:(k + 1) * nb1])
                if sink is not None:
                    sink(first + k, seg)
                    hits_out[first + k] = len(seg)
                else:
                    hits_out[first + k] = seg.copy()
            return
        host = buf.host_hits[p].numpy().view(np.uint64)
        for k in range(len(cum) - 1):
            seg = host[cum[k]:cum[k + 1]]
            if sink is not None:
                sink(first + k, seg)
                hits_out[first + k] = len(seg)
            else:
                hits_out[first + k] = seg.copy()

    i = 0
    n_sub = 0
    while i < K:
        # sub-batch [i, j): as many models as the hit buffer is expected to hold
        j = i + 1
        acc = exp[i]
        while j < K and acc + exp[j] <= hit_cap:
            acc += exp[j]
            j += 1
        n = j - i
        buf = _buffers(genome, K, n_bins, hit_cap, work_bytes, host_hits, split_blocks)
        p = n_sub & 1 if host_hits else 0
        if host_hits and buf.copied[p] is not None and pending is not None and pending[2] == p:
            deliver(*pending)                         # cannot happen with two buffers, kept for safety
            pending = None
        hp = (C.c_void_p * n)(*[h.value for h in handles_all[i:j]])
        # stats | hist | counters of n models, adjacent, at the head of the record buffer
        a, b = n * BG, n * BG + n * n_bins
        rec = buf.record[:b + (n + 1) * CT]
        d_stats, d_hist, d_ctr = rec[:a], (rec[a:b] if n_bins else None), rec[b:]
        _cabi.check(lib.cgbk_scan_batch(
            hp, n, genome.struct_ref(), genome.seqset, -1 if win_cap is None else int(win_cap),
            thr[i:j].ctypes.data if want_hits else None, flags, None, None, int(n_bins),
            d_hist.data_ptr() if d_hist is not None else None, d_stats.data_ptr(),
            buf.hits[p].data_ptr() if want_hits else None, int(buf.hit_cap) if want_hits else 0, d_ctr.data_ptr(),
            buf.work.data_ptr(), buf.work.numel(), cand_cap, stream.cuda_stream))
        launches += int(_cabi.last_launch_info()["launches"])
        if not to_host and not want_hits:
            stats_out[i:j] = d_stats.view(torch.float64).view(n, BG)
            if n_bins:
                hist_out[i:j] = d_hist.view(n, n_bins)
            i = j
            continue
        # one D2H of the record (pinned), then the host knows the counts; the sites of the previous
        # sub-batch cross PCIe on the copy stream meanwhile
        host = buf.host_record[:rec.numel()]
        host.copy_(rec, non_blocking=True)
        stream.synchronize()
        h = host.numpy()
        ctr = h[b:].reshape(n + 1, CT)
        n_done = n
        if want_hits:
            cand_over = np.nonzero(ctr[:n, _cabi.CTR_CANDIDATES] > cand_cap)[0]
            cum_all = np.concatenate([[0], np.cumsum(ctr[:n, _cabi.CTR_HITS])])
            hit_over = np.nonzero(cum_all[1:] > buf.hit_cap)[0]
            bad = min([int(x[0]) for x in (cand_over, hit_over) if len(x)] or [n])
            if bad < n:
                # records were dropped from model `bad` on: keep the models before it and run the
                # rest again, with buffers grown to what model `bad` needs if it was the first
                n_done = bad
                need = float(max(ctr[bad, _cabi.CTR_HITS], ctr[bad, _cabi.CTR_CANDIDATES])) + 4096
                exp[i + bad] = max(exp[i + bad], need)
                if need > cand_cap:
                    if pending is not None:            # the buffers are about to be re-allocated
                        deliver(*pending)
                        pending = None
                    cand_cap = int(need)
                    work_bytes = int(lib.cgbk_scan_batch_workspace_bytes(
                        genome.seqset, -1 if win_cap is None else int(win_cap), cand_cap))
                if bad == 0:
                    if not to_host and not recycle:
                        raise ValueError("the sites of the batch do not fit hit_cap=%d (to_host=False keeps "
                                         "them on the device)" % hit_cap)
                    free_bytes, _ = torch.cuda.mem_get_info(dev)
                    if need * 16 * 2 > free_bytes + 16 * buf.hit_cap * 2:
                        raise MemoryError("model %d needs %d hit records, more than the device can hold"
                                          % (i, int(need)))
                    if pending is not None:
                        deliver(*pending)
                        pending = None
                    hit_cap = max(hit_cap, int(need))
                    cand_cap = max(cand_cap, int(need))
                    work_bytes = int(lib.cgbk_scan_batch_workspace_bytes(
                        genome.seqset, -1 if win_cap is None else int(win_cap), cand_cap))
                    release_buffers(dev)
                    del buf
                    continue
        if to_host:
            stats_out[i:i + n_done] = h[:a].view(np.float64).reshape(n, BG)[:n_done]
            if n_bins:
                hist_out[i:i + n_done] = h[a:b].reshape(n, n_bins)[:n_done]
        else:
            stats_out[i:i + n_done] = d_stats.view(torch.float64).view(n, BG)[:n_done]
            if n_bins:
                hist_out[i:i + n_done] = d_hist.view(n, n_bins)[:n_done]
        ctr_out[i:i + n_done] = ctr[:n_done]
        if kernel_ms is not None and n_done:
            kernel_ms[i:i + n_done] += np.array(_cabi.profile_batch_ms(n_done))
        if site_ms is not None and n_done and want_hits:
            site_ms[i:i + n_done] += np.array(_cabi.profile_batch_ms(2 * n))[n:n + n_done]
        if want_hits and n_done:
            cum = np.concatenate([[0], np.cumsum(ctr[:n_done, _cabi.CTR_HITS])]).astype(np.int64)
            used = int(cum[-1])
            if to_host:
                # D2H on the copy stream, behind the scan; delivered while the next sub-batch runs
                ready = torch.cuda.Event()
                ready.record(stream)
                with torch.cuda.stream(buf.copy_stream):
                    buf.copy_stream.wait_event(ready)
                    if split_blocks:
                        # per model: records -> (uint16, float32) + block index, on the copy stream
                        # (ONE launch for the sub-batch: it has to find its way between the scan kernels
                        # of the next sub-batch, which keep every SM busy)
                        nb1 = split_blocks + 1
                        if n_done and int(np.diff(cum).max()) >= (1 << 31):
                            raise ValueError("a site list of 2^31 or more records cannot take the split form "
                                             "(int32 block index): use split=False")
                        buf.h_cum[p][:n_done + 1].copy_(torch.from_numpy(cum))
                        buf.d_cum[p][:n_done + 1].copy_(buf.h_cum[p][:n_done + 1], non_blocking=True)
                        _cabi.check(lib.cgbk_split_sites_batch(
                            buf.hits[p].data_ptr(), buf.d_cum[p].data_ptr(), n_done, used, int(genome.cap_bases),
                            buf.d_local[p].data_ptr(), buf.d_score[p].data_ptr(), buf.d_blocks[p].data_ptr(),
                            buf.copy_stream.cuda_stream))
                        if used:
                            buf.h_local[p][:used].copy_(buf.d_local[p][:used], non_blocking=True)
                            buf.h_score[p][:used].copy_(buf.d_score[p][:used], non_blocking=True)
                        buf.h_blocks[p][:n_done * nb1].copy_(buf.d_blocks[p][:n_done * nb1], non_blocking=True)
                    elif used:
                        buf.host_hits[p][:used].copy_(buf.hits[p][:used], non_blocking=True)
                    buf.copied[p] = torch.cuda.Event()
                    buf.copied[p].record(buf.copy_stream)
                if pending is not None:
                    deliver(*pending)
                pending = (i, cum, p, buf)     # (the tuple keeps ITS buffers alive should the cache re-allocate)
                # the next scan must not overwrite hits[p ^ 1] before ITS copy is done: deliver()
                # above has just waited for it
            elif recycle:
                for k in range(n_done):
                    hits_out[i + k] = int(cum[k + 1] - cum[k])
            else:
                if i + n_done < K or i > 0:
                    raise ValueError("the sites of the batch do not fit hit_cap=%d (to_host=False keeps them "
                                     "on the device)" % hit_cap)
                for k in range(n_done):
                    hits_out[i + k] = buf.hits[p][cum[k]:cum[k + 1]]
            n_sub += 1
        i += n_done
    if pending is not None:
        deliver(*pending)
    return {"stats": stats_out, "hist": hist_out, "hits": hits_out, "counters": ctr_out, "launches": launches}


def hits_of(genome, rec):
    """Compact records of one model -> (seq_index, pos, strand, score) like ``PSSMModel.scan_hits``."""
    gpos, strand, score = unpack_hits(rec)
    seq = np.searchsorted(genome.seq_off, gpos, side="right") - 1
    return seq, gpos - genome.seq_off[seq], strand, score
