"""PSSMModel: weighted-mixture PSWM model scored on the GPU.

Drop-in for the reference's ``cgb/pssm_model.py:14-169`` (same constructor,
properties and methods, same return shapes and exceptions), plus the batched
entry points a GPU needs so that a genome is scanned in one launch sequence
instead of one Python call per gene:

  score_seqs            many sequences per launch      (score_seq, :103-133)
  scan_hits             thresholded both-strand hits   (genome.py:433-445)
  background_stats      genome-wide soft-max histogram + moments (genome.py:215-226)
  fit_background        build_bayesian_estimator from those moments (binding_model.py:71-92)
  binding_probabilities batched promoter posteriors    (binding_model.py:94-113, gene.py:129-146)

All scoring goes through libcgbk (include/cgbk.h); host code here only does
float64 motif algebra and descriptor bookkeeping.
"""
from __future__ import annotations

import ctypes as C
import math
import threading
from functools import cached_property

import numpy as np
import torch

from . import _cabi
from .binding_model import TFBindingModel
from .genome_buffer import HostPackedGenome, PackedGenome
from .gene_regions import python_slice_bounds
from .motif_matrix import PositionWeightMatrix, psum


def log2(x):
    """cgb/misc.py:14-16."""
    return math.log(x, 2)


class PSSMModel(TFBindingModel):
    def __init__(self, collections, weights,
                 background={'A': 0.25, 'C': 0.25, 'G': 0.25, 'T': 0.25}, alphabet="ACGT",
                 device=None):
        """pssm_model.py:31-35.  ``device``: CUDA device of the scoring tables
        (default: the current device at first use)."""
        super(PSSMModel, self).__init__(collections, background, alphabet)
        self._pwm = self._combine_pwms([c.pwm for c in collections], weights, alphabet)
        self._device = device
        self._handle = None
        self._dev_ms_m = None
        self._dev_ms_bg = None
        self._bg_stats_dev = None

    @classmethod
    def from_pwm(cls, pwm_values, sites=(), background=None, alphabet="ACGT", device=None):
        """Model from a ready PWM (letter -> m probabilities), e.g. a JASPAR
        matrix; ``sites`` (optional) feed score_self / the estimator."""
        class _Coll:
            pass
        coll = _Coll()
        coll.pwm = PositionWeightMatrix(alphabet, pwm_values)
        coll.sites = list(sites)
        bg = background or {'A': 0.25, 'C': 0.25, 'G': 0.25, 'T': 0.25}
        return cls([coll], [1.0], bg, alphabet, device)

    # ------------------------------------------------------------ properties
    @cached_property
    def pwm(self):
        """pssm_model.py:37-40."""
        return self._pwm

    @cached_property
    def length(self):
        """pssm_model.py:42-45."""
        return self.pwm.length

    @cached_property
    def pssm(self):
        """pssm_model.py:47-50."""
        return self._pwm.log_odds(self.background)

    @cached_property
    def rev_comp_pssm(self):
        """pssm_model.py:52-55."""
        return self.pssm.reverse_complement()

    @cached_property
    def IC(self):
        """pssm_model.py:57-60."""
        return self.pssm.mean()

    @cached_property
    def patser_threshold(self):
        """pssm_model.py:62-70 (Hertz & Stormo 1999: -log2(FPR) = IC)."""
        dist = self.pssm.distribution(precision=10**3)
        return dist.threshold_patser()

    def threshold(self, threshold_type='patser'):
        """pssm_model.py:72-77."""
        if threshold_type == 'patser':
            thr = self.patser_threshold
        else:
            raise ValueError
        return thr

    @cached_property
    def _score_bounds(self):
        """Default histogram range: [floor(min score), ceil(max score) + 1) (soft-max <= max + 1)."""
        return float(math.floor(self.pssm.min)), float(math.ceil(self.pssm.max) + 1.0)

    @cached_property
    def sites(self):
        """pssm_model.py:79-82."""
        return [site for coll in self._collection_set for site in coll.sites]

    @staticmethod
    def _combine_pwms(pwms, weights, alphabet):
        """pssm_model.py:159-169."""
        length = pwms[0].length
        assert all(length == pwm.length for pwm in pwms)
        pwm_vals = {}
        for let in alphabet:
            cols = [(dict.__getitem__(pwm, let) if isinstance(pwm, dict) else pwm[let], w) for pwm, w in zip(pwms, weights)]
            pwm_vals[let] = [psum(col[i]*w for col, w in cols) for i in range(length)]
        return PositionWeightMatrix(alphabet, pwm_vals)

    def _calculate(self, pssm, sequence):
        """pssm_model.py:88-101: host restatement of the ambiguity rule, kept for
        API compatibility.  The scoring path applies the same rule on the device."""
        score = 0.0
        for pos in range(len(sequence)):
            try:
                score += pssm[sequence[pos]][pos]
            except KeyError:
                score += psum(pssm[l][pos] for l in 'ACGT') / 4
        return score

    # ---------------------------------------------------------- device state
    def _dev(self):
        if self._device is None:
            if not torch.cuda.is_available():
                raise _cabi.CgbkError("no CUDA device: cgb_b200 has no CPU fallback")
            self._device = torch.device("cuda", torch.cuda.current_device())
        return torch.device(self._device)

    def _model(self):
        if self._handle is None:
            if sorted(self.alphabet) != ["A", "C", "G", "T"]:
                raise ValueError("PSSM has wrong alphabet: %s - Use only with DNA motifs" % self.alphabet)
            rows = self.pssm.rows()
            m = len(rows)
            flat = (C.c_double * (4 * m))(*[x for row in rows for x in row])
            h = C.c_void_p()
            dev = self._dev()
            _cabi.check(_cabi.lib().cgbk_model_create(flat, m, dev.index or 0, C.byref(h)))
            self._handle = h
        return self._handle

    def __del__(self):
        try:
            if getattr(self, "_batch_owner", None) is not None:
                self._handle = None             # borrowed from a ModelBatch, which frees the tables
            elif getattr(self, "_handle", None) is not None and self._handle.value:
                _cabi.lib().cgbk_model_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def _stream(self):
        return torch.cuda.current_stream(self._dev()).cuda_stream

    # ---------------------------------------------------------------- scoring
    def score_seqs(self, seqs, both=True):
        """``[score_seq(s, both) for s in seqs]`` with one pack + one kernel launch."""
        m = self.length
        seqs = [str(s) if not isinstance(s, (bytes, bytearray)) else s for s in seqs]
        lens = np.array([len(s) for s in seqs], dtype=np.int64)
        if len(seqs) == 0:
            return []
        if (lens < m - 1).any():
            # Biopython allocates numpy.empty(len(seq) - m + 1): a negative count raises ...
            raise ValueError("negative dimensions are not allowed")
        if (lens == m - 1).any():
            # ... and a count of zero is an empty score array (no window): [] from score_seq
            full = [k for k in range(len(seqs)) if lens[k] >= m]
            scored = self.score_seqs([seqs[k] for k in full], both) if full else []
            out = [[] if both else np.empty(0, dtype=np.float32) for _ in seqs]
            for k, v in zip(full, scored):
                out[k] = v
            return out
        model = self._model()
        dev = self._dev()
        g = PackedGenome.from_sequences(seqs, dev)
        nwin = lens - m + 1
        out_off = np.concatenate([[0], np.cumsum(nwin)]).astype(np.int64)
        total = int(out_off[-1])
        d_start = torch.from_numpy(g.seq_off.copy()).to(dev)
        d_off = torch.from_numpy(out_off).to(dev)
        f32 = torch.empty(total, dtype=torch.float32, device=dev)
        r32 = torch.empty(total, dtype=torch.float32, device=dev)
        b64 = torch.empty(total, dtype=torch.float64, device=dev)
        f64 = torch.zeros(len(seqs), dtype=torch.float64, device=dev)
        r64 = torch.zeros(len(seqs), dtype=torch.float64, device=dev)
        _cabi.check(_cabi.lib().cgbk_score_windows(
            model, g.struct_ref(), d_start.data_ptr(), d_off.data_ptr(), len(seqs), total,
            _cabi.RC_MATRIX_ORDER | _cabi.SINGLE_WINDOW_F64,
            f32.data_ptr(), r32.data_ptr(), b64.data_ptr(), f64.data_ptr(), r64.data_ptr(), self._stream()))
        if both:
            host = b64.cpu().numpy()
            return [host[out_off[k]:out_off[k + 1]].tolist() for k in range(len(seqs))]
        host = f32.cpu().numpy()
        single = f64.cpu().numpy()
        out = []
        for k in range(len(seqs)):
            if nwin[k] == 1:
                # pssm_model.py:117-119: a bare scalar wrapped in a list; a repaired
                # (ambiguous) score stays a Python float, an ordinary one is float32
                v = float(single[k])
                out.append([np.float32(v) if float(np.float32(v)) == v else v])
            else:
                out.append(host[out_off[k]:out_off[k + 1]].copy())
        return out

    def score_seq(self, seq, both=True):
        """pssm_model.py:103-133: scores of all positions; both strands combined
        with the soft-max ``log2(2**s + 2**rc)`` unless ``both`` is False."""
        return self.score_seqs([seq], both)[0]

    def score_self(self):
        """pssm_model.py:84-86.  The sites are as long as the motif, one window each: scored by
        the library on the host in the reference's arithmetic (``cgbk_score_sites_host``; a few
        thousand table look-ups are not worth a launch).  Ragged site lists take the device path."""
        m = self.length
        colls = self._collection_set
        # the collections' own byte strings serve when every collection is one of ours with sites of
        # the motif's length; a materialised -- possibly replaced -- site list wins over them, and
        # any other collection type (PSSMModel.from_pwm, a caller's class) goes through its .sites
        fast = "sites" not in self.__dict__ and all(
            getattr(c, "_uniform", False) and getattr(c, "_raw", None) is not None and c.length == m for c in colls)
        n_sites = sum(c.site_count for c in colls) if fast else len(self.sites)
        if n_sites == 0:
            return []
        if not fast:
            sites = self.sites
            if any(len(s) != m for s in sites):
                return self.score_seqs(sites, True)
            raw = b"".join(s if isinstance(s, (bytes, bytearray)) else str(s).encode("ascii") for s in sites)
        else:
            raw = b"".join(c._raw for c in colls)           # the collections keep their sites as bytes
        flat = np.array(self.pssm.rows(), dtype=np.float64).ravel()
        out = np.empty(n_sites, dtype=np.float64)
        _cabi.check(_cabi.lib().cgbk_score_sites_host(flat.ctypes.data_as(C.POINTER(C.c_double)), m, raw, n_sites, None, None,
                                                      out.ctypes.data))
        return [[v] for v in out.tolist()]

    # ------------------------------------------------------------ tiled scans
    def scan_hits(self, genome, threshold=None, revseq_order=True, hit_cap=None):
        """All windows of every sequence of ``genome`` whose PSSM score on either
        strand is >= threshold: (seq_index, pos, strand, score) arrays sorted by
        (seq, pos, +strand first).  ``revseq_order`` selects the accumulation order
        of the minus strand: True = identify_sites (genome.py:440-445), False =
        rev_comp_pssm (pssm_model.py:115)."""
        pending = self._launch_scan(genome, threshold, revseq_order, hit_cap)
        return self._collect_scan(genome, pending)

    def _launch_scan(self, genome, threshold=None, revseq_order=True, hit_cap=None):
        """Enqueue one hit scan (no host synchronisation); see ``_collect_scan``."""
        if threshold is None:
            threshold = self.threshold()
        threshold = float(threshold)
        model = self._model()
        m = self.length
        dev = self._dev()
        if hit_cap is None:
            # expected hits at the Patser operating point, with head-room
            nwin = genome.total_windows(m)
            hit_cap = int(2 * nwin * 2.0 ** (-max(self.IC, 0.0)) * 4) + 4096
        # 16 bytes per hit + 4 x 8 bytes per candidate slot: never ask for more than half of what
        # the device has free (a low-information motif on a Gbp sequence would otherwise request
        # hundreds of GB up front; the two-phase retry in _collect_scan still grows the buffers
        # to the real count, and fails with a clear message if that cannot fit)
        if hit_cap * 48 > (1 << 30):                   # the query costs ~0.1 ms: only for big requests
            free_bytes, _ = torch.cuda.mem_get_info(dev)
            max_cap = max(4096, int(free_bytes // 2 // 48))
            if hit_cap > max_cap:
                if getattr(self, "_hit_cap_needed", 0) > max_cap:
                    raise MemoryError("the hit list needs %d entries, the device has room for %d"
                                      % (self._hit_cap_needed, max_cap))
                hit_cap = max_cap
        flags = _cabi.RC_REVSEQ_ORDER if revseq_order else _cabi.RC_MATRIX_ORDER
        counters = torch.empty(_cabi.CTR_COUNT, dtype=torch.int64, device=dev)
        cand_cap = 4 * hit_cap
        work = genome.workspace(m, cand_cap)
        hits = torch.empty(hit_cap * 2, dtype=torch.int64, device=dev)   # 16-byte records
        _cabi.check(_cabi.lib().cgbk_scan_hits(
            model, genome.struct_ref(), genome.seqset, threshold, flags, hits.data_ptr(), hit_cap,
            counters.data_ptr(), work.data_ptr(), work.numel(), cand_cap, self._stream()))
        return {"hits": hits, "counters": counters, "hit_cap": hit_cap, "cand_cap": cand_cap,
                "threshold": threshold, "revseq_order": revseq_order}

    def _collect_scan(self, genome, pending):
        """Wait for a launched scan, re-run it with larger buffers if a capacity was exceeded
        (two-phase count -> fill), and return the sorted hit arrays."""
        while True:
            c = pending["counters"].cpu().numpy()
            n_cand, n_hits = int(c[_cabi.CTR_CANDIDATES]), int(c[_cabi.CTR_HITS])
            if n_cand <= pending["cand_cap"] and n_hits <= pending["hit_cap"]:
                break
            self._hit_cap_needed = max(n_hits, (n_cand + 3) // 4) + 1024
            pending = self._launch_scan(genome, pending["threshold"], pending["revseq_order"],
                                        self._hit_cap_needed)
            self._hit_cap_needed = 0
        self.last_scan_counters = {"candidates": n_cand, "hits": n_hits,
                                   "dirty_tiles": int(c[_cabi.CTR_DIRTY_TILES])}
        # Sort on the device (by packed offset, + strand first: the (seq, pos, strand) order, as
        # sequence offsets grow with the sequence index) and bring the records over once; low-
        # information motifs produce millions of hits and a host lexsort would dominate.
        dev_rec = pending["hits"][:2 * n_hits].view(n_hits, 2)
        if n_hits > 1:
            minus = (dev_rec[:, 1] & 0xffffffff) != 1            # low word of the second int64 = strand
            order = torch.argsort(dev_rec[:, 0] * 2 + minus.to(torch.int64))
            dev_rec = dev_rec[order]
        rec = dev_rec.cpu().numpy().reshape(-1).view(
            np.dtype([("gpos", "<i8"), ("strand", "<i4"), ("score", "<f4")]))
        gpos = rec["gpos"]
        seq = np.searchsorted(genome.seq_off, gpos, side="right") - 1
        pos = gpos - genome.seq_off[seq]
        return seq, pos, rec["strand"].copy(), rec["score"].copy()

    def region_descriptors(self, genome, regions, drop_last=True):
        """(window start offsets, window counts) device tensors of a region table.

        ``regions``: iterable of ``(seq_index, starts, ends)``: slices ``sequence[start:end]`` of
        replicon ``seq_index`` (Python slice semantics, chromid.py:94).  Every m-mer start of a
        slice is a window; ``drop_last`` leaves out the last one, as ``xrange(len(seq) - m)`` does
        in Genome.build_PSSM_model (genome.py:217-218).  Slices too short for a window add none."""
        m = self.length
        all_s, all_n = [], []
        for seq_index, starts, ends in regions:
            L = genome.length(seq_index)
            s, e = python_slice_bounds(np.asarray(starts, dtype=np.int64), np.asarray(ends, dtype=np.int64), L)
            nwin = e - s - m + (0 if drop_last else 1)
            keep = nwin > 0
            all_s.append(s[keep] + int(genome.seq_off[seq_index]))
            all_n.append(nwin[keep])
        s = np.concatenate(all_s) if all_s else np.zeros(0, dtype=np.int64)
        n = np.concatenate(all_n).astype(np.int32) if all_n else np.zeros(0, dtype=np.int32)
        dev = self._dev()
        return torch.from_numpy(np.ascontiguousarray(s)).to(dev), torch.from_numpy(np.ascontiguousarray(n)).to(dev)

    def background_stats(self, genome, n_bins=256, hist_lo=None, hist_hi=None, accumulate_into=None,
                         out=None, keep_scores=False, regions=None, drop_last=True):
        """Soft-max score distribution over ALL windows of ``genome`` (both strands
        combined as in score_seq), or -- with ``regions`` (see :meth:`region_descriptors`) -- over
        the m-mers of those regions only, each scored as a sequence of its own: the reference's
        background population when the regions are the genes' upstream non-coding regions
        (genome.py:215-218).  Returns a dict with device tensors ``stats``
        (float64[8]: n, sum, sumsq, mean, std) and ``hist`` (int64[n_bins]) plus the
        bin range.  ``accumulate_into``: a previous result to add this genome to;
        ``out``: a previous result whose buffers are overwritten (no allocation).
        ``keep_scores``: also keep every window's soft-max score on the device (the
        "score plane", ~8 bytes per base) so that ``binding_probabilities`` on the
        same genome does not re-score the promoters.
        No host synchronisation happens here."""
        model = self._model()
        m = self.length
        dev = self._dev()
        if hist_lo is None:
            hist_lo = self._score_bounds[0]
        if hist_hi is None:
            hist_hi = self._score_bounds[1]
        if regions is not None:
            lib = _cabi.lib()
            d_start, d_nwin = regions if (isinstance(regions, tuple) and len(regions) == 2 and
                                          isinstance(regions[0], torch.Tensor)) \
                else self.region_descriptors(genome, regions, drop_last)
            if accumulate_into is not None:
                stats, hist, acc = accumulate_into["stats"], accumulate_into["hist"], 1
                n_bins = int(hist.numel())
            else:
                rec = torch.empty(_cabi.BG_COUNT + n_bins, dtype=torch.int64, device=dev)
                stats, hist, acc = rec[:_cabi.BG_COUNT].view(torch.float64), rec[_cabi.BG_COUNT:], 0
            work = torch.empty(int(lib.cgbk_region_stats_workspace_bytes()), dtype=torch.uint8, device=dev)
            _cabi.check(lib.cgbk_background_stats_regions(
                model, genome.struct_ref(), d_start.data_ptr(), d_nwin.data_ptr(), int(d_start.numel()),
                float(hist_lo), float(hist_hi), int(n_bins), hist.data_ptr() if n_bins else None,
                stats.data_ptr(), acc, work.data_ptr(), work.numel(), self._stream()))
            work.record_stream(torch.cuda.current_stream(dev))
            return {"stats": stats, "hist": hist, "hist_lo": float(hist_lo), "hist_hi": float(hist_hi),
                    "n_bins": int(n_bins), "plane": None, "genome": None, "work": None, "work_key": None}
        plane = None
        work = None
        if out is not None:
            stats, hist, acc = out["stats"], out["hist"], 0
            n_bins = int(hist.numel())
            plane = out.get("plane") if out.get("genome") is genome else None
            if out.get("work_key") == (id(genome), m):
                work = out["work"]
        elif accumulate_into is None:
            # one allocation: stats record | histogram | scan workspace (whose first words are
            # the scan's counters) -- the library clears all three with a single memset
            wbytes = int(_cabi.lib().cgbk_scan_workspace_bytes(genome.seqset, m, 0))
            rec = torch.empty(_cabi.BG_COUNT + n_bins + (wbytes + 7) // 8, dtype=torch.int64, device=dev)
            stats = rec[:_cabi.BG_COUNT].view(torch.float64)
            hist = rec[_cabi.BG_COUNT:_cabi.BG_COUNT + n_bins]
            work = rec[_cabi.BG_COUNT + n_bins:].view(torch.uint8)
            acc = 0
        else:
            stats, hist = accumulate_into["stats"], accumulate_into["hist"]
            acc = 1
        if keep_scores and plane is None:
            n_ints = int(_cabi.lib().cgbk_score_plane_ints(genome.seqset, m))
            plane = torch.empty(max(n_ints, 1), dtype=torch.int32, device=dev)
        if not keep_scores:
            plane = None
        own_work = work is not None
        if work is None:
            work = genome.workspace(m, 0)
        _cabi.check(_cabi.lib().cgbk_background_stats(
            model, genome.struct_ref(), genome.seqset, float(hist_lo), float(hist_hi), int(n_bins),
            hist.data_ptr(), stats.data_ptr(), acc,
            plane.data_ptr() if plane is not None else None, int(plane.numel()) if plane is not None else 0,
            work.data_ptr(), work.numel(), self._stream()))
        return {"stats": stats, "hist": hist, "hist_lo": float(hist_lo), "hist_hi": float(hist_hi),
                "n_bins": int(n_bins), "plane": plane, "genome": genome if plane is not None else None,
                "work": work if own_work else None, "work_key": (id(genome), m) if own_work else None}

    def fit_background(self, genome=None, bg=None, sync=True, keep_scores=True, regions=None):
        """build_bayesian_estimator (binding_model.py:71-92) with the background
        moments taken from the genome-wide scan, from the m-mers of ``regions`` (the
        reference's upstream non-coding population, genome.py:215-218), or from a finished
        ``background_stats`` result, e.g. after a cross-rank all-reduce.  The score
        plane of the genome-wide scan (``keep_scores``) is remembered and reused by
        ``binding_probabilities`` on the same genome."""
        if bg is None and regions is not None:
            bg = self.background_stats(genome, regions=regions)
        if bg is None:
            bg = self.background_stats(genome, keep_scores=keep_scores)
        self._bg_stats_dev = bg["stats"]
        self._last_bg = bg
        self._dev_ms_bg = bg["stats"][_cabi.BG_MEAN:_cabi.BG_STD + 1]
        if getattr(self, "_mu_m", None) is None:
            pssm_scores = self.score_self()
            self._mu_m, self._std_m = np.mean(pssm_scores), np.std(pssm_scores)
            self._dev_ms_m = None
        if self._dev_ms_m is None:
            self._dev_ms_m = torch.tensor([float(self._mu_m), float(self._std_m)], dtype=torch.float64,
                                          device=self._dev())
        if sync:
            host = bg["stats"].cpu().numpy()
            self._mu_bg, self._std_bg = float(host[_cabi.BG_MEAN]), float(host[_cabi.BG_STD])
        return bg

    def _estimator_changed(self):
        dev = self._dev()
        self._dev_ms_m = torch.tensor([float(self._mu_m), float(self._std_m)], dtype=torch.float64, device=dev)
        self._dev_ms_bg = torch.tensor([float(self._mu_bg), float(self._std_bg)], dtype=torch.float64, device=dev)

    def _require_estimator(self):
        if (self._dev_ms_m is None or self._dev_ms_bg is None) and \
                getattr(self, "_mu_m", None) is not None and getattr(self, "_mu_bg", None) is not None:
            self._estimator_changed()
        if self._dev_ms_m is None or self._dev_ms_bg is None:
            # the reference fails with AttributeError on the unset _mu_m
            raise AttributeError("'PSSMModel' object has no attribute '_mu_m' "
                                 "(call build_bayesian_estimator or fit_background first)")

    def binding_probabilities(self, genome, starts, ends, p_motif, alpha=1/350.0, seq_index=0,
                              out=None, region_tensors=None, use_plane=True):
        """Posterior of regulation for promoter regions ``sequence[start:end]``
        (Python slice semantics, gene.py:127) of one replicon: device float64
        tensor, one value per region (binding_model.py:94-113 for each).  When the
        last ``fit_background`` kept the score plane of this genome (``use_plane``),
        window scores are read from it; otherwise every window is re-scored."""
        self._require_estimator()
        model = self._model()
        m = self.length
        dev = self._dev()
        if region_tensors is None:
            region_tensors = self.promoter_descriptors(genome, starts, ends, seq_index)
        d_start, d_nwin = region_tensors
        n = int(d_start.numel())
        if out is None:
            out = torch.empty(n, dtype=torch.float64, device=dev)
        plane = None
        if use_plane:
            bg = getattr(self, "_last_bg", None)
            if bg is not None and bg.get("genome") is genome and bg.get("plane") is not None:
                plane = bg["plane"]
        def launch(pl):
            return _cabi.lib().cgbk_posteriors(
                model, genome.struct_ref(), genome.seqset, pl.data_ptr() if pl is not None else None,
                d_start.data_ptr(), d_nwin.data_ptr(), n,
                self._dev_ms_m.data_ptr(), self._dev_ms_bg.data_ptr(), float(p_motif), float(alpha),
                out.data_ptr(), self._stream())
        rc = launch(plane)
        if rc == _cabi.ERR_INVALID and plane is not None:
            # another model re-tiled the shared sequence set since the plane was written (the
            # plane's addressing is gone with the tiling): forget it and score the windows again
            self._last_bg["plane"] = None
            self._last_bg["genome"] = None
            rc = launch(None)
        _cabi.check(rc)
        return out

    def promoter_descriptors(self, genome, starts, ends, seq_index=0):
        """(start offset, window count) device tensors for cgbk_posteriors."""
        m = self.length
        L = genome.length(seq_index)
        starts = np.asarray(starts, dtype=np.int64)
        ends = np.asarray(ends, dtype=np.int64)
        if starts.min(initial=0) >= 0 and ends.min(initial=0) >= 0:
            s = np.minimum(starts, L)                   # common case: no negative slice index
            e = np.maximum(np.minimum(ends, L), s)
        else:
            s, e = python_slice_bounds(starts, ends, L)
        nwin = e - s
        nwin -= m - 1
        if nwin.min(initial=0) < 0:
            raise ValueError("negative dimensions are not allowed")   # Biopython, len(seq) < m - 1
        dev = self._dev()
        # one pinned staging buffer [n x int64 starts | n x int32 counts], one async H2D copy
        n = len(s)
        nbytes = 12 * n
        stage = getattr(self, "_desc_stage", None)
        if stage is None or stage.numel() < nbytes:
            stage = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8).pin_memory()
            self._desc_stage = stage
        ev = getattr(self, "_desc_event", None)
        if ev is not None:
            ev.synchronize()                       # the previous copy out of the staging buffer is done
        host = stage.numpy()
        host[:8 * n].view(np.int64)[:] = s + int(genome.seq_off[seq_index])
        host[8 * n:12 * n].view(np.int32)[:] = nwin
        d = stage[:nbytes].to(dev, non_blocking=True)
        self._desc_event = torch.cuda.Event()
        self._desc_event.record(torch.cuda.current_stream(dev))
        return d[:8 * n].view(torch.int64), d[8 * n:12 * n].view(torch.int32)

    def _posterior_of_sequences(self, seqs, p_motif, alpha):
        self._require_estimator()
        m = self.length
        seqs = [str(s) if not isinstance(s, (bytes, bytearray)) else s for s in seqs]
        lens = np.array([len(s) for s in seqs], dtype=np.int64)
        if (lens < m - 1).any():
            raise ValueError("negative dimensions are not allowed")
        dev = self._dev()
        g = PackedGenome.from_sequences(seqs, dev)
        d_start = torch.from_numpy(g.seq_off.copy()).to(dev)
        d_nwin = torch.from_numpy((lens - m + 1).astype(np.int32)).to(dev)
        out = self.binding_probabilities(g, None, None, p_motif, alpha, region_tensors=(d_start, d_nwin))
        return out.cpu().numpy()

    # ------------------------------------------------ one call per replicon, host in / host out
    def regulation_pipeline(self, sequence, starts, ends, p_motif, alpha=1/350.0, n_bins=256,
                            hist_lo=None, hist_hi=None):
        """Background fit over ALL windows of ``sequence`` (genome.py:215-226) followed by the
        posterior of regulation of every promoter slice ``sequence[start:end]``
        (genome.py:330-333, gene.py:129-146) in ONE library call with host buffers
        (``cgbk_regulation_pipeline``).

        ``sequence``: str / bytes / uint8 numpy array / uint8 CPU tensor (a pinned tensor is
        copied asynchronously), or a :class:`HostPackedGenome` (planes packed on the host or read
        from a packed genome file: ``cgbk_regulation_pipeline_packed``, a quarter to a half of the
        bytes cross PCIe and no pack kernel runs).  Returns ``(posteriors, stats, hist)`` as numpy arrays
        (``stats``: n, sum, sumsq, mean, std of the background) and sets ``_mu_bg`` /
        ``_std_bg`` like ``build_bayesian_estimator``."""
        packed = sequence if isinstance(sequence, HostPackedGenome) else None
        if packed is not None:
            seq_t = packed.planes
        elif isinstance(sequence, torch.Tensor):
            # the library reads these bytes from HOST memory: a device tensor, another dtype or a
            # strided view would hand it a wrong pointer or length
            if sequence.device.type != "cpu":
                raise ValueError("sequence must be a CPU tensor (the pipeline uploads it itself)")
            if sequence.dtype != torch.uint8:
                raise ValueError("sequence must be a uint8 tensor of ASCII bytes, got %s" % sequence.dtype)
            seq_t = sequence if sequence.is_contiguous() else sequence.contiguous()
            if seq_t.dim() != 1:
                seq_t = seq_t.reshape(-1)
        elif isinstance(sequence, np.ndarray):
            seq_t = torch.from_numpy(np.ascontiguousarray(sequence, dtype=np.uint8))
        else:
            raw = sequence if isinstance(sequence, (bytes, bytearray)) else str(sequence).encode("ascii")
            seq_t = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
        n_bases = packed.n_bases if packed is not None else int(seq_t.numel())
        if getattr(self, "_mu_m", None) is None:
            pssm_scores = self.score_self()
            self._mu_m, self._std_m = np.mean(pssm_scores), np.std(pssm_scores)
        # the same int64 arrays passed again (a loop over genomes sharing one gene layout, the
        # benchmark): their addresses are remembered with the plan.  Only arrays that needed no
        # conversion qualify, so in-place edits by the caller are still seen.
        plan = getattr(self, "_pipe_plan", None)
        reuse_ptrs = plan is not None and plan.get("starts") is starts and plan.get("ends") is ends
        if not reuse_ptrs:
            starts = np.ascontiguousarray(starts, dtype=np.int64)
            ends = np.ascontiguousarray(ends, dtype=np.int64)
        n = len(starts)
        if hist_lo is None:
            hist_lo = self._score_bounds[0]
        if hist_hi is None:
            hist_hi = self._score_bounds[1]
        # buffers, pointers and result views for this problem size are set up once and reused:
        # the per-call host work is argument marshalling + three small copies
        key = (n_bases, n, int(n_bins))
        if plan is None or plan["key"] != key:
            reuse_ptrs = False
            lib = _cabi.lib()
            dev = self._dev()
            shared = _pipeline_buffers(dev)
            # device scratch and pinned result buffers are shared by all models of this thread on
            # the device (a pipeline call returns synchronised, so nothing of the previous call is
            # in flight): a run over many (genome, PSWM) pairs allocates them once, not per model.
            # A plan keeps the buffers it was made with alive, so a later, larger request of another
            # model (which replaces the shared ones) leaves this plan valid.
            need = int(lib.cgbk_pipeline_scratch_bytes(n_bases, n, self.length, n_bins))
            scratch = shared.get("scratch")
            if scratch is None or scratch.numel() < need:
                scratch = torch.empty(need + need // 8, dtype=torch.uint8, device=dev)
                shared["scratch"] = scratch
            stage_bytes = int(lib.cgbk_pipeline_staging_bytes(n))
            host_need = stage_bytes + 8 * n + 64 + 8 * n_bins
            host = shared.get("host")
            if host is None or host.numel() < host_need:
                host = torch.empty(host_need + host_need // 8, dtype=torch.uint8).pin_memory()
                shared["host"] = host
            arr = host.numpy()
            hp = host.data_ptr()
            args = _cabi.PipelineArgs()
            args.model = self._model().value
            args.n_bases, args.n_regions, args.n_bins = n_bases, n, int(n_bins)
            args.d_scratch, args.scratch_bytes = scratch.data_ptr(), scratch.numel()
            args.h_staging, args.staging_bytes = hp, stage_bytes
            args.h_post, args.h_stats = hp + stage_bytes, hp + stage_bytes + 8 * n
            args.h_hist = (hp + stage_bytes + 8 * n + 64) if n_bins > 0 else None
            plan = {"key": key, "fn": lib.cgbk_regulation_pipeline_args, "args": args, "ref": C.addressof(args),
                    "scratch": scratch, "host": host,
                    "post": arr[stage_bytes:stage_bytes + 8 * n].view(np.float64),
                    "stats": arr[stage_bytes + 8 * n:stage_bytes + 8 * n + 64].view(np.float64),
                    "hist": arr[stage_bytes + 8 * n + 64:stage_bytes + 8 * n + 64 + 8 * n_bins].view(np.uint64)}
            self._pipe_plan = plan
        # one struct, rewritten in place: the call itself passes a single pointer
        args = plan["args"]
        if packed is not None:
            args.h_ascii = None
            args.h_bases2 = packed.bases2.data_ptr()
            args.h_amb = packed.amb.data_ptr() if packed.n_amb else None
            args.h_lower = packed.lower.data_ptr() if packed.n_lower else None
        else:
            args.h_ascii = seq_t.data_ptr()
        if not reuse_ptrs:
            args.h_start, args.h_end = starts.ctypes.data, ends.ctypes.data
            plan["starts"], plan["ends"] = starts, ends                 # keep them alive
        args.mu_m, args.std_m = float(self._mu_m), float(self._std_m)
        args.p_motif, args.alpha = float(p_motif), float(alpha)
        args.hist_lo, args.hist_hi = float(hist_lo), float(hist_hi)
        args.stream = self._stream()
        _cabi.check(plan["fn"](plan["ref"]))
        post, stats, hist = plan["post"].copy(), plan["stats"].copy(), plan["hist"].copy()
        self._mu_bg, self._std_bg = float(stats[_cabi.BG_MEAN]), float(stats[_cabi.BG_STD])
        self._dev_ms_bg = None                         # re-uploaded lazily by _require_estimator
        self._last_bg = None
        return post, stats, hist


_PIPE_SHARED = threading.local()


def _pipeline_buffers(dev):
    """Per (host thread, device): the regulation pipeline's device scratch and pinned host block."""
    d = getattr(_PIPE_SHARED, "by_device", None)
    if d is None:
        d = _PIPE_SHARED.by_device = {}
    return d.setdefault(dev.index or 0, {})


def prepare_thresholds(models, n_threads=0):
    """IC and Patser threshold (pssm_model.py:57-70) of many models in one library call, spread
    over host threads (``cgbk_patser_threshold_batch``); the values land in the models' cached
    ``IC`` / ``patser_threshold`` properties, identical to what each would compute alone.
    Worth it for JASPAR-scale motif sets: the score-distribution DP is ~3 ms per motif."""
    todo = [M for M in models if "patser_threshold" not in M.__dict__]
    if not todo:
        return
    rows = [M.pssm.rows() for M in todo]
    flat = [x for r in rows for row in r for x in row]
    n = len(todo)
    arr = (C.c_double * len(flat))(*flat)
    ms = (C.c_int * n)(*[len(r) for r in rows])
    thr, ic = (C.c_double * n)(), (C.c_double * n)()
    _cabi.check(_cabi.lib().cgbk_patser_threshold_batch(arr, ms, n, 10 ** 3, int(n_threads), thr, ic,
                                                        None, None, None))
    for M, t in zip(todo, thr):
        M.__dict__["patser_threshold"] = float(t)
    del ic      # IC stays with PSSM.mean(): its float64 sum is cheap and already exact


def scan_hits_many(models, genome, thresholds=None, revseq_order=True):
    """Hit scans of many motifs over one packed genome (BASELINE configs[4]: JASPAR-scale motif
    sets): all scans are enqueued back to back on the current stream and collected afterwards,
    so the GPU never waits for the host between motifs.  Returns one
    ``(seq_index, pos, strand, score)`` tuple per model, as ``PSSMModel.scan_hits``."""
    if thresholds is None:
        prepare_thresholds(models)
        thresholds = [None] * len(models)
    # launched scans hold their hit buffers until collected: keep at most a quarter of the
    # device's memory in flight (a Gbp sequence x low-information motifs is GBs per scan)
    budget = torch.cuda.get_device_properties(genome.device).total_memory // 4
    results = [None] * len(models)
    in_flight, held = [], 0
    for k, (M, thr) in enumerate(zip(models, thresholds)):
        p = M._launch_scan(genome, thr, revseq_order)
        in_flight.append((k, M, p))
        held += p["hits"].numel() * 8
        while held > budget and len(in_flight) > 1:
            j, Mj, pj = in_flight.pop(0)
            held -= pj["hits"].numel() * 8
            results[j] = Mj._collect_scan(genome, pj)
    for j, Mj, pj in in_flight:
        results[j] = Mj._collect_scan(genome, pj)
    return results
