"""Python port of the reference's CPU hot path, used ONLY as the timed CPU baseline
(bench.py ``cpu_baseline`` / ``--impl reference``) — TEST/BENCH infrastructure.

The reference tree cannot travel to the GPU box, so the two files that carry the
cost are restated with the same per-window Python work as the originals:

  score_seq            cgb/pssm_model.py:103-133: two C scans (Biopython _pwm.c,
                       restated in oracle/pwm_oracle.c), then a Python loop over EVERY
                       window for the NaN check and a Python list comprehension for
                       the soft-max ``log2(2**s + 2**rc)``
  binding_probability  cgb/binding_model.py:94-113: score_seq + two frozen scipy
                       normal distributions built per call + numpy reductions
  identify_sites loop  cgb/genome.py:433-445

Checked against the verbatim reference in tests (container) and against
np_oracle everywhere.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import scipy.stats

from . import c_oracle
from .np_oracle import _calculate_ambiguous, reverse_complement


def log2(x):
    return math.log(x, 2)            # cgb/misc.py:14-16


class PortModel:
    """Holds the [m][4] tables of an OracleModel and mirrors PSSMModel's methods."""

    def __init__(self, oracle_model):
        self.om = oracle_model
        self.length = oracle_model.length
        self.fwd = np.ascontiguousarray(oracle_model.fwd)
        self.rc = np.ascontiguousarray(oracle_model.rc)
        self._mu_m = oracle_model._mu_m
        self._std_m = oracle_model._std_m
        self._mu_bg = oracle_model._mu_bg
        self._std_bg = oracle_model._std_bg

    def _c_calculate(self, table, seq):
        n = len(seq) - self.length + 1
        if n < 0:
            raise ValueError("negative dimensions are not allowed")
        if n == 0:
            return np.empty(0, np.float32)
        out = np.empty(n, np.float32)
        c_oracle.lib().cgbo_calculate(seq, len(seq), table.ctypes.data_as(C.POINTER(C.c_double)),
                                      self.length, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out[0] if n == 1 else out

    def score_seq(self, seq, both=True):
        if isinstance(seq, str):
            seq = seq.encode("ascii")
        scores = self._c_calculate(self.fwd, seq)
        rc_scores = self._c_calculate(self.rc, seq)
        if self.length == len(seq):
            scores, rc_scores = [scores], [rc_scores]
        for i in range(len(scores)):
            if math.isnan(scores[i]):
                site = seq[i:i + self.length]
                scores[i] = _calculate_ambiguous(self.fwd, site)
                rc_scores[i] = _calculate_ambiguous(self.rc, site)
        if both:
            # pinned numpy 1.x: np.float32 ** in a Python-int context promotes to float64
            scores = [log2(2 ** float(score) + 2 ** float(rc_score))
                      for score, rc_score in zip(scores, rc_scores)]
        return scores

    def binding_probability(self, seq, p_motif, alpha=1 / 350.0):
        pssm_scores = self.score_seq(seq)
        pdf_m = scipy.stats.distributions.norm(self._mu_m, self._std_m).pdf
        pdf_bg = scipy.stats.distributions.norm(self._mu_bg, self._std_bg).pdf
        lh_m = (alpha * pdf_m(pssm_scores) + (1 - alpha) * pdf_bg(pssm_scores))
        lh_bg = pdf_bg(pssm_scores)
        lh_ratio = np.exp(np.sum(np.log(lh_bg) - np.log(lh_m)))
        return 1 / (1 + lh_ratio * (1 - p_motif) / p_motif)

    def region_hits(self, seq, start, threshold):
        """genome.py:433-445 for one region."""
        if isinstance(seq, str):
            seq = seq.encode("ascii")
        hits = []
        scores = self.score_seq(seq, both=False)
        for i, score in enumerate(scores, start=start):
            if score >= threshold:
                hits.append((i, 1, score))
        rc_scores = self.score_seq(reverse_complement(seq), both=False)
        for i, score in enumerate(reversed(rc_scores), start=start):
            if score >= threshold:
                hits.append((i, -1, score))
        return hits


def config2_step(port, seq, promoters, p_motif, alpha, bg_slice=None):
    """The CPU work of one BASELINE configs[1] step: soft-max scores of every window
    of ``seq[bg_slice]`` (background sample), mean/std, then the posterior of every
    promoter ``(start, end)``.  Returns (n_windows, n_promoters, mu_bg, std_bg, posts)."""
    s = seq if bg_slice is None else seq[bg_slice[0]:bg_slice[1]]
    bg = port.score_seq(s)
    port._mu_bg, port._std_bg = np.mean(bg), np.std(bg)
    posts = [port.binding_probability(seq[a:b], p_motif, alpha) for a, b in promoters]
    return len(bg), len(posts), port._mu_bg, port._std_bg, posts
