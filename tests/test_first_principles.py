"""Checks that stand OUTSIDE the two restatements of Biopython's motif layer.

``cgb_b200/motif_matrix.py`` (product) and ``oracle/biopython_shim.py`` (oracle) both restate
Bio.motifs 1.76 from its published algorithm, so a shared misreading would pass every parity
test.  Here nothing of either is reused: the matrices are recomputed with numpy from the raw
site strings by the textbook definitions (site_collection.py:15-21, pssm_model.py:47-70), the
information content by its closed form, and for short motifs ALL 4^m windows are enumerated
to obtain the exact score distribution under the uniform background, against which the Patser
threshold of the score-grid DP is bracketed.
"""
import itertools

import numpy as np
import pytest

from cgb_b200 import PSSMModel, SiteCollection
from oracle import biopython_shim as shim


def _random_sites(rng, m, n_sites):
    cols = rng.dirichlet([0.5] * 4, size=m)
    return ["".join("ACGT"[rng.choice(4, p=cols[j])] for j in range(m)) for _ in range(n_sites)]


def _textbook_log_odds(sites):
    """[m][4] log2((count + 1) / (N + 4) / 0.25), straight from the definition."""
    m, n = len(sites[0]), len(sites)
    counts = np.zeros((m, 4))
    for s in sites:
        for j, ch in enumerate(s):
            counts[j, "ACGT".index(ch)] += 1
    p = (counts + 1.0) / (n + 4.0)
    return p, np.log2(p / 0.25)


def _all_window_scores(lo):
    """Score of every one of the 4^m windows (uniform background => each has weight 4^-m)."""
    m = lo.shape[0]
    total = np.zeros(1)
    for j in range(m):
        total = (total[:, None] + lo[j][None, :]).reshape(-1)
    return total


CASES = [(4, 6, 1), (5, 12, 2), (6, 20, 3), (7, 9, 4), (8, 20, 5), (8, 40, 6), (6, 3, 7)]


@pytest.mark.parametrize("m,n_sites,seed", CASES)
def test_matrices_and_ic_from_the_definitions(m, n_sites, seed):
    rng = np.random.default_rng(seed)
    sites = _random_sites(rng, m, n_sites)
    p, lo = _textbook_log_odds(sites)
    M = PSSMModel([SiteCollection(sites)], [1.0])
    got_p = np.array([[M.pwm[l][j] for l in "ACGT"] for j in range(m)])
    got_lo = np.array([[M.pssm[l][j] for l in "ACGT"] for j in range(m)])
    got_rc = np.array([[M.rev_comp_pssm[l][j] for l in "ACGT"] for j in range(m)])
    assert np.max(np.abs(got_p - p)) < 1e-15
    assert np.max(np.abs(got_lo - lo)) < 1e-13
    # reverse complement: position reversed, A<->T, C<->G
    assert np.array_equal(got_rc, got_lo[::-1, ::-1])
    # information content, closed form sum p log2(p / 0.25)
    ic = float(np.sum(p * np.log2(p / 0.25)))
    assert abs(M.IC - ic) < 1e-12
    # the oracle's shim says the same
    mot = shim.create([shim.Seq(s) for s in sites])
    mot.pseudocounts = 1
    sh = mot.pwm.log_odds({"A": .25, "C": .25, "G": .25, "T": .25})
    assert abs(sh.mean() - ic) < 1e-12


@pytest.mark.parametrize("m,n_sites,seed", CASES)
def test_patser_threshold_brackets_the_exact_distribution(m, n_sites, seed):
    """threshold_patser() walks the DP's background density from the top until the tail mass
    reaches 2^-IC.  The DP rounds every cell to a grid of `step`, so a window's grid score is within
    m * step / 2 of its true score: with that slack the threshold must sit exactly where the tail of
    the TRUE distribution (all 4^m windows) crosses 2^-IC."""
    rng = np.random.default_rng(seed)
    sites = _random_sites(rng, m, n_sites)
    p, lo = _textbook_log_odds(sites)
    M = PSSMModel([SiteCollection(sites)], [1.0])
    thr = M.threshold()
    ic = float(np.sum(p * np.log2(p / 0.25)))
    fpr = 2.0 ** -ic
    scores = np.sort(_all_window_scores(lo))
    n = scores.size
    mn, mx = min(0.0, lo.min(axis=1).sum()), max(0.0, lo.max(axis=1).sum())
    step = (mx - mn) / (1000 * m - 1)
    slack = (m / 2.0 + 1.0) * step

    def tail(t):                      # P(score >= t) under the uniform background
        return (n - np.searchsorted(scores, t, side="left")) / n

    assert tail(thr - slack) >= fpr * (1 - 1e-9)            # enough mass at or above the threshold
    assert tail(thr + slack) < fpr * (1 + 1e-9) or thr + slack > scores[-1]   # and not one grid cell later
    # the exact crossing point itself lies within the slack of the DP's answer
    k = int(np.ceil(fpr * n - 1e-9))                        # number of top windows needed
    exact_thr = scores[n - k] if k >= 1 else scores[-1]
    assert abs(thr - exact_thr) <= slack, (thr, exact_thr, slack)
    # the oracle's DP (independent code: Python, class by class) lands on the same grid point (its
    # PWM here is not re-normalised a second time as _combine_pwms does: last-bit differences)
    mot = shim.create([shim.Seq(s) for s in sites])
    mot.pseudocounts = 1
    sh = mot.pwm.log_odds({"A": .25, "C": .25, "G": .25, "T": .25})
    assert abs(sh.distribution(precision=1000).threshold_patser() - thr) < 1e-9


def test_hand_kat_acgt():
    """sites = ["ACGT"], pseudocount 1 -> 0.4 / 0.2, log2(1.6), log2(0.8)."""
    M = PSSMModel([SiteCollection(["ACGT"])], [1.0])
    assert abs(M.pssm["A"][0] - np.log2(1.6)) < 1e-15 and abs(M.pssm["C"][0] - np.log2(0.8)) < 1e-15
    assert abs(M.IC - 4 * (0.4 * np.log2(1.6) + 3 * 0.2 * np.log2(0.8))) < 1e-14


@pytest.mark.gpu
@pytest.mark.parametrize("m,n_sites,seed", [(5, 12, 2), (6, 20, 3), (8, 20, 5)])
def test_gpu_scores_of_every_window_from_the_definitions(m, n_sites, seed):
    """score_seq over a sequence that contains every m-mer: each window's score against the
    textbook sum of log-odds (float64 numpy, no oracle code), 1e-4 as north_star states; the
    soft-max of the two strands against log2(2^f + 2^r)."""
    import torch
    assert torch.cuda.is_available()
    rng = np.random.default_rng(seed)
    sites = _random_sites(rng, m, n_sites)
    _, lo = _textbook_log_odds(sites)
    M = PSSMModel([SiteCollection(sites)], [1.0])
    seq = "".join("".join(w) for w in itertools.product("ACGT", repeat=m))
    idx = np.frombuffer(seq.encode(), dtype=np.uint8)
    code = np.zeros(256, dtype=np.int64)
    for k, ch in enumerate(b"ACGT"):
        code[ch] = k
    b = code[idx]
    win = np.lib.stride_tricks.sliding_window_view(b, m)
    f = lo[np.arange(m)[None, :], win].sum(axis=1)
    r = lo[::-1, ::-1][np.arange(m)[None, :], win].sum(axis=1)
    got_f = np.asarray(M.score_seq(seq, both=False), dtype=np.float64)
    assert got_f.shape == f.shape and np.max(np.abs(got_f - f)) < 1e-4
    both = np.asarray(M.score_seq(seq, both=True), dtype=np.float64)
    assert np.max(np.abs(both - np.log2(2.0 ** f + 2.0 ** r))) < 1e-4
