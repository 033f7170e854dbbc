"""numpy float64 restatement of the cgb hot path (TEST ORACLE, not product).

Every function cites the reference file:line (relative to /root/reference) it
follows.  The Biopython 1.76 ``Bio.motifs`` arithmetic is not vendored in the
reference tree; it is restated here from its published algorithm and is
therefore PARITY UNPINNED at that layer (the reference ships no tests).  The
``pssm_model.py`` / ``binding_model.py`` layer IS pinned: ``tests/golden`` holds
outputs of those two files executed verbatim (oracle/ref_loader.py).

Numeric conventions restated here
---------------------------------
* ``PSSM.calculate`` (Biopython ``_pwm.c``): per window a C ``double``
  accumulates ``matrix[j][base]`` for j ascending; the result is stored as
  **float32**; any byte outside ``ACGTacgt`` makes the window NaN.
* ``score_seq`` soft-max (pssm_model.py:129-131): ``log2(2**s + 2**r)`` with
  ``s, r`` numpy float32 scalars.  Under the reference's pinned numpy
  (1.11 / 1.16, requirements.txt:6, conda_cgb_environment.yml:9) scalar⊕Python
  int promotes to float64, so ``2**s`` is a C ``pow`` in double
  (``softmax="legacy64"``, the default and the semantics the CUDA path
  implements).  Under numpy>=2 (NEP 50) the same source line computes
  ``powf`` in float32 (``softmax="nep50"``); that mode exists only to check
  this restatement bit-for-bit against the verbatim reference run in the
  build container (numpy 2.3).
* hit test (genome.py:436,443): ``np.float32 >= float`` compares in float64
  under the pinned numpy.
"""
from __future__ import annotations

import math
from collections import namedtuple

import numpy as np

ALPHABET = "ACGT"


def sum(iterable, start=0):   # noqa: A001
    """Left-to-right float summation of the reference's Python 2.7 (CPython >= 3.12
    made the builtin ``sum`` compensated, which changes last bits)."""
    total = start
    for x in iterable:
        total = total + x
    return total


_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}

# byte -> column (A0 C1 G2 T3), -1 for anything outside ACGTacgt (_pwm.c switch)
_CODE = np.full(256, -1, dtype=np.int8)
for _i, _l in enumerate(ALPHABET):
    _CODE[ord(_l)] = _i
    _CODE[ord(_l.lower())] = _i
# byte -> column for the pure-Python repair path (pssm_model.py:96-100): only
# the upper-case keys exist in the PSSM dict, everything else is a KeyError
_CODE_UPPER = np.full(256, -1, dtype=np.int8)
for _i, _l in enumerate(ALPHABET):
    _CODE_UPPER[ord(_l)] = _i


# ---------------------------------------------------------------------------
# a1  SiteCollection.pwm  (site_collection.py:15-21,33-36)
# ---------------------------------------------------------------------------
def pwm_from_sites(sites, pseudocounts=1):
    """``motifs.create(instances)``; ``pseudocounts=1``; ``motif.pwm``.

    Biopython: counts[l][i]; ``counts.normalize(pseudocounts)`` gives
    ``float(pseudo) + count`` and ``PositionWeightMatrix.__init__`` divides
    every column by ``sum(float(v[l][i]) for l in alphabet)`` (order ACGT).
    """
    m = len(sites[0])
    assert all(len(s) == m for s in sites)
    vals = {l: [float(pseudocounts)] * m for l in ALPHABET}
    for i in range(m):
        for l in ALPHABET:
            c = sum(1 for s in sites if s[i].upper() == l)
            vals[l][i] += c
    return position_weight_matrix(vals)


def position_weight_matrix(vals):
    """``Bio.motifs.matrix.PositionWeightMatrix(alphabet, values)``: column
    re-normalisation (used by site_collection.py:36 and pssm_model.py:169)."""
    vals = {l: list(vals[l]) for l in ALPHABET}
    m = len(vals["A"])
    for i in range(m):
        total = sum(float(vals[l][i]) for l in ALPHABET)
        for l in ALPHABET:
            vals[l][i] /= total
    return {l: tuple(vals[l]) for l in ALPHABET}


# ---------------------------------------------------------------------------
# a2  PSSMModel._combine_pwms  (pssm_model.py:159-169)
# ---------------------------------------------------------------------------
def combine_pwms(pwms, weights):
    m = len(pwms[0]["A"])
    assert all(len(p["A"]) == m for p in pwms)
    vals = {l: [sum(p[l][i] * w for p, w in zip(pwms, weights)) for i in range(m)]
            for l in ALPHABET}
    return position_weight_matrix(vals)


# ---------------------------------------------------------------------------
# a3  PSSMModel.pssm -> PositionWeightMatrix.log_odds  (pssm_model.py:47-50)
# ---------------------------------------------------------------------------
def log_odds(pwm, background=None):
    if background is None:
        background = dict.fromkeys(ALPHABET, 1.0)
    else:
        background = dict(background)
    total = sum(background.values())
    for l in ALPHABET:
        background[l] /= total
    m = len(pwm["A"])
    out = {l: [] for l in ALPHABET}
    for i in range(m):
        for l in ALPHABET:
            b = background[l]
            p = pwm[l][i]
            if b > 0:
                lo = math.log(p / b, 2) if p > 0 else -math.inf
            else:
                lo = math.inf if p > 0 else math.nan
            out[l].append(lo)
    return out


# a4  PSSMModel.rev_comp_pssm  (pssm_model.py:52-55)
def reverse_complement_pssm(pssm):
    return {"A": pssm["T"][::-1], "T": pssm["A"][::-1],
            "G": pssm["C"][::-1], "C": pssm["G"][::-1]}


# a5  PSSMModel.IC -> PSSM.mean()  (pssm_model.py:57-60)
def pssm_mean(pssm, background=None):
    if background is None:
        background = dict.fromkeys(ALPHABET, 1.0)
    else:
        background = dict(background)
    total = sum(background.values())
    for l in ALPHABET:
        background[l] /= total
    sx = 0.0
    for i in range(len(pssm["A"])):
        for l in ALPHABET:
            lo = pssm[l][i]
            if math.isnan(lo):
                continue
            if math.isinf(lo) and lo < 0:
                continue
            sx += background[l] * math.pow(2, lo) * lo
    return sx


def pssm_max(pssm):
    score = 0.0
    for i in range(len(pssm["A"])):
        score += max(pssm[l][i] for l in ALPHABET)
    return score


def pssm_min(pssm):
    score = 0.0
    for i in range(len(pssm["A"])):
        score += min(pssm[l][i] for l in ALPHABET)
    return score


# ---------------------------------------------------------------------------
# a6  PSSMModel.patser_threshold  (pssm_model.py:62-70)
#     -> pssm.distribution(precision=10**3).threshold_patser()
# ---------------------------------------------------------------------------
def patser_threshold(pssm, precision=10 ** 3, literal=False):
    """Bio.motifs.thresholds.ScoreDistribution + threshold_patser.

    ``literal=True`` runs the per-point Python loops exactly as Biopython
    writes them (slow); the default vectorises the same recurrence with numpy
    (additions into a given cell happen in the same letter order; the clamped
    end cells are accumulated with ``np.add.at`` in index order).
    """
    m = len(pssm["A"])
    bgd = {l: 0.25 for l in ALPHABET}          # distribution(background=None)
    min_score = min(0.0, pssm_min(pssm))
    interval = max(0.0, pssm_max(pssm)) - min_score
    n_points = precision * m
    ic = pssm_mean(pssm, bgd)
    step = interval / (n_points - 1)

    def index_diff(x, y=0.0):
        return int((x - y + 0.5 * step) // step)

    start = -index_diff(min_score)
    if literal:
        bg_density = [0.0] * n_points
        bg_density[start] = 1.0
        for pos in range(m):
            bg_new = [0.0] * n_points
            for l in ALPHABET:
                d = index_diff(pssm[l][pos])
                bg = bgd[l]
                for i in range(n_points):
                    bg_new[max(0, min(n_points - 1, i + d))] += bg_density[i] * bg
            bg_density = bg_new
    else:
        dens = np.zeros(n_points)
        dens[start] = 1.0
        idx = np.arange(n_points)
        for pos in range(m):
            new = np.zeros(n_points)
            for l in ALPHABET:
                d = index_diff(pssm[l][pos])
                tgt = np.clip(idx + d, 0, n_points - 1)
                np.add.at(new, tgt, dens * bgd[l])
            dens = new
        bg_density = dens
    fpr = 2 ** -ic
    i = n_points
    prob = 0.0
    while prob < fpr:
        i -= 1
        prob += bg_density[i]
    return min_score + i * step


# ---------------------------------------------------------------------------
# model container
# ---------------------------------------------------------------------------
class OracleModel:
    """State of a reference ``PSSMModel`` (pssm_model.py:31-35) as plain data."""

    def __init__(self, site_lists, weights, background=None, pseudocounts=1):
        self.site_lists = [list(s) for s in site_lists]
        self.collection_pwms = [pwm_from_sites(s, pseudocounts) for s in site_lists]
        self.weights = list(weights)
        self.background = background or {l: 0.25 for l in ALPHABET}
        self.pwm = combine_pwms(self.collection_pwms, self.weights)
        self.length = len(self.pwm["A"])
        self.pssm = log_odds(self.pwm, self.background)
        self.rev_comp_pssm = reverse_complement_pssm(self.pssm)
        # [pos][ACGT] float64 tables, the layout Biopython hands to _pwm.c
        self.fwd = np.array([[self.pssm[l][j] for l in ALPHABET]
                             for j in range(self.length)], dtype=np.float64)
        self.rc = np.array([[self.rev_comp_pssm[l][j] for l in ALPHABET]
                            for j in range(self.length)], dtype=np.float64)
        self._thr = None
        self._mu_m = self._std_m = self._mu_bg = self._std_bg = None

    @classmethod
    def from_logodds(cls, fwd):
        """Model straight from a [m][4] log-odds table (no sites)."""
        self = cls.__new__(cls)
        fwd = np.asarray(fwd, dtype=np.float64)
        self.site_lists, self.weights = [], []
        self.length = fwd.shape[0]
        self.pssm = {l: [float(fwd[j, k]) for j in range(self.length)]
                     for k, l in enumerate(ALPHABET)}
        self.rev_comp_pssm = reverse_complement_pssm(self.pssm)
        self.fwd = fwd.copy()
        self.rc = np.array([[self.rev_comp_pssm[l][j] for l in ALPHABET]
                            for j in range(self.length)], dtype=np.float64)
        self.background = {l: 0.25 for l in ALPHABET}
        self._thr = None
        self._mu_m = self._std_m = self._mu_bg = self._std_bg = None
        return self

    @property
    def sites(self):                      # pssm_model.py:79-82
        return [s for coll in self.site_lists for s in coll]

    @property
    def IC(self):                         # pssm_model.py:57-60
        return pssm_mean(self.pssm)

    def threshold(self, threshold_type="patser"):   # pssm_model.py:72-77
        if threshold_type != "patser":
            raise ValueError
        if self._thr is None:
            self._thr = patser_threshold(self.pssm)
        return self._thr


# ---------------------------------------------------------------------------
# a7  PSSM.calculate (Biopython _pwm.c)  and  PSSMModel.score_seq
# ---------------------------------------------------------------------------
def calculate(table, seq_bytes):
    """``_pwm.c:calculate``: float32 scores, NaN where a window holds a byte
    outside ACGTacgt.  ``table`` is [m][4] float64.  Returns a float32 array of
    n-m+1 entries (Biopython returns the bare scalar when that is 1; the caller
    handles that, pssm_model.py:117-119)."""
    seq = np.frombuffer(seq_bytes, dtype=np.uint8)
    m = table.shape[0]
    n = seq.shape[0]
    nwin = n - m + 1
    if nwin < 0:
        raise ValueError("negative dimensions are not allowed")
    if nwin == 0:                            # len(seq) == m - 1: numpy.empty(0), no error
        return np.empty(0, dtype=np.float32)
    codes = _CODE[seq]
    bad = codes < 0
    safe = np.where(bad, 0, codes).astype(np.intp)
    acc = np.zeros(nwin, dtype=np.float64)
    for j in range(m):                       # j ascending, double accumulate
        acc += table[j][safe[j:j + nwin]]
    badwin = np.convolve(bad.astype(np.int32), np.ones(m, dtype=np.int32), "valid") > 0
    out = acc.astype(np.float32)             # (float)score
    out[badwin] = np.nan
    return out


def _calculate_ambiguous(table, window_bytes):
    """PSSMModel._calculate (pssm_model.py:88-101): only upper-case ACGT are
    keys; anything else adds the column mean ``sum(A,C,G,T)/4``."""
    score = 0.0
    for pos, ch in enumerate(window_bytes):
        k = _CODE_UPPER[ch]
        if k >= 0:
            score += table[pos][k]
        else:
            score += sum(table[pos][c] for c in range(4)) / 4
    return score


def softmax2(s, r, softmax="legacy64"):
    """pssm_model.py:129-131 ``log2(2**score + 2**rc_score)`` (misc.py:14-16
    ``math.log(x, 2)``) for float32 inputs."""
    if softmax == "legacy64":
        return math.log(math.pow(2.0, float(s)) + math.pow(2.0, float(r)), 2)
    elif softmax == "nep50":
        # numpy >= 2: 2**np.float32 stays float32; a repaired single-window score is a
        # Python float and takes the double path in either numpy
        a = np.float32(2) ** s if isinstance(s, np.float32) else 2 ** s
        b = np.float32(2) ** r if isinstance(r, np.float32) else 2 ** r
        return math.log(a + b, 2)
    raise ValueError(softmax)


def score_seq(model, seq, both=True, softmax="legacy64"):
    """PSSMModel.score_seq (pssm_model.py:103-133).

    Returns (scores, rc_scores): with ``both`` the first is the Python list of
    soft-max scores; otherwise the float32 array (or 1-element list) the
    reference returns.  ``rc_scores`` is returned for inspection only.
    """
    if isinstance(seq, str):
        seq = seq.encode("ascii")
    m = model.length
    scores = calculate(model.fwd, seq)
    rc_scores = calculate(model.rc, seq)
    if m == len(seq):
        # Biopython returned bare np.float32 scalars; the reference wraps them
        # in Python lists, so a repaired value is NOT rounded to float32.
        scores, rc_scores = [scores[0]], [rc_scores[0]]
    for i in range(len(scores)):
        if math.isnan(scores[i]):
            site = seq[i:i + m]
            scores[i] = _calculate_ambiguous(model.fwd, site)
            rc_scores[i] = _calculate_ambiguous(model.rc, site)
    if both:
        scores = [softmax2(s, r, softmax) for s, r in zip(scores, rc_scores)]
    return scores, rc_scores


def score_seq_both_vec(model, seq, softmax="legacy64"):
    """Vectorised ``score_seq(both=True)`` for long sequences (float64 array).
    Same values as :func:`score_seq` in ``legacy64`` mode up to the last ulp of
    ``log``; used where the Python loop would take minutes."""
    assert softmax == "legacy64"
    if isinstance(seq, str):
        seq = seq.encode("ascii")
    m = model.length
    f = calculate(model.fwd, seq)
    r = calculate(model.rc, seq)
    for i in np.nonzero(np.isnan(f))[0]:
        site = seq[i:i + m]
        f[i] = _calculate_ambiguous(model.fwd, site)
        r[i] = _calculate_ambiguous(model.rc, site)
    f64 = f.astype(np.float64)
    r64 = r.astype(np.float64)
    return np.log(np.exp2(f64) + np.exp2(r64)) / math.log(2.0), f, r


# a9  score_self (pssm_model.py:84-86)
def score_self(model, softmax="legacy64"):
    return [score_seq(model, s, True, softmax)[0] for s in model.sites]


# a11 build_bayesian_estimator (binding_model.py:71-92)
def build_bayesian_estimator(model, bg_scores, softmax="legacy64"):
    pssm_scores = score_self(model, softmax)
    model._mu_m, model._std_m = np.mean(pssm_scores), np.std(pssm_scores)
    model._mu_bg, model._std_bg = np.mean(bg_scores), np.std(bg_scores)
    return model._mu_m, model._std_m, model._mu_bg, model._std_bg


def _norm_pdf(x, loc, scale):
    """scipy.stats.norm(loc, scale).pdf(x): exp(-y**2/2)/sqrt(2*pi)/scale."""
    y = (np.asarray(x, dtype=np.float64) - loc) / scale
    return np.exp(-y ** 2 / 2.0) / np.sqrt(2 * np.pi) / scale


# a12 binding_probability (binding_model.py:94-113)
def binding_probability(model, seq, p_motif, alpha=1 / 350.0, softmax="legacy64",
                        scores=None):
    if scores is None:
        scores = score_seq(model, seq, True, softmax)[0]
    pssm_scores = np.asarray(scores, dtype=np.float64)
    pdf_m = _norm_pdf(pssm_scores, model._mu_m, model._std_m)
    pdf_bg = _norm_pdf(pssm_scores, model._mu_bg, model._std_bg)
    lh_m = alpha * pdf_m + (1 - alpha) * pdf_bg
    lh_bg = pdf_bg
    with np.errstate(all="ignore"):
        lh_ratio = np.exp(np.sum(np.log(lh_bg) - np.log(lh_m)))
        return float(1 / (1 + lh_ratio * (1 - p_motif) / p_motif))


# ---------------------------------------------------------------------------
# a13/a14  Gene coordinate rules (gene.py:72-127) on plain gene tables
# ---------------------------------------------------------------------------
OGene = namedtuple("OGene", "start end strand")      # strand +1 / -1


def upstream_gene(genes, idx, index=None):
    """gene.py:58-70.  ``index[idx]`` is the gene's ``_index`` (chromid.py:101-121): the running
    number of its 'gene' feature, which also counts the compound-location genes left out of the
    list -- the reference indexes the shortened list with it (IndexError past its end)."""
    g = genes[idx]
    k = idx if index is None else index[idx]
    if g.strand == 1 and k > 0:
        return genes[k - 1]
    if g.strand != 1 and k < len(genes) - 1:
        return genes[k + 1]
    return None


def upstream_noncoding_region_location(genes, idx, chrom_len, up=None, down=50, index=None):
    """gene.py:72-104 (``if up:`` treats 0 like None)."""
    g = genes[idx]
    if g.strand == 1:
        if up:
            loc_start = max(0, g.start - up)
        elif upstream_gene(genes, idx, index):
            loc_start = min(upstream_gene(genes, idx, index).end - down, g.start)
        else:
            loc_start = 0
        return loc_start, g.start + down
    if up:
        loc_end = min(chrom_len, g.end + up)
    elif upstream_gene(genes, idx, index):
        loc_end = max(upstream_gene(genes, idx, index).start + down, g.end)
    else:
        loc_end = chrom_len
    return g.end - down, loc_end


def genes_from_features(features):
    """Chromid.genes (cgb/chromid.py:99-122) over plain feature dicts
    ``{type, start, end, strand, compound, qualifiers}``: rows
    ``(index, start, end, strand, locus_tag, product_type, product, protein_id)``."""
    rows, index = [], 0
    for f, next_f in zip(features, features[1:]):
        if f["type"] == "gene":
            locus_tag = f["qualifiers"]["locus_tag"]
            next_locus_tag = next_f["qualifiers"].get("locus_tag")
            product_f = next_f if locus_tag == next_locus_tag else None
            if f["compound"]:
                index += 1
                continue
            ptype = product_f["type"] if product_f else ""
            try:
                product = product_f["qualifiers"]["product"][0]
            except (TypeError, KeyError):
                product = ""
            protein = ""
            if ptype == "CDS":
                protein = product_f["qualifiers"].get("protein_id", [""])[0]
            rows.append((index, f["start"], f["end"], f["strand"], locus_tag[0], ptype, product, protein))
            index += 1
    return rows


def promoter_region_location(gene, chrom_len, up=300, down=50):
    """gene.py:111-127 (returns the slice bounds handed to subsequence)."""
    up = -up
    if gene.strand == 1:
        return max(0, gene.start + up), gene.start + down
    return gene.end - down, min(chrom_len, gene.end - up)


def subsequence(sequence, start, end):
    """chromid.py:92-97 forward strand: a plain Python slice (negative start
    wraps, end past the tail clamps)."""
    return sequence[start:end]


def reverse_complement(seq):
    """bio_utils.py:20-28 (Biopython complement keeps case; N stays N)."""
    tbl = bytes.maketrans(b"ACGTacgtNn", b"TGCAtgcaNn")
    if isinstance(seq, str):
        return seq.translate(str.maketrans("ACGTacgtNn", "TGCAtgcaNn"))[::-1]
    return seq.translate(tbl)[::-1]


# ---------------------------------------------------------------------------
# a16 Genome.identify_sites (genome.py:394-455)
# ---------------------------------------------------------------------------
OSite = namedtuple("OSite", "start end strand score gene")


def identify_sites(model, sequence, genes, up=None, down=50, threshold=None, index=None, skip_short=False):
    """Hit loop of genome.py:419-448 over one chromid.  ``gene`` in the result
    is the gene index.  float32 score vs float64 threshold (pinned numpy).  A region more than
    one base shorter than the motif raises inside Biopython's calculate() in the reference
    (ValueError, numpy.empty of a negative count); ``skip_short`` leaves such regions out instead."""
    if threshold is None:
        threshold = model.threshold()
    m = model.length
    L = len(sequence)
    sites = []
    for gi in range(len(genes)):
        start, end = upstream_noncoding_region_location(genes, gi, L, up, down, index)
        seq = subsequence(sequence, start, end)
        if len(seq) < m - 1 and skip_short:
            continue
        scores = score_seq(model, seq, both=False)[0]
        for i, score in enumerate(scores, start=start):
            if float(score) >= threshold:
                sites.append(OSite(i, i + m, 1, score, gi))
        rc_seq = reverse_complement(seq)
        rc_scores = score_seq(model, rc_seq, both=False)[0]
        for i, score in enumerate(reversed(rc_scores), start=start):
            if float(score) >= threshold:
                sites.append(OSite(i, i + m, -1, score, gi))
    sites.sort(key=lambda s: s.score, reverse=True)
    return sites


# a10 background population of Genome.build_PSSM_model (genome.py:215-226)
def upstream_background_scores(model, sequence, genes, down=50, index=None, softmax="legacy64"):
    """``bg_scores`` of genome.py:215-223 for one replicon with ``random.sample`` taken as the
    identity: every m-mer ``seq[start:start + m]``, ``start in xrange(len(seq) - m)`` (the last
    window is dropped), of every gene's ``upstream_noncoding_region_sequence()``, each scored by
    ``score_seq`` as a sequence of its own (one window: a repaired score stays float64,
    pssm_model.py:117-127).  Returns the flat list of soft-max scores in the reference's order."""
    if isinstance(sequence, str):
        sequence = sequence.encode("ascii")
    m = model.length
    out = []
    for i in range(len(genes)):
        a, b = upstream_noncoding_region_location(genes, i, len(sequence), None, down, index)
        region = subsequence(sequence, a, b)
        n = len(region) - m
        if n <= 0:
            continue
        if softmax == "legacy64":
            both, f, r = score_seq_both_vec(model, region[:n + m - 1], softmax)
            # windows with a non-ACGT byte were repaired: as single-window sequences their strand
            # scores are Python floats, not float32 -- redo those through score_seq
            bad = np.nonzero(np.convolve((_CODE[np.frombuffer(region[:n + m - 1], dtype=np.uint8)] < 0).astype(np.int32),
                                         np.ones(m, dtype=np.int32), "valid") > 0)[0]
            both = both.tolist()
            for w in bad:
                both[w] = score_seq(model, region[w:w + m], True, softmax)[0][0]
            out.extend(both)
        else:
            out.extend(score_seq(model, region[w:w + m], True, softmax)[0][0] for w in range(n))
    return out


def scan_hits_whole(model, seq, threshold):
    """identify_sites semantics over ONE region covering the whole sequence:
    returns sorted arrays (pos, strand, score32).  Minus-strand scores are the
    forward PSSM on reverse_complement(seq) (genome.py:440-445), i.e. the same
    addends as the rc matrix but accumulated from the far end."""
    if isinstance(seq, str):
        seq = seq.encode("ascii")
    m = model.length
    f = score_seq(model, seq, both=False)[0]
    rcs = score_seq(model, reverse_complement(seq), both=False)[0]
    f = np.asarray(f, dtype=np.float64)
    r = np.asarray(rcs, dtype=np.float64)[::-1]
    pf = np.nonzero(f >= threshold)[0]
    pr = np.nonzero(r >= threshold)[0]
    pos = np.concatenate([pf, pr]).astype(np.int64)
    strand = np.concatenate([np.ones(len(pf), np.int32), -np.ones(len(pr), np.int32)])
    score = np.concatenate([f[pf], r[pr]]).astype(np.float32)
    order = np.lexsort((-strand, pos))
    return pos[order], strand[order], score[order]


# ---------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------
def synthetic_genome(n, seed, n_frac=0.0, n_run=50):
    rng = np.random.Generator(np.random.PCG64(seed))
    arr = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)]
    arr = arr.copy()
    if n_frac > 0:
        k = max(1, int(n * n_frac / n_run))
        for s in rng.integers(0, max(1, n - n_run), size=k):
            arr[s:s + n_run] = ord("N")
    return arr.tobytes()


def synthetic_genes(n, n_genes):
    """Genes every n/n_genes bp, alternating strands, ~1 kbp long."""
    step = n // n_genes
    genes = []
    for i in range(n_genes):
        s = i * step + step // 3
        e = min(n, s + min(1000, step // 2))
        genes.append(OGene(s, e, 1 if i % 2 == 0 else -1))
    return genes


# ---------------------------------------------------------------------------
# identified_sites report rows (genome.py:457-494, gene.py:290-316)
# ---------------------------------------------------------------------------
def relative_distance_to_start(gene, site_start, site_end):
    """gene.py:290-316."""
    dist = 0
    if gene.strand == 1:
        if site_start <= gene.start:
            dist = site_start - gene.start
        else:
            dist = site_end - gene.start
    else:
        if site_end >= gene.end:
            dist = gene.end - site_end
        else:
            dist = gene.end - site_start
    return dist


def identified_sites_rows(sites, sequence, genes, promoter_up, accession="chromid", annotations=None):
    """genome.py:473-494 for OSite tuples of one chromid; ``annotations[gi]`` =
    (locus_tag, protein_id, product)."""
    rows = []
    for site in sites:
        gene = genes[site.gene]
        dist = relative_distance_to_start(gene, site.start, site.end)
        category = 'operator'
        if dist < 0:
            if abs(dist) > promoter_up:
                category = 'intergenic'
        site_seq = sequence[site.start:site.end]
        if site.strand == -1:
            site_seq = reverse_complement(site_seq)
        ann = annotations[site.gene] if annotations else ('', '', '')
        rows.append([site.score, site_seq, accession, site.start, site.end, site.strand, dist, category,
                     gene.strand, gene.start, gene.end, ann[0], ann[1], ann[2]])
    return rows
