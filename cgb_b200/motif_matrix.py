"""Host float64 motif matrices (the Biopython ``Bio.motifs.matrix`` objects the
reference's TF-binding model is written against, without Biopython).

Reference call sites (under /root/reference):
  cgb/site_collection.py:15-21,33-36   motifs.create(...).pwm, pseudocounts=1
  cgb/pssm_model.py:47-60,159-169      PositionWeightMatrix, log_odds, reverse_complement, mean
  cgb/pssm_model.py:62-70              distribution(precision).threshold_patser()

All arithmetic is plain Python/C double in the order Biopython 1.76 performs it,
so the values equal the reference's bit for bit on the same libm.  The Patser
score-distribution DP runs in the C++ host code of libcgbk
(``cgbk_patser_threshold``).
"""
from __future__ import annotations

import ctypes as C
import math

ALPHABET = "ACGT"


def psum(iterable, start=0):
    """Left-to-right summation, as the builtin ``sum`` of the reference's Python 2.7.
    (CPython >= 3.12 sums floats with compensation, which changes the last bits of
    the PWM mixture, pssm_model.py:166, and of the column means, pssm_model.py:100.)"""
    total = start
    for x in iterable:
        total = total + x
    return total


class PositionMatrix(dict):
    """letter -> sequence of per-position values; ``m[letter][pos]``,
    ``m[letter, pos]`` and ``m[:, pos]`` like Biopython's GenericPositionMatrix."""

    def __init__(self, alphabet, values):
        super().__init__()
        self.alphabet = alphabet
        self.length = None
        for letter in alphabet:
            if self.length is None:
                self.length = len(values[letter])
            elif self.length != len(values[letter]):
                raise Exception("data has inconsistent lengths")
            dict.__setitem__(self, letter, list(values[letter]))

    def __getitem__(self, key):
        if isinstance(key, tuple):
            if len(key) != 2:
                raise KeyError("keys should be 1- or 2-dimensional")
            k1, k2 = key
            if isinstance(k1, slice) and k1 == slice(None):
                return {l: dict.__getitem__(self, l)[k2] for l in self.alphabet}
            return dict.__getitem__(self, k1)[k2]
        return dict.__getitem__(self, key)

    def rows(self):
        """[pos][ACGT] nested list (the layout handed to the device)."""
        return [[dict.__getitem__(self, l)[i] for l in ALPHABET] for i in range(self.length)]


class PositionWeightMatrix(PositionMatrix):
    """Column-normalised probabilities (pssm_model.py:169 builds one directly)."""

    def __init__(self, alphabet, counts):
        super().__init__(alphabet, counts)
        cols = [dict.__getitem__(self, letter) for letter in alphabet]     # the lists, in alphabet order
        for i in range(self.length):
            total = 0                                   # psum: left to right from int 0
            for col in cols:
                total = total + float(col[i])
            for col in cols:
                col[i] /= total
        for letter, col in zip(alphabet, cols):
            dict.__setitem__(self, letter, tuple(col))

    def log_odds(self, background=None):
        """log2(p / b) per cell as ``math.log(p / b, 2)`` (pssm_model.py:50)."""
        values = {}
        alphabet = self.alphabet
        if background is None:
            background = dict.fromkeys(alphabet, 1.0)
        else:
            background = dict(background)
        total = psum(background.values())
        for letter in alphabet:
            background[letter] /= total
            values[letter] = []
        cols = {letter: dict.__getitem__(self, letter) for letter in alphabet}
        for i in range(self.length):
            for letter in alphabet:
                b = background[letter]
                p = cols[letter][i]
                if b > 0:
                    logodds = math.log(p / b, 2) if p > 0 else -math.inf
                else:
                    logodds = math.inf if p > 0 else math.nan
                values[letter].append(logodds)
        return PositionSpecificScoringMatrix(alphabet, values)


class PositionSpecificScoringMatrix(PositionMatrix):
    def reverse_complement(self):
        """pssm_model.py:52-55."""
        values = {"A": self["T"][::-1], "T": self["A"][::-1],
                  "G": self["C"][::-1], "C": self["G"][::-1]}
        return self.__class__(self.alphabet, values)

    @property
    def max(self):
        score = 0.0
        for position in range(self.length):
            score += max(self[letter][position] for letter in self.alphabet)
        return score

    @property
    def min(self):
        score = 0.0
        for position in range(self.length):
            score += min(self[letter][position] for letter in self.alphabet)
        return score

    def mean(self, background=None):
        """Expected score under the motif = information content (pssm_model.py:57-60)."""
        if background is None:
            background = dict.fromkeys(self.alphabet, 1.0)
        else:
            background = dict(background)
        total = psum(background.values())
        for letter in self.alphabet:
            background[letter] /= total
        # same terms in the same order as Biopython's double loop (position outer, letter inner),
        # over plain lists instead of keyed look-ups
        cols = [(background[letter], dict.__getitem__(self, letter)) for letter in self.alphabet]
        isnan, isinf, pow2 = math.isnan, math.isinf, math.pow
        sx = 0.0
        for i in range(self.length):
            for b, col in cols:
                logodds = col[i]
                if isnan(logodds):
                    continue
                if isinf(logodds) and logodds < 0:
                    continue
                p = b * pow2(2, logodds)
                sx += p * logodds
        return sx

    def distribution(self, background=None, precision=10 ** 3):
        if background is not None:
            vals = [background[l] for l in self.alphabet]
            if max(vals) - min(vals) > 0:
                raise NotImplementedError("only the uniform background of PSSMModel is supported")
        return ScoreDistribution(self, precision)


class ScoreDistribution(object):
    """Score distribution under the uniform background; only the Patser
    threshold the reference uses is exposed (pssm_model.py:69-70)."""

    def __init__(self, pssm, precision):
        self._pssm = pssm
        self._precision = precision

    def threshold_patser(self):
        from . import _cabi
        rows = self._pssm.rows()
        m = len(rows)
        flat = (C.c_double * (4 * m))(*[x for row in rows for x in row])
        thr = C.c_double()
        _cabi.check(_cabi.lib().cgbk_patser_threshold(flat, m, int(self._precision), C.byref(thr),
                                                      None, None, None))
        return thr.value


def counts_from_instances(instances, alphabet=ALPHABET):
    """``motifs.create(instances).counts`` (site_collection.py:18)."""
    length = len(instances[0])
    counts = {letter: [0] * length for letter in alphabet}
    for instance in instances:
        if len(instance) != length:
            raise ValueError("All instances should have the same length (%d found, %d expected)"
                             % (len(instance), length))
        for position, letter in enumerate(instance):
            counts[letter][position] += 1      # KeyError for a letter outside the alphabet, like Biopython
    return counts


def normalize_counts(counts, pseudocounts, alphabet=ALPHABET):
    """``counts.normalize(pseudocounts)`` -> PositionWeightMatrix (site_collection.py:36)."""
    length = len(counts[alphabet[0]])
    vals = {}
    for letter in alphabet:
        if pseudocounts is None:
            vals[letter] = [0.0] * length
        elif isinstance(pseudocounts, dict):
            vals[letter] = [float(pseudocounts[letter])] * length
        else:
            vals[letter] = [float(pseudocounts)] * length
    for i in range(length):
        for letter in alphabet:
            vals[letter][i] += counts[letter][i]
    return PositionWeightMatrix(alphabet, vals)
