"""Minimal stand-in for the parts of Biopython 1.76 the cgb hot path calls
(TEST ORACLE, not product).

Biopython is a dependency of the reference (requirements.txt:1 ``==1.66``,
conda_cgb_environment.yml:13 ``=1.76``; the string-alphabet call sites at
pssm_model.py:32,113,169 only work with 1.7x) that is neither vendored under
/root/reference nor installed here, and there is no network.  These classes
restate the published algorithm of ``Bio.Seq.Seq``, ``Bio.motifs.create``,
``Bio.motifs.matrix`` and ``Bio.motifs.thresholds`` so that the reference's own
``pssm_model.py`` / ``binding_model.py`` / ``site_collection.py`` can be executed
VERBATIM on top of them (oracle/ref_loader.py).  PARITY UNPINNED at this layer:
the reference has no tests or golden outputs to check the shim against.

Written class-by-class after Biopython's structure, deliberately independent of
the flat restatement in np_oracle.py; tests cross-check the two.
"""
from __future__ import annotations

import math

import numpy as np


def sum(iterable, start=0):   # noqa: A001 - Python 2.7 left-to-right float summation
    total = start
    for x in iterable:
        total = total + x
    return total


class Seq(object):
    """Bio.Seq.Seq: an immutable string with an (ignored) alphabet."""

    def __init__(self, data, alphabet=None):
        self._data = str(data)
        self.alphabet = alphabet

    def __len__(self):
        return len(self._data)

    def __str__(self):
        return self._data

    def __getitem__(self, index):
        if isinstance(index, int):
            return self._data[index]
        return Seq(self._data[index], self.alphabet)

    def __iter__(self):
        return iter(self._data)

    def __eq__(self, other):
        return str(self) == str(other)

    def __hash__(self):
        return hash(self._data)

    _COMP = str.maketrans("ACGTacgtNnRYKMrykmBVDHbvdhSWsw",
                          "TGCAtgcaNnYRMKyrmkVBHDvbhdSWsw")

    def complement(self):
        return Seq(self._data.translate(self._COMP), self.alphabet)

    def reverse_complement(self):
        return Seq(self._data.translate(self._COMP)[::-1], self.alphabet)


class GenericPositionMatrix(dict):
    """Bio.motifs.matrix.GenericPositionMatrix (letter -> list of values)."""

    def __init__(self, alphabet, values):
        self.length = None
        for letter in alphabet:
            if self.length is None:
                self.length = len(values[letter])
            elif self.length != len(values[letter]):
                raise Exception("data has inconsistent lengths")
            self[letter] = list(values[letter])
        self.alphabet = alphabet

    def __getitem__(self, key):
        if isinstance(key, tuple):
            if len(key) != 2:
                raise KeyError("keys should be 1- or 2-dimensional")
            key1, key2 = key
            if isinstance(key1, slice) and key1 == slice(None):
                return {l: dict.__getitem__(self, l)[key2] for l in self.alphabet}
            return dict.__getitem__(self, key1)[key2]
        return dict.__getitem__(self, key)


class FrequencyPositionMatrix(GenericPositionMatrix):
    def normalize(self, pseudocounts=None):
        counts = {}
        if pseudocounts is None:
            for letter in self.alphabet:
                counts[letter] = [0.0] * self.length
        elif isinstance(pseudocounts, dict):
            for letter in self.alphabet:
                counts[letter] = [float(pseudocounts[letter])] * self.length
        else:
            for letter in self.alphabet:
                counts[letter] = [float(pseudocounts)] * self.length
        for i in range(self.length):
            for letter in self.alphabet:
                counts[letter][i] += self[letter][i]
        return PositionWeightMatrix(self.alphabet, counts)


class PositionWeightMatrix(GenericPositionMatrix):
    def __init__(self, alphabet, counts):
        GenericPositionMatrix.__init__(self, alphabet, counts)
        for i in range(self.length):
            total = sum(float(self[letter][i]) for letter in alphabet)
            for letter in alphabet:
                self[letter][i] /= total
        for letter in alphabet:
            self[letter] = tuple(self[letter])

    def log_odds(self, background=None):
        values = {}
        alphabet = self.alphabet
        if background is None:
            background = dict.fromkeys(self.alphabet, 1.0)
        else:
            background = dict(background)
        total = sum(background.values())
        for letter in alphabet:
            background[letter] /= total
            values[letter] = []
        for i in range(self.length):
            for letter in alphabet:
                b = background[letter]
                if b > 0:
                    p = self[letter][i]
                    if p > 0:
                        logodds = math.log(p / b, 2)
                    else:
                        logodds = -math.inf
                else:
                    p = self[letter][i]
                    if p > 0:
                        logodds = math.inf
                    else:
                        logodds = math.nan
                values[letter].append(logodds)
        return PositionSpecificScoringMatrix(alphabet, values)


def _pwm_calculate(sequence, logodds):
    """Bio/motifs/_pwm.c ``calculate``: double accumulate, float32 store."""
    seq = sequence.encode("ascii") if isinstance(sequence, str) else bytes(sequence)
    m = len(logodds)
    n = len(seq) - m + 1
    if n < 0:
        raise ValueError("negative dimensions are not allowed")
    scores = np.empty(n, dtype=np.float32)
    a, c, g, t = ord("A"), ord("C"), ord("G"), ord("T")
    la, lc, lg, lt = ord("a"), ord("c"), ord("g"), ord("t")
    for i in range(n):
        score = 0.0
        ok = True
        for j in range(m):
            ch = seq[i + j]
            if ch == a or ch == la:
                score += logodds[j][0]
            elif ch == c or ch == lc:
                score += logodds[j][1]
            elif ch == g or ch == lg:
                score += logodds[j][2]
            elif ch == t or ch == lt:
                score += logodds[j][3]
            else:
                ok = False
        scores[i] = np.float32(score) if ok else np.float32("nan")
    return scores


class PositionSpecificScoringMatrix(GenericPositionMatrix):
    def calculate(self, sequence):
        if sorted(self.alphabet) != ["A", "C", "G", "T"]:
            raise ValueError("PSSM has wrong alphabet: %s - Use only with DNA motifs"
                             % self.alphabet)
        sequence = str(sequence)
        m = self.length
        logodds = [[self[letter][i] for letter in "ACGT"] for i in range(m)]
        scores = _pwm_calculate(sequence, logodds)
        if len(scores) == 1:
            return scores[0]
        return scores

    @property
    def max(self):
        score = 0.0
        for position in range(0, self.length):
            score += max(self[letter][position] for letter in self.alphabet)
        return score

    @property
    def min(self):
        score = 0.0
        for position in range(0, self.length):
            score += min(self[letter][position] for letter in self.alphabet)
        return score

    def mean(self, background=None):
        if background is None:
            background = dict.fromkeys(self.alphabet, 1.0)
        else:
            background = dict(background)
        total = sum(background.values())
        for letter in self.alphabet:
            background[letter] /= total
        sx = 0.0
        for i in range(self.length):
            for letter in self.alphabet:
                logodds = self[letter, i]
                if math.isnan(logodds):
                    continue
                if math.isinf(logodds) and logodds < 0:
                    continue
                b = background[letter]
                p = b * math.pow(2, logodds)
                sx += p * logodds
        return sx

    def reverse_complement(self):
        values = {}
        values["A"] = self["T"][::-1]
        values["T"] = self["A"][::-1]
        values["G"] = self["C"][::-1]
        values["C"] = self["G"][::-1]
        return self.__class__(self.alphabet, values)

    def distribution(self, background=None, precision=10 ** 3):
        if background is None:
            background = dict.fromkeys(self.alphabet, 1.0)
        else:
            background = dict(background)
        total = sum(background.values())
        for letter in self.alphabet:
            background[letter] /= total
        return ScoreDistribution(precision=precision, pssm=self, background=background)


class ScoreDistribution(object):
    """Bio.motifs.thresholds.ScoreDistribution (pssm= branch only)."""

    def __init__(self, motif=None, precision=10 ** 3, pssm=None, background=None):
        assert pssm is not None
        self.min_score = min(0.0, pssm.min)
        self.interval = max(0.0, pssm.max) - self.min_score
        self.n_points = precision * pssm.length
        self.ic = pssm.mean(background)
        self.step = self.interval / (self.n_points - 1)
        self.mo_density = [0.0] * self.n_points
        self.mo_density[-self._index_diff(self.min_score)] = 1.0
        self.bg_density = [0.0] * self.n_points
        self.bg_density[-self._index_diff(self.min_score)] = 1.0
        n = self.n_points
        for position in range(pssm.length):
            mo_new = np.zeros(n)
            bg_new = np.zeros(n)
            lo = pssm[:, position]
            mo_old = np.asarray(self.mo_density)
            bg_old = np.asarray(self.bg_density)
            for letter, score in lo.items():
                bg = background[letter]
                mo = pow(2, pssm[letter, position]) * bg
                d = self._index_diff(score)
                # for i in range(n): new[_add(i, d)] += old[i] * x   (same
                # accumulation order, done with an unbuffered scatter-add)
                tgt = np.clip(np.arange(n) + d, 0, n - 1)
                np.add.at(mo_new, tgt, mo_old * mo)
                np.add.at(bg_new, tgt, bg_old * bg)
            self.mo_density = mo_new
            self.bg_density = bg_new

    def _index_diff(self, x, y=0.0):
        return int((x - y + 0.5 * self.step) // self.step)

    def _add(self, i, j):
        return max(0, min(self.n_points - 1, i + j))

    def threshold_fpr(self, fpr):
        i = self.n_points
        prob = 0.0
        while prob < fpr:
            i -= 1
            prob += self.bg_density[i]
        return self.min_score + i * self.step

    def threshold_fnr(self, fnr):
        i = -1
        prob = 0.0
        while prob < fnr:
            i += 1
            prob += self.mo_density[i]
        return self.min_score + i * self.step

    def threshold_patser(self):
        return self.threshold_fpr(fpr=2 ** -self.ic)


class Instances(list):
    def __init__(self, instances, alphabet="ACGT"):
        super().__init__(instances)
        self.alphabet = alphabet
        self.length = len(instances[0]) if instances else 0

    def count(self):
        counts = {letter: [0] * self.length for letter in self.alphabet}
        for instance in self:
            for position, letter in enumerate(str(instance)):
                counts[letter][position] += 1
        return counts


class Motif(object):
    """Bio.motifs.Motif built from instances (motifs.create)."""

    def __init__(self, alphabet="ACGT", instances=None):
        self.name = ""
        self.instances = instances
        self.alphabet = alphabet
        self.counts = FrequencyPositionMatrix(alphabet, instances.count())
        self.length = self.counts.length
        self.pseudocounts = None
        self.background = None

    @property
    def pseudocounts(self):
        return self._pseudocounts

    @pseudocounts.setter
    def pseudocounts(self, value):
        self._pseudocounts = {}
        if isinstance(value, dict):
            self._pseudocounts = {l: value[l] for l in self.alphabet}
        else:
            if value is None:
                value = 0.0
            self._pseudocounts = dict.fromkeys(self.alphabet, value)

    @property
    def pwm(self):
        return self.counts.normalize(self._pseudocounts)

    @property
    def pssm(self):
        return self.pwm.log_odds(self.background)


def create(instances, alphabet="ACGT"):
    """Bio.motifs.create."""
    return Motif(instances=Instances(instances, alphabet), alphabet=alphabet)
