"""ctypes front-end of oracle/pwm_oracle.c (TEST ORACLE, not product).

``build()`` compiles the C restatement with gcc -O2 into
``oracle/libcgb_oracle.so`` (git-ignored; travels to the GPU box with gpurun).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "pwm_oracle.c")
_LIB = os.path.join(_HERE, "libcgb_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _LIB, _SRC, "-lm"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
        L.cgbo_calculate.argtypes = [C.c_char_p, C.c_int64, dp, C.c_int, fp]
        L.cgbo_score_seq.argtypes = [C.c_char_p, C.c_int64, dp, dp, C.c_int, fp, fp, dp, dp, dp]
        L.cgbo_scan_hits.argtypes = [C.c_char_p, C.c_int64, dp, dp, C.c_int, C.c_double,
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int32), fp, C.c_int64]
        L.cgbo_scan_hits.restype = C.c_int64
        L.cgbo_binding_probability.argtypes = [C.c_char_p, C.c_int64, dp, dp, C.c_int] + [C.c_double] * 6
        L.cgbo_binding_probability.restype = C.c_double
        L.cgbo_posterior_from_scores.argtypes = [dp, C.c_int64] + [C.c_double] * 6
        L.cgbo_posterior_from_scores.restype = C.c_double
        L.cgbo_background_moments.argtypes = [C.c_char_p, C.c_int64, dp, dp, C.c_int, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _tables(model):
    return (np.ascontiguousarray(model.fwd, dtype=np.float64),
            np.ascontiguousarray(model.rc, dtype=np.float64))


def score_seq(model, seq: bytes):
    """(f32, r32, both64) of pssm_model.py:103-133 for len(seq) > m."""
    fwd, rc = _tables(model)
    n = len(seq) - model.length + 1
    f = np.empty(n, np.float32)
    r = np.empty(n, np.float32)
    b = np.empty(n, np.float64)
    lib().cgbo_score_seq(seq, len(seq), _dp(fwd), _dp(rc), model.length, _fp(f), _fp(r), _dp(b),
                         None, None)
    return f, r, b


def scan_hits(model, seq: bytes, threshold: float, cap=None):
    """genome.py:433-445 over one region; arrays sorted by (pos, +strand first)."""
    fwd, rc = _tables(model)
    cap = cap or max(1024, len(seq) // 8)
    while True:
        pos = np.empty(cap, np.int64)
        strand = np.empty(cap, np.int32)
        score = np.empty(cap, np.float32)
        k = lib().cgbo_scan_hits(seq, len(seq), _dp(fwd), _dp(rc), model.length, threshold,
                                 pos.ctypes.data_as(C.POINTER(C.c_int64)),
                                 strand.ctypes.data_as(C.POINTER(C.c_int32)), _fp(score), cap)
        if k <= cap:
            break
        cap = int(k)
    pos, strand, score = pos[:k], strand[:k], score[:k]
    order = np.lexsort((-strand, pos))
    return pos[order], strand[order], score[order]


def binding_probability(model, seq: bytes, mu_m, std_m, mu_bg, std_bg, p_motif, alpha):
    fwd, rc = _tables(model)
    return lib().cgbo_binding_probability(seq, len(seq), _dp(fwd), _dp(rc), model.length,
                                          mu_m, std_m, mu_bg, std_bg, p_motif, alpha)


def background_moments(model, seq: bytes):
    fwd, rc = _tables(model)
    out = np.zeros(3, np.float64)
    lib().cgbo_background_moments(seq, len(seq), _dp(fwd), _dp(rc), model.length, _dp(out))
    return out
