"""ctypes binding of libcgbk.so (include/cgbk.h).

There is no CPU fallback: if the CUDA library has not been built, importing a
compute entry point raises.  Status codes are mapped to the exception types the
reference raises for the same condition.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CGBK_LIB: alternative build of the library (kernel tuning experiments)
LIB_PATH = os.environ.get("CGBK_LIB") or os.path.join(_HERE, "_lib", "libcgbk.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3
MAX_MOTIF = 32
SEQ_ALIGN = 512
MAX_HIST_BINS = 4096
CTR_CANDIDATES, CTR_HITS, CTR_DIRTY_TILES, CTR_COUNT = 0, 1, 2, 8
BG_N, BG_SUM, BG_SUMSQ, BG_MEAN, BG_STD, BG_COUNT = 0, 1, 2, 3, 4, 8
RC_MATRIX_ORDER, RC_REVSEQ_ORDER, SINGLE_WINDOW_F64 = 0, 1, 2


class Genome(C.Structure):
    _fields_ = [("d_bases2", C.c_void_p), ("d_amb", C.c_void_p), ("d_lower", C.c_void_p),
                ("cap_bases", C.c_int64)]


class PipelineArgs(C.Structure):
    """cgbk_pipeline_args (include/cgbk.h)."""
    _fields_ = [("model", C.c_void_p), ("h_ascii", C.c_void_p), ("n_bases", C.c_int64),
                ("h_start", C.c_void_p), ("h_end", C.c_void_p), ("n_regions", C.c_int32), ("n_bins", C.c_int32),
                ("mu_m", C.c_double), ("std_m", C.c_double), ("p_motif", C.c_double), ("alpha", C.c_double),
                ("hist_lo", C.c_double), ("hist_hi", C.c_double),
                ("d_scratch", C.c_void_p), ("scratch_bytes", C.c_int64),
                ("h_staging", C.c_void_p), ("staging_bytes", C.c_int64),
                ("h_post", C.c_void_p), ("h_stats", C.c_void_p), ("h_hist", C.c_void_p), ("stream", C.c_void_p),
                ("h_bases2", C.c_void_p), ("h_amb", C.c_void_p), ("h_lower", C.c_void_p)]


class CgbkError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); every symbol include/cgbk.h declares
SYMBOLS = {
    "cgbk_version": (C.c_int, []),
    "cgbk_last_error": (C.c_char_p, []),
    "cgbk_patser_threshold": (C.c_int, [C.POINTER(C.c_double), C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 4),
    "cgbk_patser_threshold_batch": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int] +
                                    [C.POINTER(C.c_double)] * 4 + [C.POINTER(C.c_int)]),
    "cgbk_bases2_words": (C.c_int64, [C.c_int64]),
    "cgbk_pack_ascii": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_void_p]),
    "cgbk_model_create": (C.c_int, [C.POINTER(C.c_double), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "cgbk_model_destroy": (None, [C.c_void_p]),
    "cgbk_seqset_table_ints": (C.c_int64, [C.c_int]),
    "cgbk_seqset_create": (C.c_int, [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "cgbk_seqset_destroy": (None, [C.c_void_p]),
    "cgbk_seqset_total_windows": (C.c_int64, [C.c_void_p, C.c_int]),
    "cgbk_score_windows": (C.c_int, [C.c_void_p, C.POINTER(Genome), C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                      C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "cgbk_posteriors": (C.c_int, [C.c_void_p, C.POINTER(Genome), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                                   C.c_void_p]),
    "cgbk_score_plane_ints": (C.c_int64, [C.c_void_p, C.c_int]),
    "cgbk_scan_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_int, C.c_int64]),
    "cgbk_scan_hits": (C.c_int, [C.c_void_p, C.POINTER(Genome), C.c_void_p, C.c_double, C.c_int, C.c_void_p,
                                  C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "cgbk_background_stats": (C.c_int, [C.c_void_p, C.POINTER(Genome), C.c_void_p, C.c_double, C.c_double,
                                         C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_int64, C.c_void_p]),
    "cgbk_background_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cgbk_background_finalize_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "cgbk_profile_batch_ms": (C.c_int, [C.POINTER(C.c_float), C.c_int]),
    "cgbk_model_batch_create": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p,
                                           C.POINTER(C.c_void_p)]),
    "cgbk_model_batch_get": (C.c_void_p, [C.c_void_p, C.c_int]),
    "cgbk_model_batch_size": (C.c_int, [C.c_void_p]),
    "cgbk_model_batch_destroy": (None, [C.c_void_p]),
    "cgbk_score_sites_host": (C.c_int, [C.POINTER(C.c_double), C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "cgbk_region_stats_workspace_bytes": (C.c_int64, []),
    "cgbk_background_stats_regions": (C.c_int, [C.c_void_p, C.POINTER(Genome), C.c_void_p, C.c_void_p, C.c_int,
                                                 C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                 C.c_void_p, C.c_int64, C.c_void_p]),
    "cgbk_scan_batch_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_int64, C.c_int64]),
    "cgbk_scan_batch": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Genome), C.c_void_p, C.c_int64,
                                   C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "cgbk_sort_hits_workspace_bytes": (C.c_int64, [C.c_int64]),
    "cgbk_sort_hits": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "cgbk_split_sites_blocks": (C.c_int64, [C.c_int64]),
    "cgbk_split_sites": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cgbk_split_sites_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "cgbk_pipeline_scratch_bytes": (C.c_int64, [C.c_int64, C.c_int, C.c_int, C.c_int]),
    "cgbk_pipeline_staging_bytes": (C.c_int64, [C.c_int]),
    "cgbk_regulation_pipeline": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_double, C.c_double, C.c_double, C.c_double,
                                            C.c_double, C.c_double, C.c_int,
                                            C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cgbk_regulation_pipeline_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                   C.c_void_p, C.c_void_p, C.c_int,
                                                   C.c_double, C.c_double, C.c_double, C.c_double,
                                                   C.c_double, C.c_double, C.c_int,
                                                   C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cgbk_regulation_pipeline_args": (C.c_int, [C.c_void_p]),
    "cgbk_genbank_index": (C.c_int, [C.c_char_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "cgbk_genbank_origin": (C.c_int, [C.c_char_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "cgbk_packed_capacity": (C.c_int64, [C.c_int64]),
    "cgbk_pack_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                  C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "cgbk_last_launch_info": (C.c_int, [C.POINTER(C.c_int64)]),
    "cgbk_profile_enable": (C.c_int, [C.c_int]),
    "cgbk_profile_last_ms": (C.c_int, [C.POINTER(C.c_float)]),
}


def lib():
    """The loaded library; raises if it was never built (no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CgbkError(
                "libcgbk.so is missing (%s): build it with `python -m cgb_b200.build`; "
                "cgb_b200 has no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.cgbk_version() != 211:
            raise CgbkError("libcgbk.so version %d does not match the Python layer" % L.cgbk_version())
        _lib = L
    return _lib


def check(status):
    """Map a cgbk status code to the reference's exception types."""
    if status == OK:
        return
    msg = lib().cgbk_last_error().decode("utf-8", "replace")
    if status == ERR_INVALID:
        raise ValueError(msg)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise CgbkError(msg)


def profile_enable(on=True):
    """False / True, or 2: per-model kernel times of cgbk_scan_batch as well."""
    check(lib().cgbk_profile_enable(int(on)))


def profile_batch_ms(n):
    out = (C.c_float * n)()
    check(lib().cgbk_profile_batch_ms(out, n))
    return list(out)


def profile_last_ms():
    ms = C.c_float()
    check(lib().cgbk_profile_last_ms(C.byref(ms)))
    return ms.value


def last_launch_info():
    out = (C.c_int64 * 3)()
    check(lib().cgbk_last_launch_info(out))
    return {"launches": out[0], "tiles": out[1], "lane_bases": out[2]}
