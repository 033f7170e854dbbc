"""In-tree build of libcgbk.so (hand-written CUDA for sm_100a + the C ABI).

    python -m cgb_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The library is written to
``cgb_b200/_lib/libcgbk.so`` (git-ignored, travels with gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "libcgbk.so")
SOURCES = ["cgbk_fast_stats.cu", "cgbk_fast_lean.cu", "cgbk_fast_moments.cu", "cgbk_fast_lean_moments.cu", "cgbk_fast_filter.cu", "cgbk_api.cu", "cgbk_exact.cu",
           "cgbk_pipeline.cu", "cgbk_sort.cu", "cgbk_hostpack.cu", "cgbk_genbank.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# exact kernels must not contract a*b+c into FMA: they restate CPU double arithmetic
EXTRA = {"cgbk_exact.cu": ["-fmad=false"], "cgbk_fast_stats.cu": ["-fmad=false"],
         "cgbk_fast_moments.cu": ["-fmad=false"], "cgbk_fast_filter.cu": ["-fmad=false"],
         "cgbk_fast_lean.cu": ["-fmad=false"], "cgbk_fast_lean_moments.cu": ["-fmad=false"]}


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "cgbk.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, defines=(), out=None):
    """``defines``/``out``: tuning builds (e.g. -DCGBK_STATS_THREADS=384) written next to the
    default library and selected at run time with the CGBK_LIB environment variable."""
    global LIB
    if out is not None:
        return _build(os.path.join(LIBDIR, out), os.path.join(LIBDIR, "obj_" + out), verbose, list(defines))
    if not force and not _stale():
        return LIB
    return _build(LIB, os.path.join(LIBDIR, "obj"), verbose, list(defines))


def _build(LIB, objdir, verbose, defines):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    common = [nvcc, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "-I", os.path.join(ROOT, "include"), "-I", CSRC] + ARCH + defines
    if verbose:
        common += ["-Xptxas", "-v"]

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = common + EXTRA.get(src, []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ARCH
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
