"""The C ABI from a non-Python host: examples/cabi_demo.c is compiled with gcc as C99 against
include/cgbk.h (so the header is valid C), linked against libcgbk.so, and run -- host-only here,
with the device part on the GPU box."""
import os
import shutil
import subprocess

import pytest

from cgb_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "cabi_demo")


def _build():
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    libdir = os.path.dirname(_cabi.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(cuda, "include"), os.path.join(ROOT, "examples", "cabi_demo.c"),
           "-L", libdir, "-lcgbk", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + libdir, "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_c_host_compiles_and_runs_host_part():
    _cabi.lib()                                   # builds / finds libcgbk.so
    exe = _build()
    r = subprocess.run([exe, "--host-only"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host-only checks passed" in r.stdout


@pytest.mark.gpu
def test_c_host_full_run_on_device():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
