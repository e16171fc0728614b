"""The drop-in, mechanically (CPU side): the reference's UNMODIFIED benchmark
drivers compile and link against include/accblas/*.cuh + libaccblas_b200.so.

The drivers include their kernel headers with quotes
(/root/reference/cuda/gemv_benchmark.cu:13-16), which resolve in the
includer's own directory first, so oracle/Makefile compiles them from a scratch
directory that holds copies of the drivers and fixture headers only; nothing of
the reference is copied into the repository.  Runs where /root/reference
exists (not on the GPU box: there the prebuilt binaries are run by
tests/test_gpu_dropin_cpp.py)."""
import os
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path("/root/reference/cuda")
BIN = ROOT / "oracle" / "_ref" / "bin"


@pytest.mark.skipif(not REFERENCE.exists(), reason="reference sources not present")
@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not present")
def test_reference_drivers_build_against_the_dropin_headers():
    lib = ROOT / "accessor-blas_b200" / "libaccblas_b200.so"
    if not lib.exists():
        pytest.skip("libaccblas_b200.so not built (run __graft_entry__.build())")
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    out = subprocess.run(["make", "-C", str(ROOT / "oracle"), "dropin"], env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    for name in ("gemv_benchmark", "dot_benchmark", "trsv_benchmark"):
        exe = BIN / name
        assert exe.exists(), name
        needed = subprocess.run(["ldd", str(exe)], capture_output=True, text=True).stdout
        line = [l for l in needed.splitlines() if "libaccblas_b200.so" in l]
        assert line and "not found" not in line[0], needed
        # the kernels come from the library, not from the reference's headers
        syms = subprocess.run(["nm", "-D", "--undefined-only", str(exe)], capture_output=True,
                              text=True).stdout
        assert "accblas_" in syms
    # nothing of the reference was left behind inside the repository
    tracked = subprocess.run(["git", "ls-files"], cwd=ROOT, capture_output=True, text=True).stdout
    stray = [f for f in tracked.splitlines()
             if (f.endswith("_benchmark.cu") or f.endswith("_memory.cuh")
                 or f.endswith("matrix_helper.cuh"))
             and not f.startswith("accessor-blas_b200/drivers/")]
    assert not stray, stray
