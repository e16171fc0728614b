"""In-tree build of the sm_100a libraries (plain nvcc, no torch dependency).

Outputs (git-ignored, shipped to the GPU box by gpurun):
  accessor-blas_b200/libaccblas_b200.so        the product: kernels + C ABI
  accessor-blas_b200/libaccblas_baselines.so   cuBLAS baselines (bench only)
  accessor-blas_b200/libaccblas_b200_dev.so    build_dev(): the product sources
                                               with -DACCBLAS_DEV_HOOKS (timeline
                                               probes for tools/*_trace.py)
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
CSRC = HERE / "csrc"
BUILD = ROOT / "build" / "accblas"
LIB = HERE / "libaccblas_b200.so"
BASELINES_LIB = HERE / "libaccblas_baselines.so"

KERNEL_SOURCES = ["capi.cu", "dot.cu", "gemv.cu", "trsv.cu", "trsv_cluster_f64.cu", "trsv_cluster_f32.cu",
                  "convert_fill.cu"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC",
    f"-I{ROOT / 'include'}", f"-I{CSRC}",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found: the accblas libraries cannot be built")
    return nvcc


def _run(cmd: list[str]) -> None:
    env = dict(os.environ)
    # the image exports CXX/CC wrappers that break nvcc's host compiler lookup
    env.pop("CXX", None)
    env.pop("CC", None)
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("build step failed: " + " ".join(cmd))


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA translation unit for sm_100a and link the libraries."""
    nvcc = _nvcc()
    BUILD.mkdir(parents=True, exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + \
        list((ROOT / "include").glob("*.h"))

    def compile_one(name: str) -> Path:
        src = CSRC / name
        obj = BUILD / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *COMMON_FLAGS, *ARCH_FLAGS, "-c", str(src), "-o", str(obj)]
            if verbose:
                print(" ".join(cmd))
            _run(cmd)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as pool:
        objs = list(pool.map(compile_one, KERNEL_SOURCES + ["baselines.cu"]))
    kernel_objs = objs[:-1]
    baseline_obj = objs[-1]

    if force or _stale(LIB, kernel_objs):
        _run([nvcc, "-shared", *ARCH_FLAGS, "-o", str(LIB),
              *map(str, kernel_objs), "-cudart", "static"])
    if force or _stale(BASELINES_LIB, [baseline_obj]):
        _run([nvcc, "-shared", "-o", str(BASELINES_LIB), str(baseline_obj),
              "-cudart", "static", "-lcublas",
              "-Xlinker", "-rpath=/usr/local/cuda/lib64"])
    build_programs(force=force, verbose=verbose)
    return LIB


DEV_LIB = HERE / "libaccblas_b200_dev.so"


def build_dev(force: bool = False, verbose: bool = False) -> Path:
    """The same sources with the development hooks compiled in (phase
    timestamps of TRSV / GEMV).  Used by tools/ only; the product library never
    carries them."""
    nvcc = _nvcc()
    out = ROOT / "build" / "accblas_dev"
    out.mkdir(parents=True, exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + \
        list((ROOT / "include").glob("*.h"))

    def compile_one(name: str) -> Path:
        src = CSRC / name
        obj = out / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *COMMON_FLAGS, "-DACCBLAS_DEV_HOOKS", *ARCH_FLAGS, "-c",
                   str(src), "-o", str(obj)]
            if verbose:
                print(" ".join(cmd))
            _run(cmd)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as pool:
        objs = list(pool.map(compile_one, KERNEL_SOURCES))
    if force or _stale(DEV_LIB, objs):
        _run([nvcc, "-shared", *ARCH_FLAGS, "-o", str(DEV_LIB), *map(str, objs),
              "-cudart", "static"])
    return DEV_LIB


BIN = HERE / "bin"
PROGRAMS = {
    # name: (source, extra link flags)
    "gemv_benchmark": (HERE / "drivers" / "gemv_benchmark.cu", []),
    "dot_benchmark": (HERE / "drivers" / "dot_benchmark.cu", []),
    "trsv_benchmark": (HERE / "drivers" / "trsv_benchmark.cu", ["-lcusolver"]),
    "dropin_test": (ROOT / "tests" / "cpp" / "dropin_test.cu", []),
}


def build_programs(force: bool = False, verbose: bool = False) -> None:
    """The three benchmark drivers (reference CSV format) and the C++ test of
    the drop-in header layer, all compiled against include/accblas/*.cuh the
    way the reference's drivers are compiled against its cuda/*.cuh."""
    nvcc = _nvcc()
    BIN.mkdir(exist_ok=True)
    headers = list((ROOT / "include").rglob("*.*")) + \
        list((HERE / "drivers").glob("*.cuh")) + [LIB]

    def one(item):
        name, (src, extra) = item
        exe = BIN / name
        if force or _stale(exe, [src] + headers):
            cmd = [nvcc, "-std=c++17", "-O3", "--expt-relaxed-constexpr", "-lineinfo",
                   *ARCH_FLAGS, f"-I{ROOT / 'include'}", str(src), "-o", str(exe),
                   f"-L{HERE}", "-laccblas_b200", "-lcublas", *extra,
                   "-Xlinker", "-rpath=$ORIGIN/..",
                   "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
            if verbose:
                print(" ".join(cmd))
            _run(cmd)
        return exe

    with concurrent.futures.ThreadPoolExecutor(max_workers=4) as pool:
        list(pool.map(one, PROGRAMS.items()))


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    if "--dev" in sys.argv:
        print(build_dev(force="--force" in sys.argv, verbose=True))
