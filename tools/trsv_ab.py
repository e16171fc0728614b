"""Same-box A/B of two builds of the library on TRSV (development tool).

    python tools/trsv_ab.py [old.so]

Latency-bound kernels differ by tens of percent between gpurun boxes, so two
versions are only ever compared inside one process, interleaved.  The "old"
library is built from an earlier commit, e.g.
    git archive <commit> accessor-blas_b200/csrc include | tar -x -C /tmp/old
    nvcc -std=c++17 -O3 -Xcompiler -fPIC -I/tmp/old/include -I/tmp/old/accessor-blas_b200/csrc \
         -gencode arch=compute_100a,code=sm_100a -shared -cudart static \
         -o tools/micro/libaccblas_old.so /tmp/old/accessor-blas_b200/csrc/{capi,dot,gemv,trsv,convert_fill}.cu
"""
import ctypes
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import accessor_blas_b200 as ab  # noqa: E402
from accessor_blas_b200 import capi  # noqa: E402

old_path = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "tools" / "micro" / "libaccblas_old.so")
libs = {"new": capi.load(), "old": ctypes.CDLL(old_path)}
P = ctypes.c_void_p
handles = {}
for name, lib in libs.items():
    lib.accblas_create.argtypes = [ctypes.POINTER(P), ctypes.c_int]
    lib.accblas_trsv.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, P,
                                 ctypes.c_int64, P, ctypes.c_int64, P]
    hd = P()
    assert lib.accblas_create(ctypes.byref(hd), 0) == 0
    handles[name] = hd

dev = torch.device("cuda:0")
NAME = {torch.float64: "fp64", torch.float32: "fp32", torch.float16: "fp16"}
h = ab.Handle(0)
n = 16384
import os
lda = n + int(os.environ.get("LDA_PAD", "0"))
stream = torch.cuda.current_stream().cuda_stream


def timed(lib, hd, ar, st, uplo, diag, T, x):
    best = 1e9
    for _ in range(6):
        torch.cuda._sleep(40_000)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.accblas_trsv(hd, ab.dtype_code(ar), ab.dtype_code(st), uplo, diag, n, T.data_ptr(), lda, x.data_ptr(),
                              1, stream)
        e1.record()
        e1.synchronize()
        assert rc == 0
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3


for st in (torch.float32, torch.float64, torch.float16):
    T = torch.empty(n * lda, dtype=st, device=dev)
    b = torch.empty(n, dtype=st, device=dev)
    h.fill_uniform(n, lda, T, lda, 42, 0)
    h.fill_uniform(n, 1, b, 1, 42, n * n)
    T.mul_(0.01)
    T.view(n, lda).diagonal().fill_(1.0)
    for ar in (torch.float64, torch.float32):
        for uplo, diag in ((ab.LOWER, ab.UNIT), (ab.UPPER, ab.NON_UNIT)):
            variants = [("old", None, None), ("new", 1, -1), ("sub32", 0, -1), ("g0", 1, 0), ("g1024", 1, 1024)]
            res = {}
            for rep in range(2):
                for name, whole, ahead in variants:
                    x = b.clone()
                    if whole is not None:
                        ab.tune("trsv_whole_block_spin", whole)
                        ab.tune("trsv_l2_ahead", ahead)
                    if whole is not None:
                        # sub32 (32 entries at a time) exists in the single-CTA kernel only
                        ab.tune("trsv_variant", 1 if name == "sub32" else -1)
                    lib_name = "old" if name == "old" else "new"
                    t = timed(libs[lib_name], handles[lib_name], ar, st, uplo, diag, T, x)
                    res[name] = min(res.get(name, 1e9), t)
            sols = {}
            for lib_name in ("old", "new"):
                x = b.clone()
                timed(libs[lib_name], handles[lib_name], ar, st, uplo, diag, T, x)
                sols[lib_name] = x
            same = bool(torch.equal(sols["old"], sols["new"]))
            print(f"trsv Acc<{NAME[ar]},{NAME[st]}> {'lower' if uplo == ab.LOWER else 'upper'}/"
                  f"{'unit' if diag == ab.UNIT else 'nonunit'}: " +
                  "  ".join(f"{name} {res[name]:6.1f}" for name, _, _ in variants) +
                  f"  bit-identical solutions: {same}", flush=True)
    del T
