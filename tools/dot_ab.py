"""Same-box A/B of two builds of the library on DOT / GEMV (development tool).

    python tools/dot_ab.py [old.so]
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
    lib.accblas_dot.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, P, ctypes.c_int64,
                                P, ctypes.c_int64, P, P]
    lib.accblas_gemv.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                 P, ctypes.c_int64, P, ctypes.c_int64, ctypes.c_double, P, ctypes.c_int64, P]
    hd = P()
    assert lib.accblas_create(ctypes.byref(hd), 0) == 0
    handles[name] = hd

dev = torch.device("cuda:0")
NAME = {torch.float64: "fp64", torch.float32: "fp32", torch.float16: "fp16"}
h = ab.Handle(0)
stream = torch.cuda.current_stream().cuda_stream


def best_of(fn, reps=10):
    best = 1e9
    fn()
    torch.cuda.synchronize()
    for _ in range(reps):
        torch.cuda._sleep(40_000)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        assert fn() == 0
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3


n = 2 ** 28
m = k = 16384
for st in (torch.float32, torch.float16, torch.float64):
    x = torch.empty(n, dtype=st, device=dev)
    y = torch.empty(n, dtype=st, device=dev)
    h.fill_uniform(1, n, x, n, 42, 0)
    h.fill_uniform(1, n, y, n, 42, n)
    yv = torch.zeros(m, dtype=st, device=dev)
    for ar in (torch.float64, torch.float32):
        res = torch.zeros(1, dtype=ar, device=dev)
        out = {}
        for rep in range(2):
            for name in ("old", "new"):
                lib, hd = libs[name], handles[name]
                t = best_of(lambda: lib.accblas_dot(hd, ab.dtype_code(ar), ab.dtype_code(st), ab.dtype_code(ar), n,
                                                    x.data_ptr(), 1, y.data_ptr(), 1, res.data_ptr(), stream))
                out["dot " + name] = min(out.get("dot " + name, 1e9), t)
                # the first m*k elements of x as the matrix, y[:k] as the vector
                t = best_of(lambda: lib.accblas_gemv(hd, ab.dtype_code(ar), ab.dtype_code(st), m, k, 1.0,
                                                     x.data_ptr(), k, y.data_ptr(), 1, 0.0, yv.data_ptr(), 1,
                                                     stream))
                out["gemv " + name] = min(out.get("gemv " + name, 1e9), t)
        print(f"Acc<{NAME[ar]},{NAME[st]}>: DOT 2^28 old {out['dot old']:7.1f} us new {out['dot new']:7.1f} us | "
              f"GEMV 16384^2 old {out['gemv old']:7.1f} us new {out['gemv new']:7.1f} us", flush=True)
    del x, y
