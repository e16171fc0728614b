"""ctypes bindings of the CHECKERS (test infrastructure only):

  Oracle       oracle/liboracle.so        CPU restatement of the reference path
  RefKernels   oracle/_ref/libref_kernels.so   the reference's own CUDA kernels
                                               (needs a GPU to call)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import
this module.
"""
from __future__ import annotations

import ctypes
from ctypes import c_double, c_int, c_int64, c_uint32, c_uint64, c_void_p
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_LIB = ROOT / "oracle" / "liboracle.so"
REF_LIB = ROOT / "oracle" / "_ref" / "libref_kernels.so"

F64, F32, F16 = 0, 1, 2
_CODES = {np.dtype(np.float64): F64, np.dtype(np.float32): F32,
          np.dtype(np.float16): F16}
_NP = {F64: np.float64, F32: np.float32, F16: np.float16}


def code(dtype) -> int:
    if isinstance(dtype, int):
        return dtype
    try:
        return _CODES[np.dtype(dtype)]
    except TypeError:
        import torch
        return {torch.float64: F64, torch.float32: F32, torch.float16: F16}[dtype]


def np_dtype(c: int):
    return _NP[c]


def _ptr(a: np.ndarray) -> int:
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


class Oracle:
    def __init__(self):
        if not ORACLE_LIB.exists():
            raise ImportError(f"{ORACLE_LIB} missing: run `make -C oracle`")
        L = ctypes.CDLL(str(ORACLE_LIB))
        P, I, L64, D = c_void_p, c_int, c_int64, c_double
        L.oracle_num_threads.restype = I
        L.oracle_set_num_threads.argtypes = [I]
        L.oracle_uniform_stdlib.argtypes = [P, L64, c_uint32, c_uint64]
        L.oracle_uniform_closed_form.argtypes = [P, L64, c_uint32, c_uint64]
        L.oracle_convert.argtypes = [I, I, L64, P, P]
        gemv_args = [I, I, L64, L64, D, P, L64, P, L64, D, P, L64]
        L.oracle_ref_gemv.argtypes = gemv_args
        L.oracle_cpu_gemv.argtypes = gemv_args
        L.oracle_ref_dot.argtypes = [I, I, I, L64, P, L64, P, L64, I, P, P]
        L.oracle_cpu_dot.argtypes = [I, I, I, L64, P, L64, P, L64, P]
        trsv_args = [I, I, I, I, L64, P, L64, P, L64]
        L.oracle_ref_trsv.argtypes = trsv_args
        L.oracle_cpu_trsv.argtypes = trsv_args
        L.oracle_exact_gemv.argtypes = [I, L64, L64, D, P, L64, P, L64, D, P, L64, P]
        L.oracle_exact_dot.argtypes = [I, L64, P, L64, P, L64, P]
        L.oracle_exact_trsv.argtypes = [I, I, I, L64, P, L64, P, L64, P]
        L.oracle_l1_rel_error.argtypes = [I, L64, P, P, L64]
        L.oracle_l1_rel_error.restype = D
        self.L = L

    # -- input stream -------------------------------------------------------
    def uniform(self, count: int, seed: int = 42, first_draw: int = 0,
                closed_form: bool = True) -> np.ndarray:
        out = np.empty(count, dtype=np.float64)
        fn = self.L.oracle_uniform_closed_form if closed_form \
            else self.L.oracle_uniform_stdlib
        fn(_ptr(out), count, seed, first_draw)
        return out

    def convert(self, a: np.ndarray, dst) -> np.ndarray:
        a = np.ascontiguousarray(a)
        out = np.empty(a.shape, dtype=np_dtype(code(dst)))
        rc = self.L.oracle_convert(code(dst), code(a.dtype), a.size, _ptr(a), _ptr(out))
        assert rc == 0
        return out

    # -- order-faithful restatements of the reference kernels ---------------
    def ref_gemv(self, ar, A, m, n, lda, x, alpha, beta, y, incx=1, incy=1):
        y = np.array(y, copy=True)
        rc = self.L.oracle_ref_gemv(code(ar), code(A.dtype), m, n, alpha, _ptr(A),
                                    lda, _ptr(x), incx, beta, _ptr(y), incy)
        assert rc == 0
        return y

    def ref_dot(self, ar, x, y, res, blocks, n=None, incx=1, incy=1):
        n = x.size // incx if n is None else n
        out = np.zeros(1, dtype=np_dtype(code(res)))
        partials = np.zeros(blocks, dtype=np_dtype(code(ar)))
        rc = self.L.oracle_ref_dot(code(ar), code(x.dtype), code(res), n, _ptr(x),
                                   incx, _ptr(y), incy, blocks, _ptr(out),
                                   _ptr(partials))
        assert rc == 0
        return out[0], partials

    def ref_trsv(self, ar, A, n, lda, b, upper, unit, incx=1):
        x = np.array(b, copy=True)
        rc = self.L.oracle_ref_trsv(code(ar), code(A.dtype), int(upper), int(unit),
                                    n, _ptr(A), lda, _ptr(x), incx)
        assert rc == 0
        return x

    # -- accuracy oracle ------------------------------------------------------
    def exact_gemv(self, A, m, n, lda, x, alpha, beta, y, incx=1, incy=1):
        out = np.empty(m, dtype=np.float64)
        rc = self.L.oracle_exact_gemv(code(A.dtype), m, n, alpha, _ptr(A), lda,
                                      _ptr(x), incx, beta, _ptr(y), incy, _ptr(out))
        assert rc == 0
        return out

    def exact_dot(self, x, y, n=None, incx=1, incy=1) -> float:
        n = x.size // incx if n is None else n
        out = np.empty(1, dtype=np.float64)
        rc = self.L.oracle_exact_dot(code(x.dtype), n, _ptr(x), incx, _ptr(y), incy,
                                     _ptr(out))
        assert rc == 0
        return float(out[0])

    def exact_trsv(self, A, n, lda, b, upper, unit, incb=1):
        out = np.empty(n, dtype=np.float64)
        rc = self.L.oracle_exact_trsv(code(A.dtype), int(upper), int(unit), n,
                                      _ptr(A), lda, _ptr(b), incb, _ptr(out))
        assert rc == 0
        return out

    def l1_rel_error(self, ref: np.ndarray, res: np.ndarray, inc_res=1) -> float:
        ref = np.ascontiguousarray(ref, dtype=np.float64)
        n = ref.size
        return float(self.L.oracle_l1_rel_error(code(res.dtype), n, _ptr(ref),
                                                _ptr(res), inc_res))

    # -- CPU baseline ("port") ---------------------------------------------------
    def cpu_gemv(self, ar, A, m, n, lda, x, alpha, beta, y, incx=1, incy=1):
        rc = self.L.oracle_cpu_gemv(code(ar), code(A.dtype), m, n, alpha, _ptr(A),
                                    lda, _ptr(x), incx, beta, _ptr(y), incy)
        assert rc == 0
        return y

    def cpu_dot(self, ar, x, y, res, n=None, incx=1, incy=1):
        n = x.size // incx if n is None else n
        out = np.zeros(1, dtype=np_dtype(code(res)))
        rc = self.L.oracle_cpu_dot(code(ar), code(x.dtype), code(res), n, _ptr(x),
                                   incx, _ptr(y), incy, _ptr(out))
        assert rc == 0
        return out[0]

    def cpu_trsv(self, ar, A, n, lda, x, upper, unit, incx=1):
        rc = self.L.oracle_cpu_trsv(code(ar), code(A.dtype), int(upper), int(unit),
                                    n, _ptr(A), lda, _ptr(x), incx)
        assert rc == 0
        return x

    @property
    def num_threads(self) -> int:
        return int(self.L.oracle_num_threads())


class RefKernels:
    """The reference's own CUDA launchers (device pointers, default stream)."""

    def __init__(self):
        if not REF_LIB.exists():
            raise ImportError(f"{REF_LIB} missing: run `make -C oracle ref` "
                              "where /root/reference exists")
        L = ctypes.CDLL(str(REF_LIB))
        P, I, L64, D = c_void_p, c_int, c_int64, c_double
        L.ref_gemv.argtypes = [I, I, I, L64, L64, D, P, L64, P, L64, D, P, L64]
        L.ref_dot.argtypes = [I, I, I, I, L64, P, L64, P, L64, P]
        L.ref_trsv.argtypes = [I, I, I, I, I, L64, P, L64, P, L64, P]
        L.ref_cublas_gemv.argtypes = [I, L64, L64, D, P, L64, P, L64, D, P, L64]
        L.ref_cublas_dot.argtypes = [I, L64, P, L64, P, L64, P]
        L.ref_cublas_trsv.argtypes = [I, I, I, L64, P, L64, P, L64]
        self.L = L
        self._helper = None

    def sm_count(self) -> int:
        return int(self.L.ref_sm_count())

    def sync(self):
        assert self.L.ref_sync() == 0

    def gemv(self, ar, m, n, alpha, A, lda, x, incx, beta, y, incy, plain=False):
        rc = self.L.ref_gemv(code(ar), code(A.dtype), int(plain), m, n, alpha,
                             A.data_ptr(), lda, x.data_ptr(), incx, beta,
                             y.data_ptr(), incy)
        assert rc == 0

    def dot(self, ar, n, x, incx, y, incy, result, plain=False):
        rc = self.L.ref_dot(code(ar), code(x.dtype), code(result.dtype), int(plain),
                            n, x.data_ptr(), incx, y.data_ptr(), incy,
                            result.data_ptr())
        assert rc == 0

    def trsv(self, ar, upper, unit, n, A, lda, x, incx, plain=False):
        import torch
        if self._helper is None:
            self._helper = torch.zeros(2, dtype=torch.int32, device=A.device)
        rc = self.L.ref_trsv(code(ar), code(A.dtype), int(plain), int(upper),
                             int(unit), n, A.data_ptr(), lda, x.data_ptr(), incx,
                             self._helper.data_ptr())
        assert rc == 0

    def cublas_gemv(self, m, n, alpha, A, lda, x, incx, beta, y, incy):
        rc = self.L.ref_cublas_gemv(code(A.dtype), m, n, alpha, A.data_ptr(), lda,
                                    x.data_ptr(), incx, beta, y.data_ptr(), incy)
        assert rc == 0

    def cublas_dot(self, n, x, incx, y, incy, result):
        rc = self.L.ref_cublas_dot(code(x.dtype), n, x.data_ptr(), incx,
                                   y.data_ptr(), incy, result.data_ptr())
        assert rc == 0

    def cublas_trsv(self, upper, unit, n, A, lda, x, incx):
        rc = self.L.ref_cublas_trsv(code(A.dtype), int(upper), int(unit), n,
                                    A.data_ptr(), lda, x.data_ptr(), incx)
        assert rc == 0
