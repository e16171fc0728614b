"""ctypes binding of libaccblas_b200.so (the C ABI declared in include/accblas.h).

There is deliberately no fallback: if the shared library is missing or a call
fails, an exception is raised.  Nothing in this module computes anything on the
CPU.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_int, c_int64, c_size_t,
                    c_uint32, c_uint64, c_void_p)
from pathlib import Path

HERE = Path(__file__).resolve().parent
# ACCBLAS_LIB: development tools load libaccblas_b200_dev.so (same sources plus
# the timeline probes) through the same binding
LIB_PATH = Path(os.environ.get("ACCBLAS_LIB", HERE / "libaccblas_b200.so"))
BASELINES_PATH = HERE / "libaccblas_baselines.so"

F64, F32, F16 = 0, 1, 2
UPPER, LOWER = 0, 1
NON_UNIT, UNIT = 0, 1
ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_ALLOC, ERR_DATA, ERR_PEER = 1, 2, 3, 4, 5, 6

# every symbol include/accblas.h declares: (name, restype, argtypes)
_P = c_void_p
SYMBOLS = {
    "accblas_version": (c_int, []),
    "accblas_status_string": (c_char_p, [c_int]),
    "accblas_last_error": (c_char_p, []),
    "accblas_sizeof": (c_size_t, [c_int]),
    "accblas_create": (c_int, [POINTER(_P), c_int]),
    "accblas_destroy": (c_int, [_P]),
    "accblas_get_sm_count": (c_int, [_P, POINTER(c_int)]),
    "accblas_gemv": (c_int, [_P, c_int, c_int, c_int64, c_int64, c_double, _P,
                             c_int64, _P, c_int64, c_double, _P, c_int64, _P]),
    "accblas_dot": (c_int, [_P, c_int, c_int, c_int, c_int64, _P, c_int64, _P,
                            c_int64, _P, _P]),
    "accblas_trsv": (c_int, [_P, c_int, c_int, c_int, c_int, c_int64, _P,
                             c_int64, _P, c_int64, _P]),
    "accblas_convert": (c_int, [_P, c_int, c_int, c_int64, c_int64, _P,
                                c_int64, _P, c_int64, _P]),
    "accblas_fill_uniform": (c_int, [_P, c_int, c_int64, c_int64, _P, c_int64,
                                     c_uint32, c_uint64, _P]),
    "accblas_l1_error": (c_int, [_P, c_int, c_int, c_int64, _P, c_int64, _P,
                                 c_int64, _P, _P]),
    "accblas_gemv_host": (c_int, [_P, c_int, c_int, c_int64, c_int64, c_double,
                                  _P, c_int64, _P, c_int64, c_double, _P,
                                  c_int64, _P]),
    "accblas_dot_host": (c_int, [_P, c_int, c_int, c_int, c_int64, _P, c_int64,
                                 _P, c_int64, _P, _P]),
    "accblas_trsv_host": (c_int, [_P, c_int, c_int, c_int, c_int, c_int64, _P,
                                  c_int64, _P, c_int64, _P]),
    "accblas_tune": (c_int, [c_char_p, c_int]),
    "accblas_peer_export": (c_int, [_P, _P]),
    "accblas_peer_connect_ipc": (c_int, [_P, c_int, c_int, _P]),
    "accblas_peer_mailbox": (c_int, [_P, POINTER(_P)]),
    "accblas_peer_connect_ptrs": (c_int, [_P, c_int, c_int, POINTER(_P), POINTER(c_int)]),
    "accblas_dot_allreduce": (c_int, [_P, c_int, c_int, c_int, c_int64, _P, c_int64, _P,
                                      c_int64, _P, _P]),
    "accblas_peer_disconnect": (c_int, [_P]),
    "accblas_peer_set_timeout": (c_int, [_P, c_double]),
    "accblas_peer_status": (c_int, [_P, POINTER(c_uint64)]),
}

BASELINE_SYMBOLS = {
    "accblas_baseline_cublas_gemv": (c_int, [c_int, c_int64, c_int64, c_double,
                                             _P, c_int64, _P, c_int64, c_double,
                                             _P, c_int64, _P]),
    "accblas_baseline_cublas_dot": (c_int, [c_int, c_int64, _P, c_int64, _P,
                                            c_int64, _P, _P]),
    "accblas_baseline_cublas_trsv": (c_int, [c_int, c_int, c_int, c_int64, _P,
                                             c_int64, _P, c_int64, _P]),
}


class AccblasError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"accblas status {status}: {message}")
        self.status = status


_lib = None
_baselines = None


def load() -> ctypes.CDLL:
    """Loads libaccblas_b200.so; raises (never falls back) if it is absent."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback for the accblas kernels.")
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def load_baselines() -> ctypes.CDLL:
    global _baselines
    if _baselines is None:
        if not BASELINES_PATH.exists():
            raise ImportError(f"{BASELINES_PATH} is missing (run build())")
        lib = ctypes.CDLL(str(BASELINES_PATH))
        for name, (res, args) in BASELINE_SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _baselines = lib
    return _baselines


def check(status: int) -> None:
    if status != 0:
        lib = load()
        detail = lib.accblas_last_error().decode() or \
            lib.accblas_status_string(status).decode()
        raise AccblasError(status, detail)
