"""accessor-BLAS for B200 -- host-side mirror of the reference launcher API.

The reference's drop-in surface is a set of C++ host templates
(`acc_gemv<Ar,St>`, `acc_dot<Ar,St,Res>`, `acc_trsv<Ar,St>` and their
plain-pointer twins; /root/reference/cuda/{gemv,dot,trsv}_kernels.cuh).  The
C++ mirror of those templates lives in include/accblas/*.cuh; this Python
package mirrors the same entry points over torch CUDA tensors so the tests and
bench.py read like the reference's drivers.  All arithmetic happens in
libaccblas_b200.so (hand-written sm_100a kernels behind the C ABI of
include/accblas.h); torch only provides device memory and streams.

Import name: `accessor_blas_b200` (see accessor_blas_b200.py at the repo root;
the directory name carries a hyphen).
"""
from __future__ import annotations

import ctypes
from typing import NamedTuple, Optional, Tuple

import numpy as np
import torch

from . import capi
from .capi import (AccblasError, F16, F32, F64, LOWER, NON_UNIT, UNIT, UPPER)

__all__ = [
    "MatrixInfo", "Handle", "AccblasError", "dtype_code", "torch_dtype",
    "gemv", "acc_gemv", "dot", "acc_dot", "trsv", "acc_trsv", "convert",
    "fill_uniform", "l1_error", "tune", "UPPER", "LOWER", "UNIT", "NON_UNIT",
    "F64", "F32", "F16",
]

_TORCH_TO_CODE = {torch.float64: F64, torch.float32: F32, torch.float16: F16}
_CODE_TO_TORCH = {v: k for k, v in _TORCH_TO_CODE.items()}
_NUMPY_TO_CODE = {np.dtype(np.float64): F64, np.dtype(np.float32): F32,
                  np.dtype(np.float16): F16}


def dtype_code(dtype) -> int:
    if isinstance(dtype, int):
        return dtype
    if isinstance(dtype, torch.dtype):
        return _TORCH_TO_CODE[dtype]
    return _NUMPY_TO_CODE[np.dtype(dtype)]


def torch_dtype(code: int) -> torch.dtype:
    return _CODE_TO_TORCH[code]


class MatrixInfo(NamedTuple):
    """`matrix_info` of the reference (cuda/utils.cuh:18-56): 2-D size and the
    row stride in elements."""
    size: Tuple[int, int]
    stride: int

    @staticmethod
    def make(size, stride: Optional[int] = None) -> "MatrixInfo":
        rows, cols = int(size[0]), int(size[1])
        return MatrixInfo((rows, cols), int(cols if stride is None else stride))

    def get_1d_size(self) -> int:
        return self.size[0] * self.stride

    def get_num_elems(self) -> int:
        return self.size[0] * self.size[1]


def _stream_ptr(stream, device=None) -> int:
    """cudaStream_t of `stream`; None = the current stream of `device` (the
    handle's device, which need not be the current one)."""
    if stream is None:
        stream = torch.cuda.current_stream(device)
    return int(stream.cuda_stream) if hasattr(stream, "cuda_stream") else int(stream)


def _dev_ptr(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise ValueError("accblas device entry points need CUDA tensors")
    return t.data_ptr()


class Handle:
    """Device handle (the role `myBlasHandle` plays in the reference,
    cuda/dot_kernels.cuh:29-65, extended to all three operations)."""

    def __init__(self, device: Optional[int] = None):
        self._lib = capi.load()
        self._h = ctypes.c_void_p()
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else -1
        capi.check(self._lib.accblas_create(ctypes.byref(self._h), int(device)))
        self.device = int(device) if int(device) >= 0 else torch.cuda.current_device()

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.accblas_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sm_count(self) -> int:
        out = ctypes.c_int()
        capi.check(self._lib.accblas_get_sm_count(self._h, ctypes.byref(out)))
        return out.value

    # -- device-pointer entry points ---------------------------------------
    def gemv(self, ar, m: int, n: int, alpha: float, A: torch.Tensor, lda: int,
             x: torch.Tensor, incx: int, beta: float, y: torch.Tensor,
             incy: int, stream=None) -> None:
        st = dtype_code(A.dtype)
        assert x.dtype == A.dtype and y.dtype == A.dtype
        capi.check(self._lib.accblas_gemv(
            self._h, dtype_code(ar), st, m, n, float(alpha), _dev_ptr(A), lda,
            _dev_ptr(x), incx, float(beta), _dev_ptr(y), incy,
            _stream_ptr(stream, self.device)))

    def dot(self, ar, n: int, x: torch.Tensor, incx: int, y: torch.Tensor,
            incy: int, result: torch.Tensor, stream=None) -> None:
        assert x.dtype == y.dtype
        capi.check(self._lib.accblas_dot(
            self._h, dtype_code(ar), dtype_code(x.dtype),
            dtype_code(result.dtype), n, _dev_ptr(x), incx, _dev_ptr(y), incy,
            _dev_ptr(result), _stream_ptr(stream, self.device)))

    # ---- multi-GPU DOT with the all-reduce inside the kernel (peer memory)
    def peer_export(self) -> bytes:
        """64-byte CUDA IPC handle of this GPU's mailbox."""
        import ctypes
        buf = ctypes.create_string_buffer(64)
        capi.check(self._lib.accblas_peer_export(self._h, buf))
        return buf.raw

    def peer_connect_ipc(self, world: int, rank: int, handles: bytes) -> None:
        assert len(handles) == 64 * world
        capi.check(self._lib.accblas_peer_connect_ipc(self._h, world, rank, handles))

    def peer_mailbox(self) -> int:
        import ctypes
        ptr = ctypes.c_void_p()
        capi.check(self._lib.accblas_peer_mailbox(self._h, ctypes.byref(ptr)))
        return ptr.value

    def peer_connect_ptrs(self, world: int, rank: int, mailboxes, devices) -> None:
        import ctypes
        ptrs = (ctypes.c_void_p * world)(*mailboxes)
        devs = (ctypes.c_int * world)(*devices)
        capi.check(self._lib.accblas_peer_connect_ptrs(self._h, world, rank, ptrs, devs))

    def peer_disconnect(self) -> None:
        capi.check(self._lib.accblas_peer_disconnect(self._h))
        self._peer_group = None

    def peer_set_timeout(self, seconds: float) -> None:
        capi.check(self._lib.accblas_peer_set_timeout(self._h, float(seconds)))

    def peer_status(self) -> int:
        """0, or the number of the dot_allreduce call that timed out."""
        failed = ctypes.c_uint64(0)
        self._lib.accblas_peer_status(self._h, ctypes.byref(failed))
        return int(failed.value)

    def dot_allreduce(self, ar, n: int, x: torch.Tensor, incx: int, y: torch.Tensor,
                      incy: int, result: torch.Tensor, stream=None) -> None:
        assert x.dtype == y.dtype
        capi.check(self._lib.accblas_dot_allreduce(
            self._h, dtype_code(ar), dtype_code(x.dtype),
            dtype_code(result.dtype), n, _dev_ptr(x), incx, _dev_ptr(y), incy,
            _dev_ptr(result), _stream_ptr(stream, self.device)))

    def trsv(self, ar, uplo: int, diag: int, n: int, A: torch.Tensor, lda: int,
             x: torch.Tensor, incx: int, stream=None) -> None:
        assert x.dtype == A.dtype
        capi.check(self._lib.accblas_trsv(
            self._h, dtype_code(ar), dtype_code(A.dtype), uplo, diag, n,
            _dev_ptr(A), lda, _dev_ptr(x), incx, _stream_ptr(stream, self.device)))

    def convert(self, rows: int, cols: int, src: torch.Tensor, ld_in: int,
                dst: torch.Tensor, ld_out: int, stream=None) -> None:
        capi.check(self._lib.accblas_convert(
            self._h, dtype_code(dst.dtype), dtype_code(src.dtype), rows, cols,
            _dev_ptr(src), ld_in, _dev_ptr(dst), ld_out, _stream_ptr(stream, self.device)))

    def fill_uniform(self, rows: int, cols: int, out: torch.Tensor, ld: int,
                     seed: int = 42, first_draw: int = 0, stream=None) -> None:
        capi.check(self._lib.accblas_fill_uniform(
            self._h, dtype_code(out.dtype), rows, cols, _dev_ptr(out), ld,
            seed, first_draw, _stream_ptr(stream, self.device)))

    def l1_error(self, n: int, ref: torch.Tensor, inc_ref: int,
                 res: torch.Tensor, inc_res: int, stream=None) -> float:
        """sum|ref-res| / sum|ref| -- the reference's error metric."""
        out = torch.empty(2, dtype=torch.float64, device=ref.device)
        capi.check(self._lib.accblas_l1_error(
            self._h, dtype_code(ref.dtype), dtype_code(res.dtype), n,
            _dev_ptr(ref), inc_ref, _dev_ptr(res), inc_res, _dev_ptr(out),
            _stream_ptr(stream, self.device)))
        d, s = out.tolist()
        return d / s if s != 0.0 else float("inf") if d != 0.0 else 0.0

    # -- host-buffer entry points (numpy arrays; staged through the device) --
    def gemv_host(self, ar, m: int, n: int, alpha: float, A: np.ndarray,
                  lda: int, x: np.ndarray, incx: int, beta: float,
                  y: np.ndarray, incy: int, stream=None) -> None:
        st = dtype_code(A.dtype)
        capi.check(self._lib.accblas_gemv_host(
            self._h, dtype_code(ar), st, m, n, float(alpha), A.ctypes.data, lda,
            x.ctypes.data, incx, float(beta), y.ctypes.data, incy,
            _stream_ptr(stream, self.device)))

    def dot_host(self, ar, n: int, x: np.ndarray, incx: int, y: np.ndarray,
                 incy: int, result: np.ndarray, stream=None) -> None:
        capi.check(self._lib.accblas_dot_host(
            self._h, dtype_code(ar), dtype_code(x.dtype),
            dtype_code(result.dtype), n, x.ctypes.data, incx, y.ctypes.data,
            incy, result.ctypes.data, _stream_ptr(stream, self.device)))

    def trsv_host(self, ar, uplo: int, diag: int, n: int, A: np.ndarray,
                  lda: int, x: np.ndarray, incx: int, stream=None) -> None:
        capi.check(self._lib.accblas_trsv_host(
            self._h, dtype_code(ar), dtype_code(A.dtype), uplo, diag, n,
            A.ctypes.data, lda, x.ctypes.data, incx, _stream_ptr(stream, self.device)))


_default_handles = {}


def default_handle() -> Handle:
    dev = torch.cuda.current_device()
    if dev not in _default_handles:
        _default_handles[dev] = Handle(dev)
    return _default_handles[dev]


# ---------------------------------------------------------------------------
# Reference-shaped launchers (argument order and meaning as in the reference)
# ---------------------------------------------------------------------------
def acc_gemv(ar, m_info: MatrixInfo, alpha: float, mtx: torch.Tensor,
             x_info: MatrixInfo, x: torch.Tensor, res_info: MatrixInfo,
             beta: float, res: torch.Tensor, handle: Optional[Handle] = None,
             stream=None) -> None:
    """res = alpha * mtx * x + beta * res (cuda/gemv_kernels.cuh:168-193)."""
    h = handle or default_handle()
    h.gemv(ar, m_info.size[0], m_info.size[1], alpha, mtx, m_info.stride, x,
           x_info.stride, beta, res, res_info.stride, stream)


def gemv(m_info, alpha, mtx, x_info, x, res_info, beta, res, handle=None,
         stream=None) -> None:
    """Plain-pointer variant: arithmetic type == storage type
    (cuda/gemv_kernels.cuh:136-147)."""
    acc_gemv(mtx.dtype, m_info, alpha, mtx, x_info, x, res_info, beta, res,
             handle, stream)


def acc_dot(ar, x_info: MatrixInfo, x: torch.Tensor, y_info: MatrixInfo,
            y: torch.Tensor, res: torch.Tensor,
            handle: Optional[Handle] = None, stream=None) -> None:
    """*res = x . y, accumulated in `ar`, stored as res.dtype
    (cuda/dot_kernels.cuh:224-263).  `res` is a one-element device tensor."""
    h = handle or default_handle()
    h.dot(ar, x_info.size[0], x, x_info.stride, y, y_info.stride, res, stream)


def dot(x_info, x, y_info, y, res, handle=None, stream=None) -> None:
    """Plain-pointer variant (cuda/dot_kernels.cuh:192-206)."""
    acc_dot(x.dtype, x_info, x, y_info, y, res, handle, stream)


def acc_trsv(ar, m_info: MatrixInfo, ttype: int, dtype: int,
             mtx: torch.Tensor, x_info: MatrixInfo, x: torch.Tensor,
             handle: Optional[Handle] = None, stream=None) -> None:
    """In-place triangular solve (cuda/trsv_kernels.cuh:918-961); `ttype` is
    UPPER/LOWER, `dtype` UNIT/NON_UNIT."""
    h = handle or default_handle()
    h.trsv(ar, ttype, dtype, m_info.size[0], mtx, m_info.stride, x,
           x_info.stride, stream)


def trsv(m_info, ttype, dtype, mtx, x_info, x, handle=None, stream=None) -> None:
    """Plain-pointer variant (cuda/trsv_kernels.cuh:455-488)."""
    acc_trsv(mtx.dtype, m_info, ttype, dtype, mtx, x_info, x, handle, stream)


def convert(info: MatrixInfo, src: torch.Tensor, dst: torch.Tensor,
            handle: Optional[Handle] = None, stream=None) -> None:
    """dst = static_cast<dst.dtype>(src) over a strided 2-D view
    (cuda/matrix_helper.cuh:93-103)."""
    h = handle or default_handle()
    h.convert(info.size[0], info.size[1], src, info.stride, dst, info.stride,
              stream)


def fill_uniform(info: MatrixInfo, out: torch.Tensor, seed: int = 42,
                 first_draw: int = 0, handle: Optional[Handle] = None,
                 stream=None) -> None:
    """Device-side gen_mtx / write_random (cuda/matrix_helper.cuh:28-75) with
    uniform(-1,1) from std::default_random_engine(seed)."""
    h = handle or default_handle()
    h.fill_uniform(info.size[0], info.size[1], out, info.stride, seed,
                   first_draw, stream)


def l1_error(ref: torch.Tensor, res: torch.Tensor,
             handle: Optional[Handle] = None) -> float:
    h = handle or default_handle()
    return h.l1_error(ref.numel(), ref, 1, res, 1)


def tune(key: str, value: int) -> None:
    capi.check(capi.load().accblas_tune(key.encode(), int(value)))
