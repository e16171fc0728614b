// Drop-in replacement for the reference's gemv_kernels.cuh
// (/root/reference/cuda/gemv_kernels.cuh): same names, template parameters and
// argument meaning.
//
//   host launchers  gemv<T>, acc_gemv<Ar, St>     -> accblas_gemv (C ABI,
//                   hand-written sm_100a streaming kernel in
//                   libaccblas_b200.so)
//   kernel::gemv<block_size, T>,
//   kernel::acc_gemv<block_size>(alpha, mtx, x, beta, res)
//                   accessor-GENERIC __global__ kernels with the signatures the
//                   reference's README documents (README.md:31-50).  They work
//                   for any range/accessor type (that is what the accessor
//                   abstraction is for); the launchers above do not use them
//                   for reduced_row_major operands because the C-ABI kernel
//                   is specialised for that layout.
//   cublas_gemv     vendor baseline, called as the reference calls it.
#pragma once

#include <cinttypes>

#include <cublas_v2.h>

#include <accessor/range.hpp>
#include <accessor/reduced_row_major.hpp>
#include <accessor/row_major.hpp>

#include "kernel_utils.cuh"
#include "utils.cuh"

namespace kernel {

// res = alpha * mtx * x + beta * res, one block per row, plain pointers.
template <std::int64_t block_size, typename ValueType>
__global__ __launch_bounds__(block_size) void gemv(
    const matrix_info m_info, ValueType alpha,
    const ValueType* __restrict__ mtx, const matrix_info x_info,
    const ValueType* __restrict__ x, const matrix_info res_info,
    ValueType beta, ValueType* __restrict__ res)
{
    const std::int64_t row = blockIdx.x;
    if (row >= m_info.size[0]) {
        return;
    }
    const ValueType* a = mtx + row * m_info.stride;
    ValueType acc{};
    for (std::int64_t col = threadIdx.x; col < m_info.size[1];
         col += block_size) {
        acc += a[col] * x[col * x_info.stride];
    }
    const ValueType total = detail::block_total<block_size>(acc);
    if (threadIdx.x == 0) {
        const auto idx = row * res_info.stride;
        res[idx] = (beta == ValueType{0}) ? alpha * total
                                          : alpha * total + beta * res[idx];
    }
}

// The same through ranges: every element access goes through the accessor.
template <std::int64_t block_size, typename MtxRange, typename XRange,
          typename ResRange, typename ArType>
__global__ __launch_bounds__(block_size) void acc_gemv(ArType alpha,
                                                       MtxRange mtx, XRange x,
                                                       ArType beta,
                                                       ResRange res)
{
    using ar_type = decltype(alpha * mtx(0, 0) * x(0, 0) + beta * res(0, 0));
    static_assert(std::is_same<ArType, ar_type>::value, "Types must be equal!");
    const std::int64_t row = blockIdx.x;
    if (row >= mtx.length(0)) {
        return;
    }
    const std::int64_t num_cols = mtx.length(1);
    ar_type acc{};
    for (std::int64_t col = threadIdx.x; col < num_cols; col += block_size) {
        acc += mtx(row, col) * x(col, 0);
    }
    const ar_type total = detail::block_total<block_size>(acc);
    if (threadIdx.x == 0) {
        if (beta == ArType{0}) {
            res(row, 0) = alpha * total;
        } else {
            res(row, 0) = alpha * total + beta * res(row, 0);
        }
    }
}

}  // namespace kernel


// res = alpha * mtx * x + beta * res without the accessor: arithmetic type ==
// storage type (cuda/gemv_kernels.cuh:136-147).
template <typename ValueType>
void gemv(const matrix_info m_info, ValueType alpha, const ValueType* mtx,
          const matrix_info x_info, const ValueType* x,
          const matrix_info res_info, ValueType beta, ValueType* res)
{
    constexpr accblas_dtype t = accblas_detail::dtype_of<ValueType>::value;
    ACCBLAS_CALL(accblas_gemv(accblas_detail::default_handle(), t, t,
                              m_info.size[0], m_info.size[1],
                              static_cast<double>(alpha), mtx, m_info.stride, x,
                              x_info.stride, static_cast<double>(beta), res,
                              res_info.stride, nullptr));
}

// res = alpha * mtx * x + beta * res computed in ArType on StType storage
// (cuda/gemv_kernels.cuh:168-193).  Launches on the default stream, as the
// reference does.
template <typename ArType, typename StType>
void acc_gemv(const matrix_info m_info, ArType alpha, const StType* mtx,
              const matrix_info x_info, const StType* x,
              const matrix_info res_info, ArType beta, StType* res)
{
    ACCBLAS_CALL(accblas_gemv(accblas_detail::default_handle(),
                              accblas_detail::dtype_of<ArType>::value,
                              accblas_detail::dtype_of<StType>::value,
                              m_info.size[0], m_info.size[1],
                              static_cast<double>(alpha), mtx, m_info.stride, x,
                              x_info.stride, static_cast<double>(beta), res,
                              res_info.stride, nullptr));
}

// Range-typed convenience overload: unpacks reduced_row_major ranges and takes
// the same fast path.
template <typename ArType, typename StType>
void acc_gemv(
    ArType alpha,
    const gko::acc::range<gko::acc::reduced_row_major<2, ArType, const StType>>&
        mtx,
    const gko::acc::range<gko::acc::reduced_row_major<2, ArType, const StType>>&
        x,
    ArType beta,
    const gko::acc::range<gko::acc::reduced_row_major<2, ArType, StType>>& res)
{
    acc_gemv<ArType, StType>(
        matrix_info{{mtx.length(0), mtx.length(1)}, mtx->get_stride(0)}, alpha,
        mtx->get_const_storage(),
        matrix_info{{x.length(0), x.length(1)}, x->get_stride(0)},
        x->get_const_storage(),
        matrix_info{{res.length(0), res.length(1)}, res->get_stride(0)}, beta,
        res->get_stored_data());
}

// Plain row_major ranges (no precision change): the same fast path with
// arithmetic type == storage type -- "all other accessors work accordingly"
// (README.md:19) without falling back to the one-block-per-row template.
template <typename ValueType>
void acc_gemv(ValueType alpha,
              const gko::acc::range<gko::acc::row_major<const ValueType, 2>>& mtx,
              const gko::acc::range<gko::acc::row_major<const ValueType, 2>>& x,
              ValueType beta,
              const gko::acc::range<gko::acc::row_major<ValueType, 2>>& res)
{
    gemv<ValueType>(
        matrix_info{{mtx.length(0), mtx.length(1)}, mtx->get_stride()[0]},
        alpha, mtx->get_stored_data(),
        matrix_info{{x.length(0), x.length(1)}, x->get_stride()[0]},
        x->get_stored_data(),
        matrix_info{{res.length(0), res.length(1)}, res->get_stride()[0]}, beta,
        res->get_stored_data());
}

// Any other accessor (scaled_reduced_row_major, block_col_major, user types):
// the accessor-generic kernel template, launched as the reference launches it
// (cuda/gemv_kernels.cuh:190-192).  (Constrained on "is a range" instead of
// spelling range<Accessor> in the signature: explicit template arguments of
// the pointer launchers, acc_gemv<double>(...), must not be substituted into
// range<double>.)
namespace accblas_detail {
template <typename T>
struct is_range : std::false_type {};
template <typename Accessor>
struct is_range<gko::acc::range<Accessor>> : std::true_type {};
}  // namespace accblas_detail

template <typename ArType, typename MtxRange, typename XRange,
          typename ResRange,
          typename = typename std::enable_if<
              accblas_detail::is_range<MtxRange>::value &&
              accblas_detail::is_range<XRange>::value &&
              accblas_detail::is_range<ResRange>::value>::type>
void acc_gemv(ArType alpha, const MtxRange& mtx, const XRange& x, ArType beta,
              const ResRange& res)
{
    constexpr std::int64_t block_size = 512;
    const auto rows = mtx.length(0);
    if (rows == 0) {
        return;
    }
    kernel::acc_gemv<block_size>
        <<<static_cast<unsigned>(rows), block_size>>>(alpha, mtx, x, beta, res);
    CUDA_CALL(cudaGetLastError());
}


inline void cublas_gemv(cublasHandle_t handle, cublasOperation_t transa, int m,
                        int n, const double* alpha, const double* A, int lda,
                        const double* x, int incx, const double* beta,
                        double* y, int incy)
{
    CUBLAS_CALL(
        cublasDgemv(handle, transa, m, n, alpha, A, lda, x, incx, beta, y, incy));
}

inline void cublas_gemv(cublasHandle_t handle, cublasOperation_t transa, int m,
                        int n, const float* alpha, const float* A, int lda,
                        const float* x, int incx, const float* beta, float* y,
                        int incy)
{
    CUBLAS_CALL(
        cublasSgemv(handle, transa, m, n, alpha, A, lda, x, incx, beta, y, incy));
}

// Vendor GEMV on the row-major operands: the matrix is the transpose of what
// cuBLAS sees, so OP_T with (cols, rows) and lda = row stride.  (The reference
// passes (rows, cols), which is only right for square matrices,
// cuda/gemv_kernels.cuh:238-239.)
template <typename ValueType>
void cublas_gemv(cublasHandle_t handle, const matrix_info m_info,
                 ValueType alpha, const ValueType* mtx,
                 const matrix_info x_info, const ValueType* x,
                 const matrix_info res_info, ValueType beta, ValueType* y)
{
    cublas_gemv(handle, CUBLAS_OP_T, static_cast<int>(m_info.size[1]),
                static_cast<int>(m_info.size[0]), &alpha, mtx,
                static_cast<int>(m_info.stride), x,
                static_cast<int>(x_info.stride), &beta, y,
                static_cast<int>(res_info.stride));
}
