// Drop-in replacement for the reference's trsv_kernels.cuh
// (/root/reference/cuda/trsv_kernels.cuh): same names, template parameters and
// argument meaning.
//
//   tmtx_t, dmtx_t  triangle / diagonal selectors
//   host launchers  trsv<T>, acc_trsv<Ar, St>     -> accblas_trsv: the blocked,
//                   sync-free sm_100a solve in libaccblas_b200.so.  The
//                   `trsv_helper` argument (two uint32 the reference needs for
//                   its block ordering) is accepted and left untouched; the
//                   library keeps its progress vector in its own workspace.
//   kernel::trsv_init, kernel::{lower,upper}_trsv, kernel::acc_{lower,upper}_trsv
//                   accessor-generic __global__ kernels with the documented
//                   signatures and launch shape (grid = ceil(n / swarp_size),
//                   block = (swarp_size, swarps_per_block)).  Own
//                   implementation: ticket ordering, acquire/release progress
//                   flag (instead of volatile + __threadfence), diagonal block
//                   solved by substitution (no explicit inverse).
//   cublas_trsv     vendor baseline.
#pragma once

#include <cinttypes>

#include <cublas_v2.h>

#include <accessor/range.hpp>
#include <accessor/reduced_row_major.hpp>

#include "kernel_utils.cuh"
#include "utils.cuh"

enum class tmtx_t { upper, lower };
enum class dmtx_t { non_unit, unit };

namespace kernel {

__global__ __launch_bounds__(1) void trsv_init(std::uint32_t* block_idxs)
{
    block_idxs[0] = ~std::uint32_t{0};  // last finished block (none)
    block_idxs[1] = 0;                  // next block to hand out
}

namespace detail {

__device__ __forceinline__ std::int32_t load_acquire(const std::uint32_t* p)
{
    std::int32_t v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];"
                 : "=r"(v)
                 : "l"(p)
                 : "memory");
    return v;
}

__device__ __forceinline__ void store_release(std::uint32_t* p, std::uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v)
                 : "memory");
}

// One CTA solves one swarp_size-row block; blocks are numbered in solve order
// (top to bottom for lower, bottom to top for upper).  Element access only
// through mtx(i, j) / x(i, 0), so any accessor works.
template <std::int32_t swarp_size, std::int32_t swarps_per_block, dmtx_t dmtx,
          bool upper, typename MtxRange, typename VecRange>
__device__ __forceinline__ void trsv_block(MtxRange mtx, VecRange x,
                                           std::uint32_t* helper)
{
    static_assert(swarp_size <= WARP_SIZE && (swarp_size & (swarp_size - 1)) == 0,
                  "swarp_size must be a power of two not larger than a warp");
    static_assert(swarp_size % swarps_per_block == 0,
                  "swarp_size must be a multiple of swarps_per_block");
    using ar_type = decltype(mtx(0, 0) * x(0, 0));
    using index_type = std::int64_t;
    constexpr int S = swarp_size;
    constexpr int W = swarps_per_block;
    constexpr int rows_per_thread = S / W;

    __shared__ ar_type tri[S][S + 1];
    __shared__ ar_type rhs[S];
    __shared__ std::int32_t block_shared;

    const int tx = threadIdx.x;  // column inside a tile
    const int ty = threadIdx.y;
    const index_type n = mtx.length(0);
    if (tx == 0 && ty == 0) {
        block_shared = atomicInc(helper + 1, ~std::uint32_t{0});
    }
    __syncthreads();
    const index_type rb = block_shared;
    if (rb * S >= n) {
        return;
    }
    // global index of local row/column 0 of this block (may be negative for
    // the last upper block when n is not a multiple of S)
    const index_type base = upper ? n - (rb + 1) * S : rb * S;
    auto inside = [&](index_type g) { return g >= 0 && g < n; };

    for (int r = ty; r < S; r += W) {
        const index_type gr = base + r, gc = base + tx;
        const bool in_tri = upper ? (tx >= r) : (tx <= r);
        ar_type v = (r == tx) ? ar_type{1} : ar_type{0};
        if (in_tri && inside(gr) && inside(gc) &&
            !(dmtx == dmtx_t::unit && r == tx)) {
            v = mtx(gr, gc);
        }
        tri[r][tx] = v;
    }

    ar_type acc[rows_per_thread];
#pragma unroll
    for (int l = 0; l < rows_per_thread; ++l) {
        acc[l] = ar_type{};
    }
    for (index_type cb = 0; cb < rb; ++cb) {
        if (tx == 0 && ty == 0) {
            while (load_acquire(helper) < cb) {
            }
        }
        __syncthreads();
        const index_type gc = (upper ? n - (cb + 1) * S : cb * S) + tx;
        // written by another CTA during this launch: bypass the (incoherent) L1
        __threadfence();
        const ar_type xc = x(gc, 0);
#pragma unroll
        for (int l = 0; l < rows_per_thread; ++l) {
            const index_type gr = base + ty + l * W;
            if (inside(gr)) {
                acc[l] += mtx(gr, gc) * xc;
            }
        }
    }
    const auto tile = cg::tiled_partition<S>(cg::this_thread_block());
#pragma unroll
    for (int l = 0; l < rows_per_thread; ++l) {
        const ar_type sum =
            reduce(tile, acc[l], [](ar_type a, ar_type b) { return a + b; });
        if (tx == 0) {
            const int r = ty + l * W;
            const index_type gr = base + r;
            rhs[r] = inside(gr) ? static_cast<ar_type>(x(gr, 0)) - sum
                                : ar_type{0};
        }
    }
    __syncthreads();

    if (ty == 0) {
        // substitution inside the block: lane r owns unknown r
        ar_type mine = rhs[tx];
        const ar_type diag = tri[tx][tx];
        for (int step = 0; step < S; ++step) {
            const int c = upper ? S - 1 - step : step;
            ar_type pivot = mine;
            if (dmtx == dmtx_t::non_unit && tx == c) {
                pivot = mine / diag;
            }
            const ar_type xc = tile.shfl(pivot, c);
            if (tx == c) {
                mine = xc;
            } else if (upper ? (tx < c) : (tx > c)) {
                mine -= tri[tx][c] * xc;
            }
        }
        const index_type gr = base + tx;
        if (inside(gr)) {
            x(gr, 0) = mine;
        }
        tile.sync();
        if (tx == 0) {
            store_release(helper, static_cast<std::uint32_t>(rb));
        }
    }
}

// plain pointers wrapped into a trivial row-major accessor
template <typename ValueType>
struct plain_matrix {
    const ValueType* data;
    std::int64_t rows, stride;
    __device__ ValueType operator()(std::int64_t r, std::int64_t c) const
    {
        return data[r * stride + c];
    }
    __device__ std::int64_t length(int) const { return rows; }
};
template <typename ValueType>
struct plain_vector {
    ValueType* data;
    std::int64_t rows, stride;
    __device__ ValueType& operator()(std::int64_t r, std::int64_t) const
    {
        return data[r * stride];
    }
    __device__ std::int64_t length(int) const { return rows; }
};

}  // namespace detail

template <std::int32_t swarp_size, std::int32_t swarps_per_block, dmtx_t dmtx,
          typename ValueType>
__global__ __launch_bounds__(swarps_per_block* swarp_size) void lower_trsv(
    const matrix_info m_info, const ValueType* __restrict__ mtx,
    const matrix_info x_info, ValueType* __restrict__ x,
    std::uint32_t* col_row_global_helper)
{
    detail::trsv_block<swarp_size, swarps_per_block, dmtx, false>(
        detail::plain_matrix<ValueType>{mtx, m_info.size[0], m_info.stride},
        detail::plain_vector<ValueType>{x, x_info.size[0], x_info.stride},
        col_row_global_helper);
}

template <std::int32_t swarp_size, std::int32_t swarps_per_block, dmtx_t dmtx,
          typename ValueType>
__global__ __launch_bounds__(swarps_per_block* swarp_size) void upper_trsv(
    const matrix_info m_info, const ValueType* __restrict__ mtx,
    const matrix_info x_info, ValueType* __restrict__ x,
    std::uint32_t* col_row_global_helper)
{
    detail::trsv_block<swarp_size, swarps_per_block, dmtx, true>(
        detail::plain_matrix<ValueType>{mtx, m_info.size[0], m_info.stride},
        detail::plain_vector<ValueType>{x, x_info.size[0], x_info.stride},
        col_row_global_helper);
}

template <std::int32_t swarp_size, std::int32_t swarps_per_block, dmtx_t dmtx,
          typename MtxAccessor, typename VecAccessor>
__global__ __launch_bounds__(swarps_per_block* swarp_size) void acc_lower_trsv(
    gko::acc::range<MtxAccessor> mtx, gko::acc::range<VecAccessor> x,
    std::uint32_t* col_row_global_helper)
{
    detail::trsv_block<swarp_size, swarps_per_block, dmtx, false>(
        mtx, x, col_row_global_helper);
}

template <std::int32_t swarp_size, std::int32_t swarps_per_block, dmtx_t dmtx,
          typename MtxAccessor, typename VecAccessor>
__global__ __launch_bounds__(swarps_per_block* swarp_size) void acc_upper_trsv(
    gko::acc::range<MtxAccessor> mtx, gko::acc::range<VecAccessor> x,
    std::uint32_t* col_row_global_helper)
{
    detail::trsv_block<swarp_size, swarps_per_block, dmtx, true>(
        mtx, x, col_row_global_helper);
}

}  // namespace kernel


// In-place solve with the `ttype` triangle of the row-major matrix, arithmetic
// type == storage type (cuda/trsv_kernels.cuh:455-488).
template <typename ValueType>
void trsv(const matrix_info m_info, tmtx_t ttype, dmtx_t dtype,
          const ValueType* mtx, const matrix_info x_info, ValueType* x,
          std::uint32_t* /*trsv_helper*/)
{
    constexpr accblas_dtype t = accblas_detail::dtype_of<ValueType>::value;
    ACCBLAS_CALL(accblas_trsv(
        accblas_detail::default_handle(), t, t,
        ttype == tmtx_t::upper ? ACCBLAS_UPPER : ACCBLAS_LOWER,
        dtype == dmtx_t::unit ? ACCBLAS_UNIT : ACCBLAS_NON_UNIT, m_info.size[0],
        mtx, m_info.stride, x, x_info.stride, nullptr));
}

// The same computed in ArType on StType storage (cuda/trsv_kernels.cuh:918-961).
template <typename ArType, typename StType>
void acc_trsv(const matrix_info m_info, tmtx_t ttype, dmtx_t dtype,
              const StType* mtx, const matrix_info x_info, StType* x,
              std::uint32_t* /*trsv_helper*/)
{
    ACCBLAS_CALL(accblas_trsv(
        accblas_detail::default_handle(),
        accblas_detail::dtype_of<ArType>::value,
        accblas_detail::dtype_of<StType>::value,
        ttype == tmtx_t::upper ? ACCBLAS_UPPER : ACCBLAS_LOWER,
        dtype == dmtx_t::unit ? ACCBLAS_UNIT : ACCBLAS_NON_UNIT, m_info.size[0],
        mtx, m_info.stride, x, x_info.stride, nullptr));
}


inline void cublas_trsv(cublasHandle_t handle, cublasFillMode_t uplo,
                        cublasOperation_t trans, cublasDiagType_t diag, int n,
                        const double* A, int lda, double* x, int incx)
{
    CUBLAS_CALL(cublasDtrsv(handle, uplo, trans, diag, n, A, lda, x, incx));
}

inline void cublas_trsv(cublasHandle_t handle, cublasFillMode_t uplo,
                        cublasOperation_t trans, cublasDiagType_t diag, int n,
                        const float* A, int lda, float* x, int incx)
{
    CUBLAS_CALL(cublasStrsv(handle, uplo, trans, diag, n, A, lda, x, incx));
}

// Vendor solve on the row-major operands: cuBLAS sees the transpose, so the
// fill mode is swapped and OP_T is used.
template <typename ValueType>
void cublas_trsv(cublasHandle_t handle, tmtx_t ttype, dmtx_t dtype,
                 const matrix_info m_info, const ValueType* mtx,
                 const matrix_info x_info, ValueType* x)
{
    const cublasFillMode_t uplo = ttype == tmtx_t::upper
                                      ? CUBLAS_FILL_MODE_LOWER
                                      : CUBLAS_FILL_MODE_UPPER;
    const cublasDiagType_t diag =
        dtype == dmtx_t::unit ? CUBLAS_DIAG_UNIT : CUBLAS_DIAG_NON_UNIT;
    cublas_trsv(handle, uplo, CUBLAS_OP_T, diag,
                static_cast<int>(m_info.size[0]), mtx,
                static_cast<int>(m_info.stride), x,
                static_cast<int>(x_info.stride));
}
