// GEMV  y = alpha * A * x + beta * y  over reduced-precision row-major storage.
//
// Replaces kernel::acc_gemv / kernel::gemv (one 512-thread CTA per row, scalar
// 4/8-byte loads, x re-read and re-converted per element, a shared-memory tree
// per row; /root/reference/cuda/gemv_kernels.cuh:30-113) with a streaming
// kernel shaped for B200:
//   * a warp owns ROWS consecutive rows and walks the columns in chunks of
//     32 lanes x UNROLL 128-bit vectors, so ROWS*UNROLL independent 16-byte
//     L1-bypassing loads are in flight per lane and every request is a fully
//     coalesced 512-byte line group;
//   * the matching x vectors are loaded once per chunk through L1 (they are
//     shared by all warps of the SM), converted to the arithmetic type once and
//     reused for the ROWS rows;
//   * storage -> arithmetic conversion happens in registers, accumulation is
//     FMA in the arithmetic type, the row reduction is a warp-shuffle butterfly
//     (no shared-memory tree, no __syncthreads in the hot loop);
//   * for short matrices COLW warps share a row group (split columns) and
//     combine through shared memory in a fixed order.
// The result is rounded once to the storage type on the way out, exactly like
// the accessor's proxy assignment (cuda/gemv_kernels.cuh:106-111).
#include "common.cuh"
#include "tuning.h"

namespace accblas {
namespace {

template <typename Ar, typename St, int ROWS, int UNROLL>
struct ChunkOps {
    static constexpr int VEC = vec_traits<St>::elems;
    static constexpr int CHUNK = kWarp * VEC * UNROLL;

    // full chunk: no predicates
    static __device__ __forceinline__ void full(const St* const (&row)[ROWS],
                                                const St* __restrict__ x,
                                                std::int64_t c0, int lane,
                                                Ar (&acc)[ROWS])
    {
        uint4 xr[UNROLL];
        uint4 ar[ROWS][UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            xr[u] = ldg_cached_128(x + c0 + (u * kWarp + lane) * VEC);
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                ar[r][u] =
                    ldg_stream_128(row[r] + c0 + (u * kWarp + lane) * VEC);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            Ar xv[VEC];
            unpack_all<Ar, St>(xr[u], xv);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    acc[r] = fma_ar(unpack<Ar>(ar[r][u], i, St{}), xv[i],
                                    acc[r]);
                }
            }
        }
    }

    // last, partial chunk: whole vectors while they fit, then scalars
    static __device__ __forceinline__ void partial(
        const St* const (&row)[ROWS], const St* __restrict__ x,
        std::int64_t c0, std::int64_t n, int lane, Ar (&acc)[ROWS])
    {
        for (int u = 0; u < UNROLL; ++u) {
            const std::int64_t col = c0 + (u * kWarp + lane) * VEC;
            if (col + VEC <= n) {
                Ar xv[VEC];
                unpack_all<Ar, St>(ldg_cached_128(x + col), xv);
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const uint4 a = ldg_stream_128(row[r] + col);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        acc[r] = fma_ar(unpack<Ar>(a, i, St{}), xv[i], acc[r]);
                    }
                }
            } else {
                for (std::int64_t c = col; c < n; ++c) {
                    const Ar xv = to_ar<Ar, St>(x[c]);
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) {
                        acc[r] = fma_ar(to_ar<Ar, St>(row[r][c]), xv, acc[r]);
                    }
                }
            }
        }
    }
};

// alpha * sum (+ beta * y), rounded to storage
template <typename Ar, typename St>
__device__ __forceinline__ void write_row(St* y, std::int64_t idx, Ar alpha,
                                          Ar beta, Ar sum)
{
    Ar out;
    if (beta == Ar{0}) {
        out = alpha * sum;
    } else {
        out = fma_ar(alpha, sum, beta * to_ar<Ar, St>(y[idx]));
    }
    y[idx] = to_st<St, Ar>(out);
}

// Requirements (checked by the launcher): A and x 16-byte aligned,
// lda * sizeof(St) a multiple of 16, incx == 1.
// CTA = RG row groups x COLW column-splitting warps.
template <typename St, typename Ar, int ROWS, int UNROLL, int RG, int COLW>
__global__ __launch_bounds__(RG* COLW* kWarp) void gemv_stream_kernel(
    std::int64_t m, std::int64_t n, Ar alpha, const St* __restrict__ A,
    std::int64_t lda, const St* __restrict__ x, Ar beta, St* __restrict__ y,
    std::int64_t incy)
{
    using Ops = ChunkOps<Ar, St, ROWS, UNROLL>;
    constexpr int CHUNK = Ops::CHUNK;
    const int lane = threadIdx.x & (kWarp - 1);
    const int warp = threadIdx.x >> 5;
    const int rg = warp / COLW;  // row group inside the CTA
    const int cw = warp % COLW;  // column slot inside the row group

    const std::int64_t group = std::int64_t{blockIdx.x} * RG + rg;
    const std::int64_t row0 = group * ROWS;
    const bool active = row0 < m;

    Ar acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        acc[r] = Ar{};
    }

    if (active) {
        const St* row[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            // rows past the end re-read the last valid row; never written
            const std::int64_t ri = (row0 + r < m) ? row0 + r : m - 1;
            row[r] = A + ri * lda;
        }
        const std::int64_t full_chunks = n / CHUNK;
        std::int64_t k = cw;
        for (; k < full_chunks; k += COLW) {
            Ops::full(row, x, k * CHUNK, lane, acc);
        }
        if (k == full_chunks && full_chunks * CHUNK < n) {
            Ops::partial(row, x, full_chunks * CHUNK, n, lane, acc);
        }
    }

#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        acc[r] = warp_sum(acc[r]);
    }

    if (COLW == 1) {
        if (active && lane < ROWS && row0 + lane < m) {
            Ar mine = acc[0];
#pragma unroll
            for (int r = 1; r < ROWS; ++r) {
                mine = (lane == r) ? acc[r] : mine;
            }
            write_row<Ar, St>(y, (row0 + lane) * incy, alpha, beta, mine);
        }
    } else {
        __shared__ Ar part[RG][COLW][ROWS];
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                part[rg][cw][r] = acc[r];
            }
        }
        __syncthreads();
        if (active && cw == 0 && lane < ROWS && row0 + lane < m) {
            Ar sum = part[rg][0][lane];
#pragma unroll
            for (int c = 1; c < COLW; ++c) {
                sum += part[rg][c][lane];
            }
            write_row<Ar, St>(y, (row0 + lane) * incy, alpha, beta, sum);
        }
    }
}

// Any layout: one CTA per row, scalar (still coalesced for incx == 1) loads.
template <typename St, typename Ar, int BLOCK>
__global__ __launch_bounds__(BLOCK) void gemv_generic_kernel(
    std::int64_t m, std::int64_t n, Ar alpha, const St* __restrict__ A,
    std::int64_t lda, const St* __restrict__ x, std::int64_t incx, Ar beta,
    St* __restrict__ y, std::int64_t incy)
{
    __shared__ Ar scratch[kWarp];
    for (std::int64_t row = blockIdx.x; row < m; row += gridDim.x) {
        const St* a = A + row * lda;
        Ar acc0 = Ar{}, acc1 = Ar{};
        std::int64_t c = threadIdx.x;
        for (; c + BLOCK < n; c += 2 * BLOCK) {
            const St a0 = a[c], a1 = a[c + BLOCK];
            const St x0 = x[c * incx], x1 = x[(c + BLOCK) * incx];
            acc0 = fma_ar(to_ar<Ar, St>(a0), to_ar<Ar, St>(x0), acc0);
            acc1 = fma_ar(to_ar<Ar, St>(a1), to_ar<Ar, St>(x1), acc1);
        }
        if (c < n) {
            acc0 = fma_ar(to_ar<Ar, St>(a[c]), to_ar<Ar, St>(x[c * incx]), acc0);
        }
        const Ar sum = block_sum(acc0 + acc1, scratch);
        if (threadIdx.x == 0) {
            write_row<Ar, St>(y, row * incy, alpha, beta, sum);
        }
    }
}

template <typename St, typename Ar, int ROWS, int UNROLL, int RG, int COLW>
int launch_stream(std::int64_t m, std::int64_t n, Ar alpha, const St* A,
                  std::int64_t lda, const St* x, Ar beta, St* y,
                  std::int64_t incy, cudaStream_t stream)
{
    const std::int64_t groups = (m + ROWS - 1) / ROWS;
    const std::int64_t grid = (groups + RG - 1) / RG;
    if (grid > 0x7fffffffLL) {
        set_error("gemv: too many rows (%lld)", static_cast<long long>(m));
        return ACCBLAS_ERR_INVALID;
    }
    gemv_stream_kernel<St, Ar, ROWS, UNROLL, RG, COLW>
        <<<static_cast<unsigned>(grid), RG * COLW * kWarp, 0, stream>>>(
            m, n, alpha, A, lda, x, beta, y, incy);
    ACCBLAS_CUDA(cudaGetLastError());
    return ACCBLAS_OK;
}

template <typename St, typename Ar, int UNROLL>
int launch_by_shape(Handle* h, std::int64_t m, std::int64_t n, Ar alpha,
                    const St* A, std::int64_t lda, const St* x, Ar beta, St* y,
                    std::int64_t incy, cudaStream_t stream)
{
    int variant = tuning().gemv_variant;
    if (variant == 0) {
        // enough 4-row groups to give every SM >= 16 warps -> warp per group
        const std::int64_t warps_wanted = std::int64_t{h->sm_count} * 16;
        if ((m + 3) / 4 >= warps_wanted) {
            variant = 1;
        } else if ((m + 1) / 2 >= std::int64_t{h->sm_count} * 2) {
            variant = 2;
        } else {
            variant = 3;
        }
    }
    switch (variant) {
    case 1:
        return launch_stream<St, Ar, 4, UNROLL, 8, 1>(m, n, alpha, A, lda, x,
                                                      beta, y, incy, stream);
    case 2:
        return launch_stream<St, Ar, 2, UNROLL, 1, 8>(m, n, alpha, A, lda, x,
                                                      beta, y, incy, stream);
    default:
        return launch_stream<St, Ar, 1, UNROLL, 1, 8>(m, n, alpha, A, lda, x,
                                                      beta, y, incy, stream);
    }
}

template <typename St, typename Ar>
int launch_gemv(Handle* h, std::int64_t m, std::int64_t n, double alpha_d,
                const void* A_v, std::int64_t lda, const void* x_v,
                std::int64_t incx, double beta_d, void* y_v, std::int64_t incy,
                cudaStream_t stream)
{
    const Ar alpha = static_cast<Ar>(alpha_d);
    const Ar beta = static_cast<Ar>(beta_d);
    const St* A = static_cast<const St*>(A_v);
    const St* x = static_cast<const St*>(x_v);
    St* y = static_cast<St*>(y_v);
    if (m == 0) {
        return ACCBLAS_OK;
    }
    const bool vec_ok =
        incx == 1 &&
        ((reinterpret_cast<std::uintptr_t>(A) |
          reinterpret_cast<std::uintptr_t>(x)) & 15u) == 0 &&
        (static_cast<std::uint64_t>(lda) * sizeof(St)) % 16 == 0;
    if (vec_ok) {
        if (tuning().gemv_unroll == 4) {
            return launch_by_shape<St, Ar, 4>(h, m, n, alpha, A, lda, x, beta,
                                              y, incy, stream);
        }
        if (tuning().gemv_unroll == 1) {
            return launch_by_shape<St, Ar, 1>(h, m, n, alpha, A, lda, x, beta,
                                              y, incy, stream);
        }
        return launch_by_shape<St, Ar, 2>(h, m, n, alpha, A, lda, x, beta, y,
                                          incy, stream);
    }
    constexpr int BLOCK = 256;
    std::int64_t grid = m;
    const std::int64_t cap = std::int64_t{h->sm_count} * 64;
    if (grid > cap) {
        grid = cap;
    }
    gemv_generic_kernel<St, Ar, BLOCK>
        <<<static_cast<unsigned>(grid), BLOCK, 0, stream>>>(
            m, n, alpha, A, lda, x, incx, beta, y, incy);
    ACCBLAS_CUDA(cudaGetLastError());
    return ACCBLAS_OK;
}

}  // namespace

int gemv_impl(Handle* h, int ar, int st, std::int64_t m, std::int64_t n,
              double alpha, const void* A, std::int64_t lda, const void* x,
              std::int64_t incx, double beta, void* y, std::int64_t incy,
              cudaStream_t stream)
{
    return dispatch_ar_st(ar, st, [&](auto st_tag, auto ar_tag) {
        using St = decltype(st_tag);
        using Ar = decltype(ar_tag);
        return launch_gemv<St, Ar>(h, m, n, alpha, A, lda, x, incx, beta, y,
                                   incy, stream);
    });
}

}  // namespace accblas
