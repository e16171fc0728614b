// TRSV: blocked, single-launch, sync-free triangular solve over
// reduced-precision row-major storage.
//
// Replaces kernel::acc_{lower,upper}_trsv / kernel::{lower,upper}_trsv +
// kernel::trsv_init (/root/reference/cuda/trsv_kernels.cuh:38-42,69-432,
// 527-893: 32-row blocks, 128 threads, a volatile spin + __syncthreads +
// __threadfence per 32 columns, one warp inverting the diagonal tile with 496
// dependent shared-memory steps) with:
//   * 128-row block rows, one 512-thread CTA each, ordered by an atomic ticket
//     (a CTA only ever waits on CTAs that took an earlier ticket, so the
//     launch cannot deadlock however many CTAs are resident);
//   * off-diagonal updates done GEMV-style: warp = 8 rows, lane = 4 columns,
//     wide L1-bypassing loads issued BEFORE the wait for the matching x
//     block, converted in registers, FMA in the arithmetic type;
//   * progress communicated through the solution itself: solved entries are
//     published (already rounded through the storage type, as the reference's
//     accessor write/read does) into a workspace vector that starts out as a
//     NaN sentinel; consumers poll the values they need with volatile loads, so
//     one L2 round trip carries both "ready" and the data -- no flag, no
//     fence, no acquire/release pair on the critical path;
//   * the diagonal 128x128 tile lives in shared memory; its four 32x32
//     diagonal sub-blocks are inverted by Gauss-Jordan with all 16 warps
//     (4 per sub-block) while the CTA would otherwise be waiting, and the
//     solve walks the sub-blocks left-looking.
// The last CTA re-arms the workspace (sentinels, ticket) for the next call.
#include "common.cuh"
#include "tuning.h"

namespace accblas {
namespace {

constexpr int kB = 128;       // rows/cols per block row
constexpr int kSB = 32;       // diagonal sub-block
constexpr int kNSB = kB / kSB;
constexpr int kLD = kB + 1;   // padded leading dimension of the smem tile
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / kWarp;
constexpr int kRowsPerWarp = kB / kWarps;  // 8
constexpr int kEPL = 4;                    // elements per lane per row

template <typename Ar>
struct Sentinel;
template <>
struct Sentinel<double> {
    static __device__ __forceinline__ bool is(double v)
    {
        return __double_as_longlong(v) == -1LL;
    }
    static __device__ __forceinline__ double clean(double v)
    {
        return is(v) ? __longlong_as_double(0x7ff8000000000000LL) : v;
    }
};
template <>
struct Sentinel<float> {
    static __device__ __forceinline__ bool is(float v)
    {
        return __float_as_int(v) == -1;
    }
    static __device__ __forceinline__ float clean(float v)
    {
        return is(v) ? __int_as_float(0x7fc00000) : v;
    }
};

template <typename T>
__device__ __forceinline__ T ld_volatile(const T* p)
{
    return *reinterpret_cast<const volatile T*>(p);
}
template <typename T>
__device__ __forceinline__ void st_volatile(T* p, T v)
{
    *reinterpret_cast<volatile T*>(p) = v;
}

template <typename St>
struct Quad {
    St v[kEPL];
};

__device__ __forceinline__ uint2 ldg_stream_64(const void* p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];"
                 : "=r"(r.x), "=r"(r.y)
                 : "l"(p));
    return r;
}

template <typename St>
__device__ __forceinline__ St zero_st()
{
    return St(0);
}
template <>
__device__ __forceinline__ __half zero_st<__half>()
{
    return __ushort_as_half(0);
}

// four consecutive elements; `valid` of them are inside the matrix
template <typename St, bool VECTOR>
__device__ __forceinline__ Quad<St> load_quad(const St* p, int valid)
{
    Quad<St> q;
    if (VECTOR && valid == kEPL) {
        if (sizeof(St) == 8) {
            const uint4 a = ldg_stream_128(p);
            const uint4 b = ldg_stream_128(p + 2);
            uint4* dst = reinterpret_cast<uint4*>(&q);
            dst[0] = a;
            dst[1] = b;
        } else if (sizeof(St) == 4) {
            *reinterpret_cast<uint4*>(&q) = ldg_stream_128(p);
        } else {
            *reinterpret_cast<uint2*>(&q) = ldg_stream_64(p);
        }
    } else {
#pragma unroll
        for (int e = 0; e < kEPL; ++e) {
            q.v[e] = (e < valid) ? p[e] : zero_st<St>();
        }
    }
    return q;
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void group_barrier(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// In-place inverse of the four 32x32 diagonal sub-blocks of D (row-major,
// leading dimension kLD).  Same elimination sequence per entry as the
// reference's Gauss-Jordan (cuda/trsv_kernels.cuh:583-620 / 784-821); rows of
// one elimination step are independent, so 4 warps share a sub-block.
template <typename Ar, bool UPPER, bool UNIT>
__device__ __forceinline__ void invert_diag_subblocks(Ar* D, int warp, int lane)
{
    const int g = warp >> 2;  // sub-block
    const int q = warp & 3;   // warp inside the group
    Ar* T = D + (g * kSB) * kLD + g * kSB;
    const int c = lane;
    if (!UNIT) {
        for (int row = q; row < kSB; row += 4) {
            const Ar inv = Ar{1} / T[row * kLD + row];
            const bool in_tri = UPPER ? (c > row) : (c < row);
            const Ar cur = T[row * kLD + c];
            __syncwarp();
            if (c == row) {
                T[row * kLD + c] = inv;
            } else if (in_tri) {
                T[row * kLD + c] = cur * inv;
            }
        }
        group_barrier(1 + g, 4 * kWarp);
    }
    if (!UPPER) {
        for (int d = 0; d < kSB; ++d) {
            const Ar diag_el = T[d * kLD + d];
            const Ar piv = T[d * kLD + c];
            for (int row = d + 1 + q; row < kSB; row += 4) {
                const Ar factor = -T[row * kLD + d];
                const Ar cur = T[row * kLD + c];
                __syncwarp();
                if (c < row) {
                    T[row * kLD + c] =
                        (c == d) ? factor * diag_el : fma_ar(factor, piv, cur);
                }
            }
            group_barrier(1 + g, 4 * kWarp);
        }
    } else {
        for (int d = kSB - 1; d >= 0; --d) {
            const Ar diag_el = T[d * kLD + d];
            const Ar piv = T[d * kLD + c];
            for (int row = d - 1 - q; row >= 0; row -= 4) {
                const Ar factor = -T[row * kLD + d];
                const Ar cur = T[row * kLD + c];
                __syncwarp();
                if (c > row) {
                    T[row * kLD + c] =
                        (c == d) ? factor * diag_el : fma_ar(factor, piv, cur);
                }
            }
            group_barrier(1 + g, 4 * kWarp);
        }
    }
}

template <typename St, typename Ar, bool UPPER, bool UNIT, bool VECTOR>
__global__ __launch_bounds__(kThreads, 1) void trsv_kernel(
    std::int64_t n, const St* __restrict__ A, std::int64_t lda,
    St* __restrict__ x, std::int64_t incx, Ar* xs,
    unsigned* __restrict__ ticket, long long* __restrict__ trace)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ar* D = reinterpret_cast<Ar*>(smem_raw);  // kB x kLD
    Ar* xcol = D + kB * kLD;                  // 2 x kB, staged x blocks
    Ar* rhs = xcol + 2 * kB;                  // kB
    Ar* xsol = rhs + kB;                      // kB
    __shared__ unsigned k_shared;

    const int tid = threadIdx.x;
    const int lane = tid & (kWarp - 1);
    const int warp = tid >> 5;

    if (tid == 0) {
        k_shared = atomicAdd(ticket, 1u);
    }
    __syncthreads();
    const std::int64_t k = k_shared;  // position in the solve order
    // development aid: per-CTA phase timestamps (SM cycles / global ns)
#define ACCBLAS_TRACE(slot, value)                  \
    if (trace != nullptr && tid == 0) {             \
        trace[k * 16 + (slot)] = (value);           \
    }
    ACCBLAS_TRACE(0, clock64());
    const std::int64_t nb = (n + kB - 1) / kB;
    const std::int64_t pb = UPPER ? nb - 1 - k : k;  // physical block row
    const std::int64_t r0 = pb * kB;
    const int bs = static_cast<int>((n - r0 < kB) ? (n - r0) : kB);

    // ---- diagonal tile -> shared memory (identity padding past the edge)
    for (int idx = tid; idx < kB * kB; idx += kThreads) {
        const int r = idx / kB;
        const int c = idx % kB;
        const bool in_tri = UPPER ? (c >= r) : (c <= r);
        Ar val;
        if (r < bs && c < bs && in_tri) {
            val = (UNIT && r == c)
                      ? Ar{1}
                      : to_ar<Ar, St>(A[(r0 + r) * lda + r0 + c]);
        } else {
            val = (r == c) ? Ar{1} : Ar{0};
        }
        D[r * kLD + c] = val;
    }
    if (tid < kB) {
        rhs[tid] = (tid < bs) ? to_ar<Ar, St>(x[(r0 + tid) * incx]) : Ar{0};
        xsol[tid] = Ar{0};
    }
    __syncthreads();
    ACCBLAS_TRACE(1, clock64());
    invert_diag_subblocks<Ar, UPPER, UNIT>(D, warp, lane);
    ACCBLAS_TRACE(2, clock64());

    // ---- off-diagonal blocks, in solve order
    Ar acc[kRowsPerWarp];
#pragma unroll
    for (int i = 0; i < kRowsPerWarp; ++i) {
        acc[i] = Ar{};
    }
    const St* row_ptr[kRowsPerWarp];
#pragma unroll
    for (int i = 0; i < kRowsPerWarp; ++i) {
        std::int64_t r = r0 + warp * kRowsPerWarp + i;
        r = (r < n) ? r : n - 1;  // padded rows re-read a valid row
        row_ptr[i] = A + r * lda;
    }
    int buf = 0;
    for (std::int64_t jj = 0; jj < k; ++jj) {
        const std::int64_t pbj = UPPER ? nb - 1 - jj : jj;
        const std::int64_t c0 = pbj * kB + lane * kEPL;
        const std::int64_t left = n - c0;
        const int valid = left >= kEPL ? kEPL : (left > 0 ? static_cast<int>(left) : 0);
        Quad<St> raw[kRowsPerWarp];
#pragma unroll
        for (int i = 0; i < kRowsPerWarp; ++i) {
            raw[i] = load_quad<St, VECTOR>(row_ptr[i] + c0, valid);
        }
        if (jj == k - 1) {
            ACCBLAS_TRACE(3, clock64());
        }
        if (warp == 0) {
            Ar v[kEPL];
            bool ok;
            do {
                ok = true;
#pragma unroll
                for (int e = 0; e < kEPL; ++e) {
                    if (e < valid) {
                        v[e] = ld_volatile(xs + c0 + e);
                        ok = ok && !Sentinel<Ar>::is(v[e]);
                    } else {
                        v[e] = Ar{0};
                    }
                }
            } while (!__all_sync(0xffffffffu, ok));
#pragma unroll
            for (int e = 0; e < kEPL; ++e) {
                xcol[buf * kB + lane * kEPL + e] = v[e];
            }
            if (jj == k - 1) {
                ACCBLAS_TRACE(4, clock64());
                ACCBLAS_TRACE(13, static_cast<long long>(globaltimer_ns()));
            }
        }
        __syncthreads();
        if (jj == k - 1) {
            ACCBLAS_TRACE(5, clock64());
        }
        Ar xv[kEPL];
#pragma unroll
        for (int e = 0; e < kEPL; ++e) {
            xv[e] = xcol[buf * kB + lane * kEPL + e];
        }
#pragma unroll
        for (int i = 0; i < kRowsPerWarp; ++i) {
#pragma unroll
            for (int e = 0; e < kEPL; ++e) {
                acc[i] = fma_ar(to_ar<Ar, St>(raw[i].v[e]), xv[e], acc[i]);
            }
        }
        buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < kRowsPerWarp; ++i) {
        acc[i] = warp_sum(acc[i]);
    }
    if (lane < kRowsPerWarp) {
        Ar mine = acc[0];
#pragma unroll
        for (int i = 1; i < kRowsPerWarp; ++i) {
            mine = (lane == i) ? acc[i] : mine;
        }
        if (r0 + warp * kRowsPerWarp + lane < n) {  // padded rows stay zero
            rhs[warp * kRowsPerWarp + lane] -= mine;
        }
    }
    __syncthreads();
    ACCBLAS_TRACE(6, clock64());

    // ---- diagonal block: left-looking over the 32-wide sub-blocks
    for (int step = 0; step < kNSB; ++step) {
        const int s = UPPER ? kNSB - 1 - step : step;
        if (step > 0) {
            // rhs_s -= D[s, solved sub-blocks] * xsol[solved]
            const int first_col = UPPER ? (s + 1) * kSB : 0;
            const int ncols = step * kSB;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int r = s * kSB + warp * 2 + rr;
                Ar sum = Ar{};
                for (int cc = lane; cc < ncols; cc += kWarp) {
                    sum = fma_ar(D[r * kLD + first_col + cc],
                                 xsol[first_col + cc], sum);
                }
                sum = warp_sum(sum);
                if (lane == 0) {
                    rhs[r] -= sum;
                }
            }
            __syncthreads();
        }
        if (warp == 0) {
            const int r = s * kSB + lane;
            const Ar* Trow = D + r * kLD + s * kSB;
            const Ar* v = rhs + s * kSB;
            Ar p0 = Ar{}, p1 = Ar{}, p2 = Ar{}, p3 = Ar{};
#pragma unroll
            for (int cidx = 0; cidx < kSB; cidx += 4) {
                p0 = fma_ar(Trow[cidx + 0], v[cidx + 0], p0);
                p1 = fma_ar(Trow[cidx + 1], v[cidx + 1], p1);
                p2 = fma_ar(Trow[cidx + 2], v[cidx + 2], p2);
                p3 = fma_ar(Trow[cidx + 3], v[cidx + 3], p3);
            }
            const Ar sol = (p0 + p1) + (p2 + p3);
            // round through storage: later rows see what the accessor re-reads
            const St stored = to_st<St, Ar>(sol);
            const Ar back = to_ar<Ar, St>(stored);
            xsol[r] = back;
            const std::int64_t gi = r0 + r;
            if (gi < n) {
                st_volatile(xs + gi, Sentinel<Ar>::clean(back));
                x[gi * incx] = stored;
            }
        }
        __syncthreads();
        ACCBLAS_TRACE(7 + step, clock64());
    }
    ACCBLAS_TRACE(12, static_cast<long long>(globaltimer_ns()));
#undef ACCBLAS_TRACE

    // ---- the last CTA re-arms the workspace for the next call
    if (k == nb - 1) {
        Ar sentinel;
        memset(&sentinel, 0xff, sizeof(Ar));
        for (std::int64_t i = tid; i < n; i += kThreads) {
            xs[i] = sentinel;
        }
        if (tid == 0) {
            *ticket = 0u;
        }
    }
}

template <typename St, typename Ar, bool UPPER, bool UNIT, bool VECTOR>
int launch_one(std::int64_t n, const St* A, std::int64_t lda, St* x,
               std::int64_t incx, Ar* xs, unsigned* ticket, long long* trace,
               cudaStream_t stream)
{
    auto kernel = trsv_kernel<St, Ar, UPPER, UNIT, VECTOR>;
    const size_t smem = sizeof(Ar) * (kB * kLD + 4 * kB);
    static bool configured = false;  // per instantiation
    if (!configured) {
        ACCBLAS_CUDA(cudaFuncSetAttribute(
            kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
            static_cast<int>(smem)));
        configured = true;
    }
    const std::int64_t nb = (n + kB - 1) / kB;
    kernel<<<static_cast<unsigned>(nb), kThreads, smem, stream>>>(
        n, A, lda, x, incx, xs, ticket, trace);
    ACCBLAS_CUDA(cudaGetLastError());
    return ACCBLAS_OK;
}

template <typename St, typename Ar>
int launch_trsv(Handle* h, int uplo, int diag, std::int64_t n, const void* A_v,
                std::int64_t lda, void* x_v, std::int64_t incx,
                long long* trace, cudaStream_t stream)
{
    if (n == 0) {
        return ACCBLAS_OK;
    }
    const St* A = static_cast<const St*>(A_v);
    St* x = static_cast<St*>(x_v);
    // progress vector: n arithmetic values, all sentinel between calls
    const size_t need = static_cast<size_t>(n) * sizeof(Ar);
    void* old_ws = h->ws;
    int rc = ensure_workspace(h, kScratchBytes + need, stream);
    if (rc != ACCBLAS_OK) {
        return rc;
    }
    if (h->ws != old_ws) {
        h->trsv_armed_bytes = 0;
    }
    if (h->trsv_armed_bytes < need) {
        // first use at this size: arm every byte that will be polled (all-ones
        // bytes are the sentinel for fp32 and fp64 alike); afterwards the
        // kernel's last CTA re-arms what it used
        ACCBLAS_CUDA(cudaMemsetAsync(trsv_region(h), 0xff, need, stream));
    }
    h->trsv_armed_bytes = need;
    Ar* xs = static_cast<Ar*>(trsv_region(h));
    unsigned* ticket = control_words(h) + kCtlTrsvTicket;

    const bool vec =
        reinterpret_cast<std::uintptr_t>(A) % 16 == 0 &&
        (static_cast<std::uint64_t>(lda) * sizeof(St)) % 16 == 0;
    const bool upper = uplo == ACCBLAS_UPPER;
    const bool unit = diag == ACCBLAS_UNIT;
#define ACCBLAS_TRSV_CASE(U, N, V)                                          \
    if (upper == U && unit == N && vec == V) {                              \
        return launch_one<St, Ar, U, N, V>(n, A, lda, x, incx, xs, ticket,  \
                                           trace, stream);                  \
    }
    ACCBLAS_TRSV_CASE(false, false, false)
    ACCBLAS_TRSV_CASE(false, false, true)
    ACCBLAS_TRSV_CASE(false, true, false)
    ACCBLAS_TRSV_CASE(false, true, true)
    ACCBLAS_TRSV_CASE(true, false, false)
    ACCBLAS_TRSV_CASE(true, false, true)
    ACCBLAS_TRSV_CASE(true, true, false)
    ACCBLAS_TRSV_CASE(true, true, true)
#undef ACCBLAS_TRSV_CASE
    return ACCBLAS_ERR_INVALID;
}

}  // namespace

int trsv_impl(Handle* h, int ar, int st, int uplo, int diag, std::int64_t n,
              const void* A, std::int64_t lda, void* x, std::int64_t incx,
              cudaStream_t stream, long long* trace)
{
    return dispatch_ar_st(ar, st, [&](auto st_tag, auto ar_tag) {
        using St = decltype(st_tag);
        using Ar = decltype(ar_tag);
        return launch_trsv<St, Ar>(h, uplo, diag, n, A, lda, x, incx, trace,
                                   stream);
    });
}

}  // namespace accblas
