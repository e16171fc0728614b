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
//   * off-diagonal updates done GEMV-style with 4 lanes per row (16
//     consecutive elements per load instruction and row, full sectors): the
//     next panel's L1-bypassing loads are in flight while the current one is
//     consumed, the panel is widened to the arithmetic type BEFORE the wait
//     for the matching x block, and a row sum needs two shuffle levels;
//   * progress communicated through the solution itself: solved entries are
//     published (already rounded through the storage type, as the reference's
//     accessor write/read does) into a workspace vector that starts out as a
//     NaN sentinel; consumers poll the values they need with volatile loads, so
//     one L2 round trip carries both "ready" and the data -- no flag, no
//     fence, no acquire/release pair on the critical path;
//   * the diagonal 128x128 tile lives in shared memory (leading dimension 136
//     so 16-byte loads are conflict-free); each of its four 32x32 diagonal
//     sub-blocks is inverted by one warp, lane j = column j by substitution in
//     registers, and the solve walks the sub-blocks left-looking;
//   * the reduction + diagonal-solve code is rehearsed once on scratch data
//     while the CTA waits, so it is warm in the instruction cache when it
//     runs for real on the critical path.
// The last CTA re-arms the workspace (sentinels, ticket) for the next call.
#include "common.cuh"
#include "tuning.h"

namespace accblas {
namespace {

constexpr int kB = 128;       // rows/cols per block row
constexpr int kSB = 32;       // diagonal sub-block
constexpr int kNSB = kB / kSB;
// Leading dimension of the smem tile: 136 = 8 * 17, so that 16-byte (two
// double) loads by 2 rows x 4 column slots per quarter-warp hit eight distinct
// 16-byte bank groups (row stride 136 doubles = 68 groups = 4 mod 8).
constexpr int kLD = kB + 8;
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / kWarp;
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

template <typename Ar>
struct alignas(2 * sizeof(Ar)) Pair {
    Ar a, b;
};

// keeps a converted value in a register at this point of the program (the
// compiler would otherwise sink the conversion below the barrier that follows)
__device__ __forceinline__ void pin_register(double& v)
{
    asm volatile("" : "+d"(v));
}
__device__ __forceinline__ void pin_register(float& v)
{
    asm volatile("" : "+f"(v));
}

__device__ __forceinline__ void named_barrier_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_barrier_arrive(int id, int threads)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
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


// Inverse of one 32x32 triangular sub-block T of D (row-major, leading
// dimension kLD), in place, by ONE warp: lane j computes column j of T^-1 by
// substitution entirely in registers (every T(i,k) is a shared-memory
// broadcast), so the 496 steps of the reference's Gauss-Jordan sweep
// (/root/reference/cuda/trsv_kernels.cuh:583-620 / 784-821), each separated by
// a warp barrier, become one barrier-free unrolled pass.
template <typename Ar, bool UPPER, bool UNIT>
__device__ __forceinline__ void invert_subblock(Ar* T, Ar* inv_diag, int lane)
{
    if (!UNIT) {
        inv_diag[lane] = Ar{1} / T[lane * kLD + lane];
    }
    __syncwarp();
    Ar z[kSB];
#pragma unroll
    for (int step = 0; step < kSB; ++step) {
        const int i = UPPER ? kSB - 1 - step : step;
        Ar s0 = (i == lane) ? Ar{1} : Ar{0};
        Ar s1 = Ar{0};
#pragma unroll
        for (int t = 0; t < step; ++t) {
            const int k = UPPER ? kSB - 1 - t : t;
            const Ar a = T[i * kLD + k];
            if (t & 1) {
                s1 = fma_ar(-a, z[k], s1);
            } else {
                s0 = fma_ar(-a, z[k], s0);
            }
        }
        const Ar sum = s0 + s1;
        z[i] = UNIT ? sum : sum * inv_diag[i];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kSB; ++i) {
        T[i * kLD + lane] = z[i];
    }
}

// Thread layout for 128 x 128 tiles: 4 consecutive lanes share a row
// (row = tid / 4, seg = tid % 4) and a lane owns the columns
// 16*i + 4*seg + e (i < 8, e < 4).  Each load instruction therefore reads 16
// consecutive elements per row for 8 rows (full 32-byte sectors), and a row
// sum needs only TWO shuffle levels -- the shuffle unit handles one warp per
// cycle per SM, so the previous "lane = column" layout (five levels for each
// of 8 rows per warp) spent ~1300 cycles of every block step just shuffling.
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
    Ar* inv_diag = xsol + kB;                 // kB
    Ar* scratch = inv_diag + kB;              // kB, rehearsal right-hand side
    __shared__ unsigned k_shared;

    const int tid = threadIdx.x;
    const int lane = tid & (kWarp - 1);
    const int warp = tid >> 5;
    const int trow = tid >> 2;  // row of the tile this thread works on
    const int seg = tid & 3;

    if (tid == 0) {
        k_shared = atomicAdd(ticket, 1u);
    }
    __syncthreads();
    const std::int64_t k = k_shared;  // position in the solve order
    // development aid: per-CTA phase timestamps (SM cycles / global ns)
#define ACCBLAS_TRACE(slot, value)                  \
    if (trace != nullptr && tid == 0) {             \
        trace[k * 64 + (slot)] = (value);           \
    }
    ACCBLAS_TRACE(0, clock64());
    const std::int64_t nb = (n + kB - 1) / kB;
    const std::int64_t pb = UPPER ? nb - 1 - k : k;  // physical block row
    const std::int64_t r0 = pb * kB;
    const int bs = static_cast<int>((n - r0 < kB) ? (n - r0) : kB);

    // ---- diagonal tile -> shared memory (identity padding past the edge);
    //      all eight loads of a thread are issued before the first is used
    {
        constexpr int kIters = kB * (kB / kEPL) / kThreads;  // 8
        Quad<St> q[kIters];
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            const int idx = tid + it * kThreads;
            const int r = idx / (kB / kEPL);
            const int c = (idx % (kB / kEPL)) * kEPL;
            const int left = bs - c;
            const int valid =
                (r < bs) ? (left >= kEPL ? kEPL : (left > 0 ? left : 0)) : 0;
            const std::int64_t rr = (r < bs) ? r0 + r : r0;
            q[it] = load_quad<St, VECTOR>(A + rr * lda + r0 + c, valid);
        }
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            const int idx = tid + it * kThreads;
            const int r = idx / (kB / kEPL);
            const int c = (idx % (kB / kEPL)) * kEPL;
#pragma unroll
            for (int e = 0; e < kEPL; ++e) {
                const int cc = c + e;
                const bool in_tri = UPPER ? (cc >= r) : (cc <= r);
                Ar val;
                if (r < bs && cc < bs && in_tri && !(UNIT && r == cc)) {
                    val = to_ar<Ar, St>(q[it].v[e]);
                } else {
                    val = (r == cc) ? Ar{1} : Ar{0};
                }
                D[r * kLD + cc] = val;
            }
        }
    }
    if (tid < kB) {
        rhs[tid] = (tid < bs) ? to_ar<Ar, St>(x[(r0 + tid) * incx]) : Ar{0};
        scratch[tid] = rhs[tid];
        xsol[tid] = Ar{0};
    }
    __syncthreads();
    ACCBLAS_TRACE(1, clock64());
    if (warp < kNSB) {
        invert_subblock<Ar, UPPER, UNIT>(D + (warp * kSB) * kLD + warp * kSB,
                                         inv_diag + warp * kSB, lane);
    }
    ACCBLAS_TRACE(2, clock64());

    // Everything from here to the end of the diagonal solve runs TWICE:
    // pass 0 is a rehearsal on scratch data with nothing published.  A CTA
    // executes the reduction + diagonal-solve code exactly once for real, on
    // the critical path of the whole solve, and at that point none of it is in
    // the SM's instruction cache (measured: the cold run costs ~2x the warm
    // one).  The rehearsal happens while the CTA would be waiting for its
    // predecessors anyway.
    constexpr int NP = (sizeof(St) == 8) ? 2 : 1;  // panels per 128-col block
    constexpr int Q = 8 / NP;                      // quads per panel per thread
    const St* row_ptr;
    {
        std::int64_t r = r0 + trow;
        r = (r < n) ? r : n - 1;  // padded rows re-read a valid row
        row_ptr = A + r * lda + kEPL * seg;
    }
    auto load_panel = [&](std::int64_t jj, int p, Quad<St> (&dst)[Q]) {
        const std::int64_t pbj = UPPER ? nb - 1 - jj : jj;
        const std::int64_t c0 = pbj * kB + p * (16 * Q);
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            const std::int64_t col = c0 + 16 * i + kEPL * seg;
            const std::int64_t left = n - col;
            const int valid =
                left >= kEPL ? kEPL : (left > 0 ? static_cast<int>(left) : 0);
            dst[i] = load_quad<St, VECTOR>(row_ptr + c0 + 16 * i, valid);
        }
    };

#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const bool real = pass == 1;
        Ar* rhs_cur = real ? rhs : scratch;
        // fp32 arithmetic: one accumulator per quad slot, so a chain is no
        // longer than the reference's (one term per 32-column block)
        constexpr int NACC = std::is_same<Ar, float>::value ? 8 : 4;
        Ar acc[NACC];
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            acc[i] = Ar{};
        }
        const std::int64_t deps = real ? k : 0;

        // ---- off-diagonal blocks, in solve order; the loads of the next
        //      panel are in flight while the current one is consumed
        Quad<St> cur[Q];
        if (deps > 0) {
            load_panel(0, 0, cur);
        }
        int buf = 0;
        for (std::int64_t jj = 0; jj < deps; ++jj) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                Quad<St> nxt[Q];
                if (p + 1 < NP) {
                    load_panel(jj, p + 1, nxt);
                } else if (jj + 1 < deps) {
                    load_panel(jj + 1, 0, nxt);
                }
                // widen the panel BEFORE waiting for x: the 64-bit conversions
                // go through the XU pipe at 16 lanes/clk/SM (1024 cycles for a
                // 128x128 tile) and would otherwise sit on the critical path
                // between "x arrived" and "row sums ready"
                Ar cv[Q][kEPL];
#pragma unroll
                for (int i = 0; i < Q; ++i) {
#pragma unroll
                    for (int e = 0; e < kEPL; ++e) {
                        cv[i][e] = to_ar<Ar, St>(cur[i].v[e]);
                        pin_register(cv[i][e]);
                    }
                }
                if (p == 0) {
                    if (jj == deps - 1) {
                        ACCBLAS_TRACE(3, clock64());
                    }
                    if (warp == 0) {
                        const std::int64_t pbj = UPPER ? nb - 1 - jj : jj;
                        const std::int64_t pc = pbj * kB + lane * kEPL;
                        const std::int64_t left = n - pc;
                        const int valid =
                            left >= kEPL
                                ? kEPL
                                : (left > 0 ? static_cast<int>(left) : 0);
                        Ar v[kEPL];
                        bool ok;
                        do {
                            ok = true;
#pragma unroll
                            for (int e = 0; e < kEPL; ++e) {
                                if (e < valid) {
                                    v[e] = ld_volatile(xs + pc + e);
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
                        if (jj == deps - 1) {
                            ACCBLAS_TRACE(4, clock64());
                            ACCBLAS_TRACE(13, static_cast<long long>(
                                                  globaltimer_ns()));
                        }
                    }
                    __syncthreads();
                    if (jj == deps - 1) {
                        ACCBLAS_TRACE(5, clock64());
                    }
                }
                const Ar* xb = xcol + buf * kB + p * (16 * Q) + kEPL * seg;
                const bool tile_probe =
                    trace != nullptr && tid == 0 && jj == deps - 1 && p == NP - 1;
                if (tile_probe) {
                    Ar first = xb[0];
                    pin_register(first);
                    trace[k * 64 + 32] = clock64();
                }
#pragma unroll
                for (int i = 0; i < Q; ++i) {
#pragma unroll
                    for (int e = 0; e < kEPL; ++e) {
                        const int slot = (p * Q + i) % NACC;
                        acc[slot] = fma_ar(cv[i][e], xb[16 * i + e], acc[slot]);
                    }
                }
                if (tile_probe) {
#pragma unroll
                    for (int i = 0; i < NACC; ++i) {
                        pin_register(acc[i]);
                    }
                    trace[k * 64 + 33] = clock64();
                }
#pragma unroll
                for (int i = 0; i < Q; ++i) {
                    cur[i] = nxt[i];
                }
            }
            buf ^= 1;
        }
        {
#pragma unroll
            for (int width = NACC / 2; width > 0; width /= 2) {
#pragma unroll
                for (int i = 0; i < width; ++i) {
                    acc[i] += acc[i + width];
                }
            }
            Ar v = acc[0];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (seg == 0 && r0 + trow < n) {  // padded rows stay zero
                rhs_cur[trow] -= v;
            }
            if (real && trace != nullptr && tid == 0) {
                pin_register(v);
                trace[k * 64 + 34] = clock64();
            }
        }
        __syncthreads();
        if (real) {
            ACCBLAS_TRACE(6, clock64());
        }

        // ---- diagonal block: the four 32-wide sub-blocks in solve order.
        //      Warp group g (4 warps = the 128 threads whose rows lie in the
        //      g-th sub-block to be solved) multiplies by that sub-block's
        //      inverse and publishes; every later group then subtracts the new
        //      entries from its own rows.  Synchronisation is by NAMED barriers
        //      between exactly the warps involved (the solving group only
        //      arrives and moves on), not by CTA-wide barriers:
        //        id 1+g : "x of group g is in shared memory"  (group g arrives,
        //                 later groups wait)          128 * (4 - g) threads
        //        id 4+g : "rhs of group g is complete" (within group g)  128
        const int mem_sub = trow >> 5;
        const int grp = UPPER ? kNSB - 1 - mem_sub : mem_sub;
        const bool probe = real && trace != nullptr && seg == 0;
#pragma unroll 1
        for (int step = 0; step < kNSB; ++step) {
            const int s = UPPER ? kNSB - 1 - step : step;  // memory order
            // a thread owns the column pairs 32 s + 2 seg + 8 e + {0, 1}
            const int c2 = s * kSB + 2 * seg;
            if (grp == step) {
                if (step > 0) {
                    named_barrier_sync(4 + step, 4 * kWarp);
                }
                if (probe && (trow & 31) == 0) {
                    trace[k * 64 + 16 + 4 * step] = clock64();
                }
                const Pair<Ar>* Trow =
                    reinterpret_cast<const Pair<Ar>*>(D + trow * kLD + c2);
                const Pair<Ar>* v =
                    reinterpret_cast<const Pair<Ar>*>(rhs_cur + c2);
                Ar p0 = Ar{}, p1 = Ar{};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const Pair<Ar> t = Trow[4 * e];
                    const Pair<Ar> w = v[4 * e];
                    p0 = fma_ar(t.a, w.a, p0);
                    p1 = fma_ar(t.b, w.b, p1);
                }
                Ar sol = p0 + p1;
                sol += __shfl_xor_sync(0xffffffffu, sol, 1);
                sol += __shfl_xor_sync(0xffffffffu, sol, 2);
                if (seg == 0) {
                    // round through storage: later rows see what the
                    // accessor re-reads
                    const St stored = to_st<St, Ar>(sol);
                    const Ar back = to_ar<Ar, St>(stored);
                    xsol[trow] = back;
                    if (probe && (trow & 31) == 0) {
                        trace[k * 64 + 17 + 4 * step] = clock64();
                    }
                }
                if (step + 1 < kNSB) {
                    named_barrier_arrive(1 + step, 4 * kWarp * (kNSB - step));
                }
                if (real && seg == 0) {
                    // publish: progress vector first (the next block row spins
                    // on it), then the caller's x.  No barrier follows in this
                    // warp group, so the stores cost nothing here.
                    const std::int64_t gi = r0 + trow;
                    if (gi < n) {
                        const Ar val = xsol[trow];
                        st_volatile(xs + gi, Sentinel<Ar>::clean(val));
                        x[gi * incx] = to_st<St, Ar>(val);
                    }
                }
            } else if (grp > step) {
                named_barrier_sync(1 + step, 4 * kWarp * (kNSB - step));
                const Pair<Ar>* Drow =
                    reinterpret_cast<const Pair<Ar>*>(D + trow * kLD + c2);
                const Pair<Ar>* xv =
                    reinterpret_cast<const Pair<Ar>*>(xsol + c2);
                Ar p0 = Ar{}, p1 = Ar{};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const Pair<Ar> t = Drow[4 * e];
                    const Pair<Ar> w = xv[4 * e];
                    p0 = fma_ar(t.a, w.a, p0);
                    p1 = fma_ar(t.b, w.b, p1);
                }
                Ar sum = p0 + p1;
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                if (seg == 0) {
                    rhs_cur[trow] -= sum;
                }
                if (probe && grp == step + 1 && (trow & 31) == 0) {
                    trace[k * 64 + 18 + 4 * step] = clock64();
                }
            }
        }
        // both passes end with the whole CTA in step (the rehearsal must not
        // run into the real pass's barriers)
        __syncthreads();
        if (real) {
            ACCBLAS_TRACE(10, clock64());
        }
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
    const size_t smem = sizeof(Ar) * (kB * kLD + 6 * kB);
    // the opt-in is per device (and per instantiation)
    static bool configured[64] = {};
    int device = 0;
    ACCBLAS_CUDA(cudaGetDevice(&device));
    if (device < 0 || device >= 64 || !configured[device]) {
        ACCBLAS_CUDA(cudaFuncSetAttribute(
            kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
            static_cast<int>(smem)));
        if (device >= 0 && device < 64) {
            configured[device] = true;
        }
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
