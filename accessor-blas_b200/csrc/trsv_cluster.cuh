// TRSV, cluster kernel: blocked single-launch triangular solve whose critical
// path -- the hand-off of a solved block to the block row that needs it next
// -- runs through distributed shared memory instead of L2.
//
// Replaces kernel::acc_{lower,upper}_trsv / kernel::{lower,upper}_trsv +
// kernel::trsv_init (/root/reference/cuda/trsv_kernels.cuh:38-42,69-432,
// 527-893).  Same block algorithm and numerics as trsv.cu (128-row block rows,
// 32x32 sub-block inverses, x_g = Inv_g rhs_g - sum M(g,t) x_t, every solved
// entry rounded through the storage type before anyone consumes it); what
// changes is how the pieces talk to each other:
//
//   * Block rows are handed out to thread-block CLUSTERS of up to 8 CTAs in
//     solve order (one atomic ticket per cluster, rank inside the cluster =
//     position inside the group of 8 block rows).  A CTA only waits on block
//     rows with an earlier ticket or a lower rank of its own cluster, all of
//     which are resident or finished: no deadlock whatever the residency.
//   * The block row that comes next sits in the same cluster 7 times out of
//     8.  The solving CTA pushes every 32-entry sub-block of its solution
//     straight into that CTA's shared memory with st.async (one transaction
//     carries the value and completes the receiver's mbarrier transaction
//     count): ~220 cycles SM to SM, against a store to L2, a polling round
//     trip and a CTA barrier (~1700 cycles measured in trsv.cu).  The
//     receiver sleeps on four mbarriers, one per sub-block, and starts on the
//     32 columns that have arrived; after the last sub-block only 8 FMAs per
//     thread, two shuffle levels and the diagonal solve remain.  The solution
//     is pushed to the NEXT TWO block rows: the second one streams that panel
//     like any other, but does not have to wait for it to come round through
//     L2 (that detour -- a poll of the progress vector, a staged copy, a
//     panel, the tail -- was a second critical path three block rows long).
//   * Everything else of a block row -- the panels left of the last one -- is
//     streamed GEMV-style by 16 independent warps: a warp owns 8 rows, a lane
//     4 consecutive columns of a 128-column panel (16-byte L1-bypassing
//     loads, one 512-byte row segment per warp request), the next chunk's
//     loads are in flight while the current one is converted and multiplied,
//     and the lane's slice of x sits in registers.  No CTA-wide barrier in the
//     loop: whole x blocks are staged from the progress vector (the solution
//     itself, NaN sentinel = not solved yet) into a 4-deep shared-memory ring
//     with full/empty mbarriers per slot, two blocks ahead of the fastest
//     warp, by whichever warp gets there first (a claim counter in shared
//     memory; a dedicated 17th warp would put five warps on one scheduler and
//     cap every thread at 96 registers).  The row sums are folded across the
//     warp once, at the end (a 9-shuffle reduce-scatter that leaves row q's
//     sum with the four lanes that own row q in the layout of the last panel
//     and the diagonal solve).
//   * The diagonal tile, its sub-block inverses and products, the rehearsal of
//     the reduction + solve code and the chain of named barriers between the
//     warp groups are those of trsv.cu.
// The last CTA re-arms the workspace (sentinels, ticket) for the next call.
#pragma once

#include "common.cuh"
#include "trsv_common.cuh"
#include "tuning.h"

namespace accblas {
namespace {

using namespace trsv_detail;

constexpr int kRing = 4;        // x blocks staged in shared memory
constexpr int kLookAhead = 2;   // blocks staged ahead of the warp that asks
constexpr int kTail = 2;        // blocks before this one received by push
constexpr int kMaxCluster = 8;

// ---------------------------------------------------------------------------
// cluster, mbarrier and DSMEM primitives
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned cluster_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_size()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile(
        "barrier.cluster.arrive.release.aligned;\n"
        "barrier.cluster.wait.acquire.aligned;\n" ::
            : "memory");
}
__device__ __forceinline__ unsigned smem_addr(const void* p)
{
    return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
// shared::cluster address of `local` (a shared::cta address) in CTA `rank`
__device__ __forceinline__ unsigned map_to_rank(unsigned local, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                 : "=r"(r)
                 : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ unsigned ld_cluster_u32(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar),
                 "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar,
                                                      unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                     bar),
                 "r"(bytes)
                 : "memory");
}
// CTA-scope acquire: data written by threads of this CTA
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// cluster-scope acquire: data written by st.async of another CTA
__device__ __forceinline__ void mbar_wait_cluster(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// one value into (possibly another CTA's) shared memory; the same transaction
// completes sizeof(value) bytes on the mbarrier next to it
__device__ __forceinline__ void st_async(unsigned addr, double v, unsigned bar)
{
    asm volatile(
        "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 "
        "[%0], %1, [%2];" ::"r"(addr),
        "l"(__double_as_longlong(v)), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void st_async(unsigned addr, float v, unsigned bar)
{
    asm volatile(
        "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 "
        "[%0], %1, [%2];" ::"r"(addr),
        "r"(__float_as_uint(v)), "r"(bar)
        : "memory");
}

// ---------------------------------------------------------------------------
// streaming layout: a lane's slice of one row of one chunk
//   fp64: 2 elements (16 bytes), chunk = 64 columns, 2 chunks per panel
//   fp32: 4 elements (16 bytes), chunk = 128 columns
//   fp16: 4 elements ( 8 bytes), chunk = 128 columns
// RAW words, taken apart only when they are consumed (see Quad in
// trsv_common.cuh for why).
// ---------------------------------------------------------------------------
template <typename St>
struct Span {
    static constexpr int kElems = sizeof(St) == 8 ? 2 : 4;
    static constexpr int kWords = kElems * static_cast<int>(sizeof(St)) / 4;
    unsigned w[kWords];

    __device__ __forceinline__ void pin()
    {
#pragma unroll
        for (int i = 0; i < kWords; ++i) {
            asm volatile("" : "+r"(w[i]));
        }
    }
    template <typename Ar>
    __device__ __forceinline__ Ar get(int e) const
    {
        if constexpr (sizeof(St) == 8) {
            return to_ar<Ar, double>(__hiloint2double(
                static_cast<int>(w[2 * e + 1]), static_cast<int>(w[2 * e])));
        } else if constexpr (sizeof(St) == 4) {
            return to_ar<Ar, float>(__uint_as_float(w[e]));
        } else {
            const unsigned short bits = static_cast<unsigned short>(
                (e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xffffu));
            return to_ar<Ar, __half>(__ushort_as_half(bits));
        }
    }
};

// VW = vector width the rows allow (16 / 8 / 0 = scalar), as in trsv.cu
template <typename St, int VW>
__device__ __forceinline__ Span<St> load_span(const St* p, int valid)
{
    Span<St> s;
    constexpr int E = Span<St>::kElems;
    constexpr int W = Span<St>::kWords;
    if (VW != 0 && valid == E) {
        if (W == 4 && VW == 16) {
            const uint4 a = ldg_stream_128(p);
            s.w[0] = a.x;
            s.w[1] = a.y;
            s.w[2 % W] = a.z;
            s.w[3 % W] = a.w;
        } else {
#pragma unroll
            for (int i = 0; i < W / 2; ++i) {
                const uint2 a =
                    ldg_stream_64(reinterpret_cast<const char*>(p) + 8 * i);
                s.w[2 * i] = a.x;
                s.w[2 * i + 1] = a.y;
            }
        }
    } else {
        Quad<St> q;  // reuse the element-wise packing of the quad
#pragma unroll
        for (int i = 0; i < Quad<St>::kWords; ++i) {
            q.w[i] = 0u;
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if (e < valid) {
                q.set(e, p[e]);
            }
        }
#pragma unroll
        for (int i = 0; i < W; ++i) {
            s.w[i] = q.w[i];
        }
    }
    return s;
}

template <typename Ar>
__device__ __forceinline__ Ar shfl_xor(Ar v, int mask)
{
    return __shfl_xor_sync(0xffffffffu, v, mask);
}

// v[q] = this lane's partial sum of row q (q < 8) -> returns the sum over all
// 32 lanes of row (lane >> 2): the four lanes of a quad end with the same
// value.  Fixed order.
template <typename Ar>
__device__ __forceinline__ Ar reduce_scatter_rows(const Ar (&v)[8], int lane)
{
    Ar t[4];
    {
        const bool hi = (lane & 16) != 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const Ar send = hi ? v[q] : v[q + 4];
            const Ar keep = hi ? v[q + 4] : v[q];
            t[q] = keep + shfl_xor(send, 16);
        }
    }
    Ar u[2];
    {
        const bool hi = (lane & 8) != 0;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const Ar send = hi ? t[q] : t[q + 2];
            const Ar keep = hi ? t[q + 2] : t[q];
            u[q] = keep + shfl_xor(send, 8);
        }
    }
    Ar w;
    {
        const bool hi = (lane & 4) != 0;
        const Ar send = hi ? u[0] : u[1];
        const Ar keep = hi ? u[1] : u[0];
        w = keep + shfl_xor(send, 4);
    }
    w += shfl_xor(w, 1);
    w += shfl_xor(w, 2);
    return w;
}

// Cold path kept out of line: inlined into the unrolled tail it costs registers
// the hot path needs (ptxas spilled around it).
// A block solved by ANOTHER cluster reaches this CTA through L2 only: one warp
// polls the progress vector and delivers into this CTA's own push buffer with
// the st.async a neighbour would have used.  Delivers the block's sub-blocks
// in arrival order up to and including arrival `upto`; `next` = arrivals
// delivered so far (returned updated).  ALL pending sub-blocks are polled in
// the same round (their loads are in flight together), so consecutive
// sub-blocks cost one L2 round trip between them only if they really are a
// round trip apart -- polled one after the other, the last of the four was
// delivered four round trips after the first became visible.
template <typename Ar, bool UPPER>
__device__ __noinline__ int deliver_from_l2(const Ar* xs_block, int valid,
                                            int next, int upto,
                                            unsigned dst_block,
                                            unsigned bar_block, int lane)
{
    while (next <= upto) {
        Ar v[kNSB];
#pragma unroll
        for (int u = 0; u < kNSB; ++u) {
            const int idx = (UPPER ? kNSB - 1 - u : u) * kSB + lane;
            v[u] = Ar{0};
            if (u >= next && idx < valid) {
                v[u] = ld_volatile(xs_block + idx);
            }
        }
#pragma unroll
        for (int u = 0; u < kNSB; ++u) {
            if (u == next &&
                __all_sync(0xffffffffu, !Sentinel<Ar>::is(v[u]))) {
                const int idx = (UPPER ? kNSB - 1 - u : u) * kSB + lane;
                st_async(dst_block + idx * static_cast<unsigned>(sizeof(Ar)),
                         v[u], bar_block + 8 * u);
                ++next;
            }
        }
    }
    return next;
}

template <typename St, typename Ar, bool UPPER, bool UNIT, int VW, bool TRACE>
__global__ __launch_bounds__(kThreads, 1) void trsv_cluster_kernel(
    std::int64_t n, const St* __restrict__ A, std::int64_t lda,
    St* __restrict__ x, std::int64_t incx, Ar* xs,
    unsigned* __restrict__ ticket, long long* __restrict__ trace_arg,
    int l2_ahead)
{
    long long* const trace = TRACE ? trace_arg : nullptr;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ar* D = reinterpret_cast<Ar*>(smem_raw);  // kB x kLD
    Ar* ring = D + kB * kLD;                  // kRing x kB, staged x blocks
    Ar* xpush = ring + kRing * kB;            // kTail x kB, the blocks solved last
    Ar* rhs = xpush + kTail * kB;             // kB
    Ar* xsol = rhs + kB;                      // kB
    Ar* inv_diag = xsol + kB;                 // kB
    Ar* scratch = inv_diag + kB;              // kB, rehearsal right-hand side
    __shared__ __align__(8) unsigned long long bars[2 * kRing + kTail * kNSB];
    __shared__ unsigned ticket_s;
    __shared__ unsigned next_fetch;  // first x block nobody has claimed yet

    const int tid = threadIdx.x;
    const int lane = tid & (kWarp - 1);
    const int warp = tid >> 5;
    const unsigned crank = cluster_rank();
    const unsigned csize = cluster_size();
    const unsigned bar_full = smem_addr(&bars[0]);
    const unsigned bar_empty = smem_addr(&bars[kRing]);
    const unsigned bar_push = smem_addr(&bars[2 * kRing]);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kRing; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, kThreads / kWarp);
        }
#pragma unroll
        for (int t = 0; t < kTail * kNSB; ++t) {
            mbar_init(bar_push + 8 * t, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int t = 0; t < kTail * kNSB; ++t) {
            // the one arrival; the phase completes when the 32 values are in
            mbar_arrive_expect_tx(bar_push + 8 * t,
                                  kSB * static_cast<unsigned>(sizeof(Ar)));
        }
        next_fetch = 0u;
        if (crank == 0) {
            ticket_s = atomicAdd(ticket, 1u);
        }
    }
    __syncthreads();
    cluster_sync_all();  // barriers armed and the ticket visible cluster-wide
    const std::int64_t k =
        static_cast<std::int64_t>(
            ld_cluster_u32(map_to_rank(smem_addr(&ticket_s), 0))) *
            csize +
        crank;  // position in the solve order
    const std::int64_t nb = (n + kB - 1) / kB;

#define ACCBLAS_TRACE(slot, value)                           \
    if (TRACE && trace != nullptr && tid == 0) {             \
        trace[k * 64 + (slot)] = (value);                    \
    }

    if (k < nb) {
        const std::int64_t pb = UPPER ? nb - 1 - k : k;  // physical block row
        const std::int64_t r0 = pb * kB;
        const int bs = static_cast<int>((n - r0 < kB) ? (n - r0) : kB);
        const std::int64_t deps = k;  // block rows solved before this one

        {
            const int trow = tid >> 2;  // row of the tile in the quad layout
            const int seg = tid & 3;
            ACCBLAS_TRACE(0, clock64());
            // ---- diagonal tile -> shared memory (identity padding past the
            //      edge); all eight loads of a thread are issued before the
            //      first is used
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
                        (r < bs) ? (left >= kEPL ? kEPL : (left > 0 ? left : 0))
                                 : 0;
                    const std::int64_t rr = (r < bs) ? r0 + r : r0;
                    q[it] = load_quad<St, VW>(A + rr * lda + r0 + c, valid);
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
                            val = q[it].template get<Ar>(e);
                        } else {
                            val = (r == cc) ? Ar{1} : Ar{0};
                        }
                        D[r * kLD + cc] = val;
                    }
                }
            }
            if (tid < kB) {
                rhs[tid] =
                    (tid < bs) ? to_ar<Ar, St>(x[(r0 + tid) * incx]) : Ar{0};
                scratch[tid] = rhs[tid];
                xsol[tid] = Ar{0};
            }
            __syncthreads();
            ACCBLAS_TRACE(1, clock64());
            if (warp < kNSB && warp * kSB < bs) {
                invert_subblock<Ar, UPPER, UNIT>(
                    D + (warp * kSB) * kLD + warp * kSB, inv_diag + warp * kSB,
                    lane);
            }
            __syncthreads();
            // ---- M(g,t) = Inv_g * D(g,t), in place (see trsv.cu)
            {
                const int i = tid >> 4;         // row inside the 32 x 32 block
                const int jp = (tid & 15) * 2;  // column pair
                Ar out[6][2];
                int slot = 0;
#pragma unroll
                for (int g = 1; g < kNSB; ++g) {
#pragma unroll
                    for (int t = 0; t < g; ++t) {
                        const int mg = UPPER ? kNSB - 1 - g : g;
                        const int mt = UPPER ? kNSB - 1 - t : t;
                        const Ar* inv = D + (mg * kSB + i) * kLD + mg * kSB;
                        const Ar* blk = D + (mg * kSB) * kLD + mt * kSB + jp;
                        Ar o0 = Ar{}, o1 = Ar{};
                        if (mg * kSB < bs && mt * kSB < bs) {
#pragma unroll 8
                            for (int kk = 0; kk < kSB; ++kk) {
                                const Ar a = inv[kk];
                                const Pair<Ar> d =
                                    *reinterpret_cast<const Pair<Ar>*>(
                                        blk + kk * kLD);
                                o0 = fma_ar(a, d.a, o0);
                                o1 = fma_ar(a, d.b, o1);
                            }
                        }
                        out[slot][0] = o0;
                        out[slot][1] = o1;
                        ++slot;
                    }
                }
                __syncthreads();
                slot = 0;
#pragma unroll
                for (int g = 1; g < kNSB; ++g) {
#pragma unroll
                    for (int t = 0; t < g; ++t) {
                        const int mg = UPPER ? kNSB - 1 - g : g;
                        const int mt = UPPER ? kNSB - 1 - t : t;
                        Pair<Ar> o;
                        o.a = out[slot][0];
                        o.b = out[slot][1];
                        *reinterpret_cast<Pair<Ar>*>(
                            D + (mg * kSB + i) * kLD + mt * kSB + jp) = o;
                        ++slot;
                    }
                }
            }
            __syncthreads();
            ACCBLAS_TRACE(2, clock64());

            const int mem_sub = trow >> 5;
            const int grp = UPPER ? kNSB - 1 - mem_sub : mem_sub;  // solve index
            // streaming geometry
            constexpr int EPL = Span<St>::kElems;  // elements per lane and row
            constexpr int CW = kWarp * EPL;        // columns per chunk
            constexpr int CPP = kB / CW;           // chunks per panel
            constexpr int RW = kB / (kThreads / kWarp);  // rows per warp: 8
            // fp32 arithmetic: two accumulators per row, chains as short as
            // trsv.cu's
            constexpr int NA = std::is_same<Ar, float>::value ? 2 : 1;
#pragma unroll 1
            // (the first block row of the solve order has nobody to wait for:
            // no rehearsal)
            for (int pass = (k == 0 ? 1 : 0); pass < 2; ++pass) {
                const bool real = pass == 1;
                Ar* rhs_cur = real ? rhs : scratch;
                // streamed panels: blocks 0 .. deps - 2 of the solve order.  The
                // x of the last of them (the block solved two positions back)
                // arrives by push, like the tail's; the others come through
                // the ring from L2
                const std::int64_t panels = real ? (deps > 1 ? deps - 1 : 0) : 0;
                const std::int64_t ring_panels = panels > 0 ? panels - 1 : 0;
                const std::int64_t items = panels * CPP;
                // The streaming geometry is (re)computed inside the pass: kept
                // outside, it stays live through the tail and the solve of
                // BOTH passes, where registers are what the kernel is short
                // of.  `zero` is opaque to the compiler, so nothing here can be
                // hoisted out of the loop.
                int zero = 0;
                asm volatile("" : "+r"(zero));
                const St* rowbase;
                int rmax;  // rows past the end of the matrix re-read a valid row
                {
                    std::int64_t rw = r0 + RW * warp + zero;
                    const std::int64_t room = n - 1 - rw;
                    rmax = room >= RW - 1
                               ? RW - 1
                               : (room > 0 ? static_cast<int>(room) : 0);
                    rw = rw < n ? rw : n - 1;
                    rowbase = A + rw * lda + EPL * lane;
                }
                const std::int64_t tail_prefetch_at =
                    panels > 6 ? panels - 6 : 0;
                const unsigned self_push = map_to_rank(smem_addr(xpush), crank);
                const unsigned self_bar = map_to_rank(bar_push, crank);
                const int group_blocks =
                    l2_ahead + zero > 0
                        ? max(1, l2_ahead / (kB * static_cast<int>(sizeof(St))))
                        : 0;

                Ar acc[RW][NA];
#pragma unroll
                for (int r = 0; r < RW; ++r) {
#pragma unroll
                    for (int a = 0; a < NA; ++a) {
                        acc[r][a] = Ar{};
                    }
                }

                auto load_chunk = [&](std::int64_t it, Span<St> (&dst)[RW]) {
                    const std::int64_t jj = it / CPP;
                    const int c = static_cast<int>(it % CPP);
                    const std::int64_t pbj = UPPER ? nb - 1 - jj : jj;
                    const std::int64_t col0 = pbj * kB + c * CW;
                    const std::int64_t left = n - (col0 + EPL * lane);
                    const int valid =
                        left >= EPL ? EPL
                                    : (left > 0 ? static_cast<int>(left) : 0);
#pragma unroll
                    for (int r = 0; r < RW; ++r) {
                        const int rr = r < rmax ? r : rmax;
                        dst[r] = load_span<St, VW>(rowbase + rr * lda + col0,
                                                   valid);
                    }
                };
                // L2 prefetch of a group of panels for this warp's 8 rows:
                // lane -> row (lane >> 2), lines (lane & 3), +4, ...
                auto l2_prefetch_group = [&](std::int64_t j0) {
                    // (the panel of the block solved last is read in the quad
                    // layout, after the loop: it belongs to the groups too)
                    const std::int64_t limit = real ? deps : 0;
                    std::int64_t j1 = j0 + group_blocks;
                    j1 = j1 < limit ? j1 : limit;
                    if (j0 >= j1) {
                        return;
                    }
                    const std::int64_t c_lo = (UPPER ? nb - j1 : j0) * kB;
                    std::int64_t c_hi = (UPPER ? nb - j0 : j1) * kB;
                    c_hi = c_hi < n ? c_hi : n;
                    constexpr int kLineElems = 128 / static_cast<int>(sizeof(St));
                    const int r = lane >> 2;
                    const St* row_l2 =
                        rowbase - EPL * lane + (r < rmax ? r : rmax) * lda;
                    for (std::int64_t c = c_lo + (lane & 3) * kLineElems;
                         c < c_hi; c += 4 * kLineElems) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(row_l2 + c));
                    }
                };
                // ---- staging of x blocks: progress vector (L2) -> ring slot.
                //      Blocks are claimed strictly in order (`next_fetch`).
                //      * A warp that NEEDS block jj and finds it unclaimed
                //        claims it and polls until it is there: it would have
                //        to wait for it anyway.
                //      * Ahead of need, block b is looked at once -- without
                //        waiting -- by warp b % 16 when that warp starts a
                //        panel up to kLookAhead blocks earlier, and staged only
                //        if it has been published completely.  Never is a warp
                //        parked on a block of the future while panels whose x
                //        is already there are waiting for it (that cost a
                //        caught-up CTA three panels of delay per block).
                auto load_block = [&](std::int64_t blk, Ar (&v)[4]) {
                    const std::int64_t pbj = UPPER ? nb - 1 - blk : blk;
                    const std::int64_t base = pbj * kB + 4 * lane;
                    bool missing = false;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (base + e < n &&
                            (Sentinel<Ar>::is(v[e]))) {
                            v[e] = ld_volatile(xs + base + e);
                            missing = missing || Sentinel<Ar>::is(v[e]);
                        }
                    }
                    return __any_sync(0xffffffffu, missing);
                };
                auto init_block = [&](std::int64_t blk, Ar (&v)[4]) {
                    const std::int64_t pbj = UPPER ? nb - 1 - blk : blk;
                    const std::int64_t base = pbj * kB + 4 * lane;
                    Ar sentinel;
                    memset(&sentinel, 0xff, sizeof(Ar));
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        v[e] = (base + e < n) ? sentinel : Ar{0};
                    }
                };
                auto publish_block = [&](std::int64_t blk, const Ar (&v)[4]) {
                    const int slot = static_cast<int>(blk % kRing);
                    const unsigned round = static_cast<unsigned>(blk / kRing);
                    if (round > 0) {
                        // every warp is past the block that used this slot
                        mbar_wait(bar_empty + 8 * slot, (round - 1) & 1u);
                    }
                    Ar* dst = ring + slot * kB + 4 * lane;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        dst[e] = v[e];
                    }
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(bar_full + 8 * slot);
                    }
                };
                auto try_claim = [&](std::int64_t blk) {
                    unsigned won = 0u;
                    if (lane == 0) {
                        won = atomicCAS(&next_fetch, static_cast<unsigned>(blk),
                                        static_cast<unsigned>(blk) + 1u) ==
                              static_cast<unsigned>(blk);
                    }
                    return __shfl_sync(0xffffffffu, won, 0) != 0u;
                };
                auto first_unclaimed = [&]() {
                    unsigned cur = 0u;
                    if (lane == 0) {
                        cur = *reinterpret_cast<volatile unsigned*>(&next_fetch);
                    }
                    return static_cast<std::int64_t>(
                        __shfl_sync(0xffffffffu, cur, 0));
                };
                // block jj is needed now
                auto ensure_staged = [&](std::int64_t jj) {
                    for (;;) {
                        const std::int64_t cur = first_unclaimed();
                        if (cur > jj) {
                            return;  // claimed by somebody: wait on `full`
                        }
                        if (try_claim(cur)) {
                            Ar v[4];
                            init_block(cur, v);
                            while (load_block(cur, v)) {
                            }
                            publish_block(cur, v);
                        }
                    }
                };
                // look once at the next unclaimed block if it is this warp's
                auto stage_ahead = [&](std::int64_t jj) {
                    const std::int64_t cur = first_unclaimed();
                    if (cur > jj && cur <= jj + kLookAhead && cur < ring_panels &&
                        (cur & 15) == warp) {
                        Ar v[4];
                        init_block(cur, v);
                        if (!load_block(cur, v) && try_claim(cur)) {
                            publish_block(cur, v);
                        }
                    }
                };
                auto consume = [&](std::int64_t it, Span<St> (&buf)[RW]) {
                    const std::int64_t jj = it / CPP;
                    const int c = static_cast<int>(it % CPP);
                    const int slot = static_cast<int>(jj % kRing);
                    if (c == 0) {
                        if (jj == tail_prefetch_at) {
                            // the tiles of the kTail blocks solved last are
                            // read in the quad layout, two sub-blocks at a
                            // time, with next to nothing in flight: they must
                            // come out of L2, not out of DRAM
                            const std::int64_t c_lo =
                                (UPPER ? pb + 1 : pb - kTail) * kB;
                            std::int64_t c_hi = c_lo + kTail * kB;
                            c_hi = c_hi < n ? c_hi : n;
                            constexpr int kLineElems =
                                128 / static_cast<int>(sizeof(St));
                            const int r = lane >> 2;
                            const St* row_l2 = rowbase - EPL * lane +
                                               (r < rmax ? r : rmax) * lda;
                            for (std::int64_t cc =
                                     (c_lo > 0 ? c_lo : 0) + (lane & 3) * kLineElems;
                                 cc < c_hi; cc += 4 * kLineElems) {
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(
                                    row_l2 + cc));
                            }
                        }
                        if (jj < ring_panels) {
                            ensure_staged(jj);
                            stage_ahead(jj);
                            mbar_wait(bar_full + 8 * slot,
                                      static_cast<unsigned>(jj / kRing) & 1u);
                            if (group_blocks > 0 && jj % group_blocks == 0) {
                                l2_prefetch_group(jj + group_blocks);
                            }
                        } else {
                            // the block solved two positions back: pushed by
                            // the CTA of rank - 2 -- or, when that one lives in
                            // another cluster, fetched from L2 by warp 0
                            if (crank < 2u && warp == 0) {
                                const std::int64_t pbl = UPPER ? pb + 2 : pb - 2;
                                const std::int64_t room = n - pbl * kB;
                                deliver_from_l2<Ar, UPPER>(
                                    xs + pbl * kB,
                                    room < kB ? static_cast<int>(room) : kB, 0,
                                    kNSB - 1,
                                    self_push + kB * static_cast<unsigned>(
                                                         sizeof(Ar)),
                                    self_bar + 8 * kNSB, lane);
                            }
#pragma unroll
                            for (int t = 0; t < kNSB; ++t) {
                                mbar_wait_cluster(bar_push + 8 * (kNSB + t), 0u);
                            }
                        }
                    }
                    Ar xv[EPL];
                    {
                        const Ar* xb =
                            (jj < ring_panels ? ring + slot * kB : xpush + kB) +
                            c * CW + EPL * lane;
#pragma unroll
                        for (int e = 0; e < EPL; e += 2) {
                            const Pair<Ar> p2 =
                                *reinterpret_cast<const Pair<Ar>*>(xb + e);
                            xv[e] = p2.a;
                            xv[e + 1] = p2.b;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < RW; ++r) {
                        buf[r].pin();
#pragma unroll
                        for (int e = 0; e < EPL; ++e) {
                            acc[r][e % NA] =
                                fma_ar(buf[r].template get<Ar>(e), xv[e],
                                       acc[r][e % NA]);
                        }
                    }
                    if (c == CPP - 1 && jj < ring_panels) {
                        // the slot may be refilled once all 16 warps are past it
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(bar_empty + 8 * slot);
                        }
                    }
                };

                // ---- the tile of the block solved LAST is read in the quad
                //      layout, sub-block by sub-block as the pushes arrive.  A
                //      thread's share of a 32-column sub-block is two quads.  The
                //      first two sub-blocks are requested while the last
                //      streamed chunk is still being waited for (its "next
                //      chunk" buffer is free by then), the other two as soon as
                //      the streaming buffers are dead: nothing of this tile's
                //      latency may end up behind the arrival of x.  (A rolling
                //      window of two sub-blocks, refilled after each arrival,
                //      exposed an L2 round trip per pair whenever the CTA was
                //      not early.)
                const bool has_tail = real && deps > 0;
                Quad<St> wq[kNSB][2];
                // t-th sub-block (in arrival order) of the last block's tile
                // (lower triangle: the last block's tile never reaches past the
                // end of the matrix; upper triangle: only for the second block
                // row of the solve order when n is not a multiple of 128)
                const std::int64_t pb_last = UPPER ? pb + 1 : pb - 1;
                const bool tail_edge = has_tail && (pb_last + 1) * kB > n;
                auto load_sub = [&](int t, Quad<St> (&dst)[2]) {
                    const int sbm = UPPER ? kNSB - 1 - t : t;
                    std::int64_t r = r0 + trow;
                    r = (r < n) ? r : n - 1;
                    const St* src =
                        A + r * lda + pb_last * kB + 32 * sbm + kEPL * seg;
                    dst[0] = load_quad<St, VW>(src, kEPL);
                    dst[1] = load_quad<St, VW>(src + 16, kEPL);
                };
                auto load_tail_window = [&]() {
                    if (has_tail && !tail_edge) {
                        load_sub(0, wq[0]);
                        load_sub(1, wq[1]);
                    }
                };

                // ---- streamed panels: register double buffer.  The last one or
                //      two chunks are peeled off the loop so that the tail
                //      window can be requested right before the LAST chunk is
                //      waited for without being live inside the loop (inside it
                //      the window's registers turned into spill traffic as
                //      large as the matrix stream itself for fp64 storage).
                {
                    Span<St> buf_a[RW];
                    Span<St> buf_b[RW];
                    if (items > 0) {
                        load_chunk(0, buf_a);
                        if (group_blocks > 0) {
                            l2_prefetch_group(0);
                        }
                    }
                    std::int64_t it = 0;
#pragma unroll 1
                    for (; it + 2 < items; it += 2) {
                        load_chunk(it + 1, buf_b);
                        consume(it, buf_a);
                        load_chunk(it + 2, buf_a);
                        consume(it + 1, buf_b);
                    }
                    // chunk `it` (if any) is in buf_a
                    if (items - it == 2) {
                        load_chunk(it + 1, buf_b);
                        consume(it, buf_a);
                        load_tail_window();
                        consume(it + 1, buf_b);
                    } else if (items - it == 1) {
                        load_tail_window();
                        consume(it, buf_a);
                    } else {
                        load_tail_window();
                    }
                }
                if (real) {
                    ACCBLAS_TRACE(3, clock64());
                }
                if (has_tail && !tail_edge) {
                    load_sub(2, wq[2]);
                    load_sub(3, wq[3]);
                }
                {
                    Ar v[RW];
#pragma unroll
                    for (int r = 0; r < RW; ++r) {
                        v[r] = acc[r][0];
#pragma unroll
                        for (int a = 1; a < NA; ++a) {
                            v[r] += acc[r][a];
                        }
                    }
                    const Ar streamed = reduce_scatter_rows(v, lane);
                    if (seg == 0 && r0 + trow < n) {  // padded rows stay zero
                        rhs_cur[trow] -= streamed;
                    }
                }
                if (tail_edge) {
                    // the rare tile that straddles the end of the matrix:
                    // whole block at once, element by element
                    if (crank == 0 && warp == 0) {
                        const std::int64_t room = n - pb_last * kB;
                        deliver_from_l2<Ar, UPPER>(
                            xs + pb_last * kB,
                            room < kB ? static_cast<int>(room) : kB, 0, kNSB - 1,
                            self_push, self_bar, lane);
                    }
#pragma unroll 1
                    for (int t = 0; t < kNSB; ++t) {
                        mbar_wait_cluster(bar_push + 8 * t, 0u);
                    }
                    std::int64_t r = r0 + trow;
                    r = (r < n) ? r : n - 1;
                    const St* src = A + r * lda + pb_last * kB;
                    const int cols = static_cast<int>(n - pb_last * kB);
                    Ar v = Ar{};
#pragma unroll 1
                    for (int c = seg; c < cols; c += 4) {
                        v = fma_ar(to_ar<Ar, St>(src[c]), xpush[c], v);
                    }
                    v += shfl_xor(v, 1);
                    v += shfl_xor(v, 2);
                    if (seg == 0 && r0 + trow < n) {
                        rhs_cur[trow] -= v;
                    }
                } else if (has_tail) {
                    Ar a0 = Ar{}, a1 = Ar{};
                    int l2_next = 0;  // arrivals delivered by warp 0
#pragma unroll
                    for (int t = 0; t < kNSB; ++t) {
                        const int sbm = UPPER ? kNSB - 1 - t : t;
                        // a predecessor outside this cluster publishes through
                        // L2 only: warp 0 polls the progress vector and
                        // delivers with the st.async a neighbour would have used
                        if (crank == 0 && warp == 0) {
                            const std::int64_t room = n - pb_last * kB;
                            l2_next = deliver_from_l2<Ar, UPPER>(
                                xs + pb_last * kB,
                                room < kB ? static_cast<int>(room) : kB, l2_next,
                                t, self_push, self_bar, lane);
                        }
                        // The very last sub-block is widened BEFORE the wait
                        // (its conversions would sit on the critical path); the
                        // others are widened after their own wait, in the
                        // shadow of the next one.
                        const bool ahead = t == kNSB - 1;
                        Ar cv[2][kEPL];
                        if (ahead) {
#pragma unroll
                            for (int ii = 0; ii < 2; ++ii) {
                                wq[t][ii].pin();
#pragma unroll
                                for (int e = 0; e < kEPL; ++e) {
                                    cv[ii][e] = wq[t][ii].template get<Ar>(e);
                                    pin_register(cv[ii][e]);
                                }
                            }
                        }
                        mbar_wait_cluster(bar_push + 8 * t, 0u);
                        if (TRACE && trace != nullptr && tid == 0) {
                            trace[k * 64 + 44 + t] =
                                static_cast<long long>(globaltimer_ns());
                        }
                        const Ar* xb = xpush + kEPL * seg;
                        if (!ahead) {
#pragma unroll
                            for (int ii = 0; ii < 2; ++ii) {
                                wq[t][ii].pin();
#pragma unroll
                                for (int e = 0; e < kEPL; ++e) {
                                    cv[ii][e] = wq[t][ii].template get<Ar>(e);
                                }
                            }
                        }
#pragma unroll
                        for (int ii = 0; ii < 2; ++ii) {
                            const int i = 2 * sbm + ii;
#pragma unroll
                            for (int e = 0; e < kEPL; e += 2) {
                                const Pair<Ar> p2 =
                                    *reinterpret_cast<const Pair<Ar>*>(
                                        xb + 16 * i + e);
                                a0 = fma_ar(cv[ii][e], p2.a, a0);
                                a1 = fma_ar(cv[ii][e + 1], p2.b, a1);
                            }
                        }
                    }
                    ACCBLAS_TRACE(4, clock64());
                    ACCBLAS_TRACE(13, static_cast<long long>(globaltimer_ns()));
                    Ar v = a0 + a1;
                    v += shfl_xor(v, 1);
                    v += shfl_xor(v, 2);
                    if (seg == 0 && r0 + trow < n) {
                        rhs_cur[trow] -= v;
                    }
                }
                // "rhs of my sub-block is complete": its 32 rows belong to
                // the four warps of this group only
                named_barrier_sync(1 + grp, 4 * kWarp);
                if (real) {
                    ACCBLAS_TRACE(6, clock64());
                }

                // where the solution of this row is pushed: this thread's slot
                // in the push buffer of the next kTail CTAs of the cluster
                bool push_to[kTail];
                unsigned push_dst[kTail];
                unsigned push_bar[kTail];
#pragma unroll
                for (int d = 1; d <= kTail; ++d) {
                    push_to[d - 1] = crank + d < csize && k + d < nb;
                    const unsigned target = push_to[d - 1] ? crank + d : crank;
                    push_dst[d - 1] = map_to_rank(
                        smem_addr(xpush + (d - 1) * kB + trow), target);
                    push_bar[d - 1] = map_to_rank(
                        bar_push + 8 * ((d - 1) * kNSB), target);
                }
                // ---- diagonal block (see trsv.cu): y_g = Inv_g rhs_g for all
                //      groups at once, then x_g = y_g - sum_{t<g} M(g,t) x_t
                {
                    const int c_own = mem_sub * kSB + 2 * seg;
                    Ar y;
                    {
                        const Pair<Ar>* Trow = reinterpret_cast<const Pair<Ar>*>(
                            D + trow * kLD + c_own);
                        const Pair<Ar>* v =
                            reinterpret_cast<const Pair<Ar>*>(rhs_cur + c_own);
                        Ar p0 = Ar{}, p1 = Ar{};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const Pair<Ar> t = Trow[4 * e];
                            const Pair<Ar> w = v[4 * e];
                            p0 = fma_ar(t.a, w.a, p0);
                            p1 = fma_ar(t.b, w.b, p1);
                        }
                        y = p0 + p1;
                        y += shfl_xor(y, 1);
                        y += shfl_xor(y, 2);
                    }
                    Ar corr = Ar{};
#pragma unroll
                    for (int step = 0; step < kNSB; ++step) {
                        const int ms = UPPER ? kNSB - 1 - step : step;
                        if (grp == step) {
                            if (TRACE && real && trace != nullptr && seg == 0 &&
                                (trow & 31) == 0) {
                                trace[k * 64 + 16 + 2 * step] = clock64();
                            }
                            // round through storage: later rows see what the
                            // accessor re-reads
                            const St stored = to_st<St, Ar>(y - corr);
                            const Ar back = to_ar<Ar, St>(stored);
                            if (seg == 0) {
                                xsol[trow] = back;
                                if (real) {
                                    // the block rows that come next first:
                                    // they are the ones waiting
#pragma unroll
                                    for (int d = 1; d <= kTail; ++d) {
                                        if (push_to[d - 1]) {
                                            st_async(push_dst[d - 1], back,
                                                     push_bar[d - 1] + 8 * step);
                                        }
                                    }
                                }
                            }
                            if (step + 1 < kNSB) {
                                named_barrier_arrive(5 + step,
                                                     4 * kWarp * (kNSB - step));
                            }
                            if (real && seg == 0) {
                                const std::int64_t gi = r0 + trow;
                                if (gi < n) {
                                    st_volatile(xs + gi,
                                                Sentinel<Ar>::clean(back));
                                    x[gi * incx] = stored;
                                }
                            }
                            if (TRACE && real && trace != nullptr && seg == 0 &&
                                (trow & 31) == 0) {
                                trace[k * 64 + 40 + step] =
                                    static_cast<long long>(globaltimer_ns());
                                trace[k * 64 + 17 + 2 * step] = clock64();
                            }
                        } else if (grp > step) {
                            const int c2 = ms * kSB + 2 * seg;
                            const Pair<Ar>* Mrow =
                                reinterpret_cast<const Pair<Ar>*>(D + trow * kLD +
                                                                  c2);
                            Pair<Ar> mreg[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                mreg[e] = Mrow[4 * e];
                                pin_register(mreg[e].a);
                                pin_register(mreg[e].b);
                            }
                            named_barrier_sync(5 + step,
                                               4 * kWarp * (kNSB - step));
                            const Pair<Ar>* xv =
                                reinterpret_cast<const Pair<Ar>*>(xsol + c2);
                            Ar p0 = Ar{}, p1 = Ar{};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const Pair<Ar> w = xv[4 * e];
                                p0 = fma_ar(mreg[e].a, w.a, p0);
                                p1 = fma_ar(mreg[e].b, w.b, p1);
                            }
                            Ar sum = p0 + p1;
                            sum += shfl_xor(sum, 1);
                            sum += shfl_xor(sum, 2);
                            corr += sum;
                        }
                    }
                }
                // both passes end with all consumer warps in step (the
                // rehearsal must not run into the real pass's barriers)
                __syncthreads();
                if (real) {
                    ACCBLAS_TRACE(10, clock64());
                }
            }
            ACCBLAS_TRACE(12, static_cast<long long>(globaltimer_ns()));

            // ---- the last block row re-arms the workspace for the next call
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
    }
#undef ACCBLAS_TRACE
    // nobody leaves while a neighbour may still read its ticket or write its
    // mailbox
    cluster_sync_all();
}

template <typename St, typename Ar, bool UPPER, bool UNIT, int VW,
          bool TRACE = false>
int launch_cluster(Handle* h, std::int64_t n, const St* A, std::int64_t lda,
                   St* x, std::int64_t incx, Ar* xs, unsigned* ticket,
                   long long* trace, cudaStream_t stream)
{
    auto kernel = trsv_cluster_kernel<St, Ar, UPPER, UNIT, VW, TRACE>;
    const size_t smem = sizeof(Ar) * (kB * kLD + (kRing + kTail + 4) * kB);
    // the opt-in and the cluster occupancy are per device (and instantiation)
    static int clusters_on[64] = {};  // 0 = not asked yet, -1 = none fit
    const int slot = (h->device >= 0 && h->device < 64) ? h->device : 0;
    if (clusters_on[slot] == 0 || slot != h->device) {
        ACCBLAS_CUDA(cudaFuncSetAttribute(
            kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
            static_cast<int>(smem)));
        cudaLaunchConfig_t probe = {};
        probe.gridDim = dim3(kMaxCluster);
        probe.blockDim = dim3(kThreads);
        probe.dynamicSmemBytes = smem;
        cudaLaunchAttribute pa[1];
        pa[0].id = cudaLaunchAttributeClusterDimension;
        pa[0].val.clusterDim.x = kMaxCluster;
        pa[0].val.clusterDim.y = 1;
        pa[0].val.clusterDim.z = 1;
        probe.attrs = pa;
        probe.numAttrs = 1;
        int clusters = 0;
        cudaError_t err = cudaOccupancyMaxActiveClusters(&clusters, kernel, &probe);
        if (err != cudaSuccess) {
            cudaGetLastError();
            clusters = 0;
        }
        if (slot == h->device) {
            clusters_on[slot] = clusters > 0 ? clusters : -1;
        } else if (clusters <= 0) {
            return ACCBLAS_ERR_UNSUPPORTED;
        }
    }
    if (slot == h->device && clusters_on[slot] < 0) {
        return ACCBLAS_ERR_UNSUPPORTED;  // caller falls back to trsv.cu
    }
    const std::int64_t nb = (n + kB - 1) / kB;
    int cs = kMaxCluster;
    while (cs > 1 && cs / 2 >= nb) {
        cs /= 2;  // small systems: no more padding CTAs than necessary
    }
    const std::int64_t grid = (nb + cs - 1) / cs * cs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = static_cast<unsigned>(cs);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int l2_ahead = tuning().trsv_l2_ahead >= 0
                             ? tuning().trsv_l2_ahead
                             : trsv_default_l2_ahead<St, Ar>();
    ACCBLAS_CUDA(cudaLaunchKernelEx(&cfg, kernel, n, A, lda, x, incx, xs, ticket,
                                    trace, l2_ahead));
    return ACCBLAS_OK;
}

template <typename St, typename Ar>
int launch_trsv_cluster(Handle* h, bool upper, bool unit, int vw,
                        std::int64_t n, const St* A, std::int64_t lda, St* x,
                        std::int64_t incx, Ar* xs, unsigned* ticket,
                        long long* trace, cudaStream_t stream)
{
#if defined(ACCBLAS_DEV_HOOKS)
    if (trace != nullptr) {
        // development timeline (tools/trsv_trace.py): one instantiation only
        constexpr int kTraceVw = sizeof(St) == 2 ? 8 : 16;
        if (upper || !unit || vw != kTraceVw) {
            set_error("trsv trace: lower / unit / 16-byte aligned operands only");
            return ACCBLAS_ERR_UNSUPPORTED;
        }
        return launch_cluster<St, Ar, false, true, kTraceVw, true>(
            h, n, A, lda, x, incx, xs, ticket, trace, stream);
    }
#endif
#define ACCBLAS_TRSV_CASE(U, N, V)                                            \
    if (upper == U && unit == N && vw == V) {                                 \
        return launch_cluster<St, Ar, U, N, V>(h, n, A, lda, x, incx, xs,     \
                                               ticket, trace, stream);        \
    }
#define ACCBLAS_TRSV_WIDTHS(U, N)                  \
    ACCBLAS_TRSV_CASE(U, N, 0)                     \
    ACCBLAS_TRSV_CASE(U, N, 8)                     \
    if constexpr (sizeof(St) != 2) {               \
        ACCBLAS_TRSV_CASE(U, N, 16)                \
    }
    ACCBLAS_TRSV_WIDTHS(false, false)
    ACCBLAS_TRSV_WIDTHS(false, true)
    ACCBLAS_TRSV_WIDTHS(true, false)
    ACCBLAS_TRSV_WIDTHS(true, true)
#undef ACCBLAS_TRSV_WIDTHS
#undef ACCBLAS_TRSV_CASE
    return ACCBLAS_ERR_INVALID;
}

// Entry point of one arithmetic type (instantiated by trsv_cluster_f64.cu /
// trsv_cluster_f32.cu: two translation units compile in parallel).
template <typename Ar>
int trsv_cluster_ar(Handle* h, int st, bool upper, bool unit, int vw,
                    std::int64_t n, const void* A, std::int64_t lda, void* x,
                    std::int64_t incx, void* xs, unsigned* ticket,
                    long long* trace, cudaStream_t stream)
{
    auto run = [&](auto st_tag) {
        using St = decltype(st_tag);
        return launch_trsv_cluster<St, Ar>(
            h, upper, unit, vw, n, static_cast<const St*>(A), lda,
            static_cast<St*>(x), incx, static_cast<Ar*>(xs), ticket, trace,
            stream);
    };
    switch (st) {
    case ACCBLAS_F64:
        return run(double{});
    case ACCBLAS_F32:
        return run(float{});
    case ACCBLAS_F16:
        return run(__half{});
    default:
        set_error("invalid storage dtype %d", st);
        return ACCBLAS_ERR_INVALID;
    }
}

}  // namespace
}  // namespace accblas
