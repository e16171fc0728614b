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
//     consumed (raw-word registers nothing touches before the conversion
//     phase), the panel is widened to the arithmetic type BEFORE the matching
//     x block is looked at, and a row sum needs two shuffle levels; tiles of
//     later blocks are requested into L2 ahead of time;
//   * progress communicated through the solution itself: solved entries are
//     published (already rounded through the storage type, as the reference's
//     accessor write/read does) into a workspace vector that starts out as a
//     NaN sentinel -- no flag, no fence, no acquire/release pair.  Consumers
//     fetch the next x block one block ahead with cp.async (L2 -> shared
//     memory); a CTA that is behind the chain never waits for L2, a CTA that
//     has caught up re-polls the missing entries, all of them in flight at
//     once;
//   * the diagonal 128x128 tile lives in shared memory (leading dimension 136
//     so 16-byte loads are conflict-free); each of its four 32x32 diagonal
//     sub-blocks is inverted by one warp, lane j = column j by substitution in
//     registers; then M(g,t) = Inv_g D(g,t) is formed in place, so that the
//     solve x_g = Inv_g rhs_g - sum_{t<g} M(g,t) x_t has ONE 32x32 product
//     between consecutive sub-block solutions; warp groups hand the sub-block
//     solutions to each other through named barriers;
//   * the reduction + diagonal-solve code is rehearsed once on scratch data
//     while the CTA waits, so it is warm in the instruction cache when it
//     runs for real on the critical path.
// The last CTA re-arms the workspace (sentinels, ticket) for the next call.
#include "common.cuh"
#include "trsv_common.cuh"
#include "tuning.h"

namespace accblas {

// trsv_cluster_f64.cu / trsv_cluster_f32.cu (kernel in trsv_cluster.cuh);
// ACCBLAS_ERR_UNSUPPORTED = no cluster of that kernel fits this device
int trsv_cluster_f64(Handle* h, int st, bool upper, bool unit, int vw,
                     std::int64_t n, const void* A, std::int64_t lda, void* x,
                     std::int64_t incx, void* xs, unsigned* ticket,
                     long long* trace, cudaStream_t stream);
int trsv_cluster_f32(Handle* h, int st, bool upper, bool unit, int vw,
                     std::int64_t n, const void* A, std::int64_t lda, void* x,
                     std::int64_t incx, void* xs, unsigned* ticket,
                     long long* trace, cudaStream_t stream);

namespace {

using namespace trsv_detail;

// Thread layout for 128 x 128 tiles: 4 consecutive lanes share a row
// (row = tid / 4, seg = tid % 4) and a lane owns the columns
// 16*i + 4*seg + e (i < 8, e < 4).  Each load instruction therefore reads 16
// consecutive elements per row for 8 rows (full 32-byte sectors), and a row
// sum needs only TWO shuffle levels -- the shuffle unit handles one warp per
// cycle per SM, so the previous "lane = column" layout (five levels for each
// of 8 rows per warp) spent ~1300 cycles of every block step just shuffling.
// TRACE = false (production): the timeline probes below compile away entirely
// -- even predicated off they cost registers (this kernel sits at the 128
// register limit of a 512-thread CTA) and scoreboard waits.
template <typename St, typename Ar, bool UPPER, bool UNIT, int VW,
          bool TRACE>
__global__ __launch_bounds__(kThreads, 1) void trsv_kernel(
    std::int64_t n, const St* __restrict__ A, std::int64_t lda,
    St* __restrict__ x, std::int64_t incx, Ar* xs,
    unsigned* __restrict__ ticket, long long* __restrict__ trace_arg,
    int l2_ahead, int whole_block_spin)
{
    long long* const trace = TRACE ? trace_arg : nullptr;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ar* D = reinterpret_cast<Ar*>(smem_raw);  // kB x kLD
    Ar* xcol = D + kB * kLD;                  // kXRing x kB, staged x blocks
    Ar* rhs = xcol + kXRing * kB;             // kB
    Ar* xsol = rhs + kB;                      // kB
    Ar* inv_diag = xsol + kB;                 // kB
    Ar* scratch = inv_diag + kB;              // kB, rehearsal right-hand side
    __shared__ unsigned k_shared;
    __shared__ int mode_s[kXRing];

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
    if (TRACE && trace != nullptr && tid == 0) {    \
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
        rhs[tid] = (tid < bs) ? to_ar<Ar, St>(x[(r0 + tid) * incx]) : Ar{0};
        scratch[tid] = rhs[tid];
        xsol[tid] = Ar{0};
    }
    __syncthreads();
    ACCBLAS_TRACE(1, clock64());
    if (warp < kNSB && warp * kSB < bs) {  // padding-only sub-blocks are identity
        invert_subblock<Ar, UPPER, UNIT>(D + (warp * kSB) * kLD + warp * kSB,
                                         inv_diag + warp * kSB, lane);
    }
    __syncthreads();
    // ---- M(g,t) = Inv_g * D(g,t) for every sub-block pair with t solved
    //      before g, in place.  With these the solve of the diagonal tile is
    //          x_g = Inv_g rhs_g - sum_{t<g} M(g,t) x_t
    //      i.e. ONE 32x32 matrix-vector product between consecutive
    //      sub-block solutions instead of two (update, then multiply by the
    //      inverse).  192 k FMAs per CTA, done once, while the CTA would be
    //      waiting for its predecessors anyway.
    {
        const int i = tid >> 4;         // row inside the 32 x 32 block
        const int jp = (tid & 15) * 2;  // column pair
        Ar out[6][2];
        int slot = 0;
#pragma unroll
        for (int g = 1; g < kNSB; ++g) {
#pragma unroll
            for (int t = 0; t < g; ++t) {
                const int mg = UPPER ? kNSB - 1 - g : g;  // memory sub-block
                const int mt = UPPER ? kNSB - 1 - t : t;
                const Ar* inv = D + (mg * kSB + i) * kLD + mg * kSB;
                const Ar* blk = D + (mg * kSB) * kLD + mt * kSB + jp;
                Ar o0 = Ar{}, o1 = Ar{};
                // sub-blocks made of padding only (short last block, small
                // systems) hold zeros already: nothing to multiply
                if (mg * kSB < bs && mt * kSB < bs) {
#pragma unroll 8
                    for (int kk = 0; kk < kSB; ++kk) {
                        const Ar a = inv[kk];
                        const Pair<Ar> d =
                            *reinterpret_cast<const Pair<Ar>*>(blk + kk * kLD);
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
                *reinterpret_cast<Pair<Ar>*>(D + (mg * kSB + i) * kLD +
                                             mt * kSB + jp) = o;
                ++slot;
            }
        }
    }
    __syncthreads();
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
    // panels of a 128-column block are walked in the order their x entries
    // are produced (ascending columns for lower, descending for upper)
    auto load_panel = [&](std::int64_t jj, int p, Quad<St> (&dst)[Q]) {
        const std::int64_t pbj = UPPER ? nb - 1 - jj : jj;
        const int pm = UPPER ? NP - 1 - p : p;  // memory panel
        const std::int64_t c0 = pbj * kB + pm * (16 * Q);
        // Only the LAST physical block can be cut off by the edge of the
        // matrix, and only `upper` ever streams it (lower: pbj < pb).  Every
        // other tile takes the path without per-quad edge arithmetic (the
        // 64-bit compares, selects and reconvergence points of the general
        // path were ~100 of the ~345 instructions of a warp's block
        // iteration; ncu source view, session r02z).
        if (!UPPER || (pbj + 1) * kB <= n) {
#pragma unroll
            for (int i = 0; i < Q; ++i) {
                dst[i] = load_quad<St, VW>(row_ptr + c0 + 16 * i, kEPL);
            }
        } else {
#pragma unroll
            for (int i = 0; i < Q; ++i) {
                const std::int64_t col = c0 + 16 * i + kEPL * seg;
                const std::int64_t left = n - col;
                const int valid =
                    left >= kEPL ? kEPL
                                 : (left > 0 ? static_cast<int>(left) : 0);
                dst[i] = load_quad<St, VW>(row_ptr + c0 + 16 * i, valid);
            }
        }
    };
    // warp 0: copy the 128 progress-vector entries of physical block `pblock`
    // into x buffer `b` (L2 -> shared memory, 16 bytes per copy)
    auto prefetch_x = [&](std::int64_t pblock, int b) {
        constexpr int kChunks = kB * static_cast<int>(sizeof(Ar)) / 16;
        const char* src = reinterpret_cast<const char*>(xs + pblock * kB);
        const unsigned dst = static_cast<unsigned>(
            __cvta_generic_to_shared(xcol + b * kB));
#pragma unroll
        for (int c = lane; c < kChunks; c += kWarp) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                             dst + c * 16),
                         "l"(src + c * 16)
                         : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // one commit group per block iteration, empty past the last dependency, so
    // that "all but the kXAhead - 1 youngest groups" is always "block jj"
    auto prefetch_x_slot = [&](std::int64_t j, std::int64_t deps_, int b) {
        if (j < deps_) {
            prefetch_x(UPPER ? nb - 1 - j : j, b);
        } else {
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };
    // Ask L2 for the off-diagonal tiles of a GROUP of consecutive blocks of
    // the solve order, one group ahead.  Two reasons:
    //  * the register pipeline above keeps only ONE panel in flight per CTA,
    //    so a block iteration would be bounded by the DRAM latency; with the
    //    lines already in L2 the same loads return in a few hundred cycles;
    //  * a 128-column tile is only 256 / 512 / 1024 bytes per row (fp16 / fp32
    //    / fp64) at a row stride of tens of KB -- poor DRAM page locality.  A
    //    group is l2_group_bytes (>= 1 KB) of CONSECUTIVE bytes per row.
    // Thread = lines seg, seg + 4, ... of its own row.
    const int group_blocks =
        l2_ahead > 0 ? max(1, l2_ahead / (kB * static_cast<int>(sizeof(St))))
                     : 0;
    auto l2_prefetch_group = [&](std::int64_t j0) {
        // blocks j0 .. j0 + group_blocks - 1 of the solve order (dependencies
        // only: jj < k), as one contiguous column range
        std::int64_t j1 = j0 + group_blocks;
        j1 = j1 < k ? j1 : k;
        if (j0 >= j1) {
            return;
        }
        const std::int64_t c_lo = (UPPER ? nb - j1 : j0) * kB;
        std::int64_t c_hi = (UPPER ? nb - j0 : j1) * kB;
        c_hi = c_hi < n ? c_hi : n;
        constexpr int kLineElems = 128 / static_cast<int>(sizeof(St));
        const St* row_base = row_ptr - kEPL * seg;
        // (one cp.async.bulk.prefetch.L2 per row and group instead of one
        // prefetch per line measured the same: 318 vs 319.5 us, session r02y --
        // UBLKPF takes uniform registers, so the compiler serialises the lanes)
        for (std::int64_t c = c_lo + seg * kLineElems; c < c_hi;
             c += 4 * kLineElems) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(row_base + c));
        }
    };
    const int mem_sub = trow >> 5;
    const int grp = UPPER ? kNSB - 1 - mem_sub : mem_sub;  // solve index

#pragma unroll 1
    // (the first CTA of the solve order has nobody to wait for: no rehearsal)
    for (int pass = (k == 0 ? 1 : 0); pass < 2; ++pass) {
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
        //      panel are in flight while the current one is consumed.
        //      The x block a panel needs is fetched one block AHEAD with
        //      cp.async (L2 -> shared memory, no registers): a CTA that is
        //      behind the chain finds its x already in shared memory and pays
        //      no L2 round trip per block ("fast": one barrier, 128 columns of
        //      FMAs).  A CTA that has caught up finds sentinels; it then polls
        //      the block 32 entries at a time, in the order the producing CTA
        //      publishes its sub-block solutions, so that only the last 32
        //      columns' FMAs follow the arrival of the block's last sub-block.
        Quad<St> cur[Q];
        if (deps > 0) {
            load_panel(0, 0, cur);
            if (warp == 0) {
#pragma unroll
                for (int a = 0; a < kXAhead; ++a) {
                    prefetch_x_slot(a, deps, a);
                }
            }
            if (group_blocks > 0) {
                // the rest of group 0 (block 0 itself is being loaded)
                l2_prefetch_group(0);
            }
        }
        int buf = 0;
        const bool phase_log =
            TRACE && trace != nullptr && tid == 0 && k == nb - 1;
        long long ph[5] = {0, 0, 0, 0, 0};
        long long ph_t = 0;
        for (std::int64_t jj = 0; jj < deps; ++jj) {
            const std::int64_t pbj = UPPER ? nb - 1 - jj : jj;
            if (phase_log) {
                ph_t = clock64();
            }
            Ar* xblk = xcol + buf * kB;
            bool fast = false;
            if (group_blocks > 0 && jj % group_blocks == 0) {
                l2_prefetch_group(jj + group_blocks);
            }
            // warp 0: is the prefetched copy of this x block complete?  If not,
            // ask L2 once more for the missing entries -- all of them in one
            // go, and the answers are looked at only after the panel has been
            // widened, so the round trip is hidden
            Ar repoll[kNSB];
            bool ok = true;
            if (warp == 0) {
                if (jj == deps - 1) {
                    ACCBLAS_TRACE(3, clock64());
                }
                asm volatile("cp.async.wait_group %0;" ::"n"(kXAhead - 1)
                             : "memory");
                __syncwarp();  // lanes read entries other lanes copied
#pragma unroll
                for (int sb = 0; sb < kNSB; ++sb) {
                    const int idx = sb * kSB + lane;
                    repoll[sb] = Ar{0};
                    if (pbj * kB + idx < n) {
                        repoll[sb] = xblk[idx];
                        ok = ok && !Sentinel<Ar>::is(repoll[sb]);
                    } else {
                        xblk[idx] = Ar{0};
                    }
                }
                ok = __all_sync(0xffffffffu, ok);
                if (!ok) {
#pragma unroll
                    for (int sb = 0; sb < kNSB; ++sb) {
                        const std::int64_t pc = pbj * kB + sb * kSB + lane;
                        if (pc < n && Sentinel<Ar>::is(repoll[sb])) {
                            repoll[sb] = ld_volatile(xs + pc);
                        }
                    }
                }
            }
            if (phase_log) {
                const long long now = clock64();
                ph[0] += now - ph_t;  // x prefetch wait + check (+ re-poll issue)
                ph_t = now;
            }
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const int pm = UPPER ? NP - 1 - p : p;
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
                        cv[i][e] = cur[i].template get<Ar>(e);
                        if (pre_convert<St, Ar>::value) {
                            pin_register(cv[i][e]);
                        }
                    }
                }
                if (p == 0 && phase_log) {
                    const long long now = clock64();
                    ph[1] += now - ph_t;  // next-panel loads issued, panel widened
                    ph_t = now;
                }
                if (p == 0 && warp == 0) {
                    if (!ok) {
                        ok = true;
#pragma unroll
                        for (int sb = 0; sb < kNSB; ++sb) {
                            const int idx = sb * kSB + lane;
                            if (pbj * kB + idx < n) {
                                xblk[idx] = repoll[sb];
                                ok = ok && !Sentinel<Ar>::is(repoll[sb]);
                            }
                        }
                        ok = __all_sync(0xffffffffu, ok);
                    }
                    if (!ok && whole_block_spin) {
                        // caught up with the chain: spin on whatever is still
                        // missing, all of it in flight at once
                        do {
                            ok = true;
#pragma unroll
                            for (int sb = 0; sb < kNSB; ++sb) {
                                const std::int64_t pc =
                                    pbj * kB + sb * kSB + lane;
                                if (pc < n && Sentinel<Ar>::is(repoll[sb])) {
                                    repoll[sb] = ld_volatile(xs + pc);
                                }
                            }
#pragma unroll
                            for (int sb = 0; sb < kNSB; ++sb) {
                                const int idx = sb * kSB + lane;
                                if (pbj * kB + idx < n) {
                                    xblk[idx] = repoll[sb];
                                    ok = ok && !Sentinel<Ar>::is(repoll[sb]);
                                }
                            }
                            ok = __all_sync(0xffffffffu, ok);
                        } while (!ok);
                    }
                    fast = ok;
                    if (lane == 0) {
                        mode_s[buf] = fast ? 1 : 0;
                    }
                }
                constexpr int SUBS = kNSB / NP;  // sub-blocks per panel
#pragma unroll
                for (int t = 0; t < SUBS; ++t) {
                    // t-th sub-block of this panel to arrive; ts = its
                    // position inside the panel in memory order
                    const int ts = UPPER ? SUBS - 1 - t : t;
                    const int sbm = pm * SUBS + ts;  // memory sub-block
                    const bool first = p == 0 && t == 0;
                    if (first || !fast) {
                        if (warp == 0 && !fast) {
                            const int idx = sbm * kSB + lane;
                            const std::int64_t pc = pbj * kB + idx;
                            if (pc < n) {
                                Ar v = xblk[idx];
                                while (Sentinel<Ar>::is(v)) {
                                    v = ld_volatile(xs + pc);
                                }
                                xblk[idx] = v;
                            }
                        }
                        if (warp == 0 && jj == deps - 1 &&
                            (fast || (p == NP - 1 && t == SUBS - 1))) {
                            ACCBLAS_TRACE(4, clock64());
                            ACCBLAS_TRACE(13, static_cast<long long>(
                                                  globaltimer_ns()));
                        }
                        __syncthreads();
                        if (first && phase_log) {
                            const long long now = clock64();
                            ph[2] += now - ph_t;  // re-poll evaluation + first barrier
                            ph_t = now;
                        }
                        if (first) {
                            fast = mode_s[buf] != 0;
                            // every warp is past block jj - 1: its x buffer
                            // can take block jj + kXAhead
                            if (warp == 0) {
                                prefetch_x_slot(jj + kXAhead, deps,
                                                (buf + kXAhead) % kXRing);
                            }
                        }
                    }
                    const Ar* xb = xblk + pm * (16 * Q) + kEPL * seg;
#pragma unroll
                    for (int ii = 0; ii < 2; ++ii) {
                        const int i = 2 * ts + ii;
#pragma unroll
                        for (int e = 0; e < kEPL; ++e) {
                            const int slot = (p * Q + i) % NACC;
                            acc[slot] =
                                fma_ar(cv[i][e], xb[16 * i + e], acc[slot]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < Q; ++i) {
                    // nothing touches the words of the next panel before this
                    // point: its loads stay in flight across the iteration
                    nxt[i].pin();
                    cur[i] = nxt[i];
                }
            }
            if (phase_log) {
#pragma unroll
                for (int i = 0; i < NACC; ++i) {
                    pin_register(acc[i]);
                }
                const long long now = clock64();
                ph[3] += now - ph_t;  // FMAs (and later barriers on the slow path)
                ph_t = now;
            }
            if (phase_log && fast) {
                ++ph[4];
            }
            buf = (buf + 1) % kXRing;
        }
        if (real) {
            ACCBLAS_TRACE(5, clock64());
            if (phase_log) {
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    trace[k * 64 + 48 + i] = ph[i];
                }
            }
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
        }
        // "rhs of my sub-block is complete": its 32 rows belong to the four
        // warps of this group only
        named_barrier_sync(1 + grp, 4 * kWarp);
        if (real) {
            ACCBLAS_TRACE(6, clock64());
        }

        // ---- diagonal block.  Warp group g = the 128 threads whose rows lie
        //      in the g-th sub-block to be solved.
        //        y_g = Inv_g rhs_g                  all groups at once
        //        x_g = y_g - sum_{t<g} M(g,t) x_t   chain, one product per link
        //      Synchronisation by NAMED barriers between exactly the warps
        //      involved; the solving group only arrives and moves on:
        //        id 1+g : "rhs of group g is complete"      128 threads
        //        id 5+s : "x_s is in shared memory"         128 * (4 - s)
        {
            const bool probe = TRACE && real && trace != nullptr && seg == 0 &&
                               (trow & 31) == 0;
            const int c_own = mem_sub * kSB + 2 * seg;
            Ar y;
            {
                const Pair<Ar>* Trow =
                    reinterpret_cast<const Pair<Ar>*>(D + trow * kLD + c_own);
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
                if (probe && grp == 0) {
                    pin_register(y);
                    trace[k * 64 + 35] = clock64();
                }
                y += __shfl_xor_sync(0xffffffffu, y, 1);
                y += __shfl_xor_sync(0xffffffffu, y, 2);
                if (probe && grp == 0) {
                    pin_register(y);
                    trace[k * 64 + 36] = clock64();
                }
            }
            Ar corr = Ar{};
#pragma unroll
            for (int step = 0; step < kNSB; ++step) {
                const int ms = UPPER ? kNSB - 1 - step : step;  // memory order
                if (grp == step) {
                    if (probe) {
                        trace[k * 64 + 16 + 4 * step] = clock64();
                    }
                    // round through storage: later rows see what the accessor
                    // re-reads
                    const St stored = to_st<St, Ar>(y - corr);
                    const Ar back = to_ar<Ar, St>(stored);
                    if (seg == 0) {
                        xsol[trow] = back;
                    }
                    if (step + 1 < kNSB) {
                        // (PTX ISA, bar.arrive example: a store to shared
                        // memory followed by bar.arrive is visible to the
                        // threads that bar.sync on the same barrier)
                        named_barrier_arrive(5 + step,
                                             4 * kWarp * (kNSB - step));
                    }
                    if (probe) {
                        trace[k * 64 + 17 + 4 * step] = clock64();
                        trace[k * 64 + 40 + step] =
                            static_cast<long long>(globaltimer_ns());
                    }
                    if (real && seg == 0) {
                        // publish: progress vector first (the next block row
                        // spins on it), then the caller's x
                        const std::int64_t gi = r0 + trow;
                        if (gi < n) {
                            st_volatile(xs + gi, Sentinel<Ar>::clean(back));
                            x[gi * incx] = stored;
                        }
                    }
                } else if (grp > step) {
                    // M(g, step) entries of my row: constants, fetched before
                    // the wait
                    const int c2 = ms * kSB + 2 * seg;
                    const Pair<Ar>* Mrow =
                        reinterpret_cast<const Pair<Ar>*>(D + trow * kLD + c2);
                    Pair<Ar> mreg[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        mreg[e] = Mrow[4 * e];
                        pin_register(mreg[e].a);
                        pin_register(mreg[e].b);
                    }
                    named_barrier_sync(5 + step, 4 * kWarp * (kNSB - step));
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
                    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                    corr += sum;
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

template <typename St, typename Ar, bool UPPER, bool UNIT, int VW,
          bool TRACE = false>
int launch_one(std::int64_t n, const St* A, std::int64_t lda, St* x,
               std::int64_t incx, Ar* xs, unsigned* ticket, long long* trace,
               cudaStream_t stream)
{
    auto kernel = trsv_kernel<St, Ar, UPPER, UNIT, VW, TRACE>;
    const size_t smem = sizeof(Ar) * (kB * kLD + (4 + kXRing) * kB);
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
        n, A, lda, x, incx, xs, ticket, trace,
        tuning().trsv_l2_ahead >= 0 ? tuning().trsv_l2_ahead
                                    : trsv_default_l2_ahead<St, Ar>(),
        tuning().trsv_whole_block_spin);
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
    // whole 128-entry blocks: the kernel prefetches x a block at a time
    const size_t need =
        static_cast<size_t>((n + kB - 1) / kB) * kB * sizeof(Ar);
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

    // vector loads need 8-byte aligned rows; 16-byte aligned rows additionally
    // let fp32 / fp64 quads use 128-bit loads (an fp16 quad is 8 bytes)
    const std::uintptr_t align_bits =
        reinterpret_cast<std::uintptr_t>(A) |
        static_cast<std::uintptr_t>(static_cast<std::uint64_t>(lda) * sizeof(St));
    // 16 -> 128-bit loads, 8 -> 64-bit loads, 0 -> scalar; an fp16 quad is 8
    // bytes, so its 16-byte case is the 8-byte one
    int vw = (align_bits & 15u) == 0 ? 16 : ((align_bits & 7u) == 0 ? 8 : 0);
    if (sizeof(St) == 2 && vw == 16) {
        vw = 8;
    }
    const bool upper = uplo == ACCBLAS_UPPER;
    const bool unit = diag == ACCBLAS_UNIT;
    // which kernel: same-box A/B at n = 16384 (tools/trsv_check.py,
    // profiles/r02_trsv_ab.txt) -- the cluster kernel is 12-24 % faster for fp16
    // storage and 12-30 % slower for fp32 / fp64 storage
    const int variant = tuning().trsv_variant;
    if (variant == 0 || (variant < 0 && sizeof(St) == 2)) {
        constexpr int st_code =
            std::is_same<St, double>::value
                ? ACCBLAS_F64
                : (std::is_same<St, float>::value ? ACCBLAS_F32 : ACCBLAS_F16);
        if constexpr (std::is_same<Ar, double>::value) {
            rc = trsv_cluster_f64(h, st_code, upper, unit, vw, n, A, lda, x,
                                  incx, xs, ticket, trace, stream);
        } else {
            rc = trsv_cluster_f32(h, st_code, upper, unit, vw, n, A, lda, x,
                                  incx, xs, ticket, trace, stream);
        }
        if (rc != ACCBLAS_ERR_UNSUPPORTED) {
            return rc;
        }
        // no cluster of that kernel fits this device: one CTA per block row
    }
#if defined(ACCBLAS_DEV_HOOKS)
    if (trace != nullptr) {
        // development timeline (tools/trsv_trace.py): one instantiation only
        constexpr int kTraceVw = sizeof(St) == 2 ? 8 : 16;
        if (upper || !unit || vw != kTraceVw) {
            set_error("trsv trace: lower / unit / 16-byte aligned operands only");
            return ACCBLAS_ERR_UNSUPPORTED;
        }
        return launch_one<St, Ar, false, true, kTraceVw, true>(
            n, A, lda, x, incx, xs, ticket, trace, stream);
    }
#endif
#define ACCBLAS_TRSV_CASE(U, N, V)                                          \
    if (upper == U && unit == N && vw == V) {                               \
        return launch_one<St, Ar, U, N, V>(n, A, lda, x, incx, xs, ticket,  \
                                           trace, stream);                  \
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
