// GEMV  y = alpha * A * x + beta * y  over reduced-precision row-major storage.
//
// Replaces kernel::acc_gemv / kernel::gemv (one 512-thread CTA per row, scalar
// 4/8-byte loads, x re-read and re-converted per element, a shared-memory tree
// per row; /root/reference/cuda/gemv_kernels.cuh:30-113) with a streaming
// kernel shaped for B200:
//   * a CTA owns RG row groups of ROWS consecutive rows; COLW warps share a
//     row group and walk interleaved column chunks of 32 lanes x UNROLL
//     128-bit vectors, so ROWS*UNROLL independent 16-byte L1-bypassing loads
//     are in flight per lane and every request is a fully coalesced 512-byte
//     line group.  Many small CTAs (thousands) keep the SMs evenly loaded to
//     the very end of the launch -- measured on B200 the "one warp owns 4
//     whole rows" shape loses 20 % to the tail, the column-split shape does
//     not (profiles/r01_summary.md);
//   * the matching x vectors are loaded once per chunk through L1 (shared by
//     all CTAs of the SM), converted to the arithmetic type once and reused
//     for the ROWS rows;
//   * storage -> arithmetic conversion happens in registers, accumulation is
//     FMA in the arithmetic type, the row reduction is a warp-shuffle butterfly
//     plus one shared-memory hop between the COLW warps (fixed order);
//   * fp16 storage with fp64 arithmetic would be bound by the 64-bit
//     conversion pipe (F2F issues at a quarter of the rate the HBM stream
//     needs), so that pair widens half of every vector with integer
//     instructions instead (see HalfToDoubleScaled below) and runs a register
//     software pipeline fed by 64-bit loads; results are bit-identical;
//   * operands that are only 8- or 4-byte aligned (odd row strides, sliced
//     bases) take the same pipeline with 64- / 32-bit loads, a strided x is
//     packed once: the scalar kernel is left with fp16 rows of odd stride;
//   * back-to-back calls overlap: programmatic dependent launch lets the next
//     GEMV's CTAs occupy the SM slots the last wave leaves empty and prefetch
//     into L2 while they wait for the predecessor to finish.
// Selectable alternatives that were measured and did not win (tools/tune.py):
// a TMA bulk-copy ring (gemv_bulk_kernel), a cp.async ring (AsyncOps).
// The result is rounded once to the storage type on the way out, exactly like
// the accessor's proxy assignment (cuda/gemv_kernels.cuh:106-111).
#include "common.cuh"
#include "tuning.h"

namespace accblas {

#if defined(ACCBLAS_DEV_HOOKS)
// development aid (libaccblas_b200_dev.so only): when non-null, every CTA of
// gemv_stream_kernel records its start / end time (globaltimer ns) and SM id
// at [3*blockIdx.x ...]
__device__ unsigned long long* g_gemv_trace = nullptr;

int set_gemv_trace(unsigned long long* ptr)
{
    ACCBLAS_CUDA(cudaMemcpyToSymbol(g_gemv_trace, &ptr, sizeof(ptr)));
    return ACCBLAS_OK;
}
#endif

namespace {

__device__ __forceinline__ unsigned long long gemv_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ---------------------------------------------------------------------------
// fp16 -> fp64 without the conversion pipe.
//
// Put the 15 magnitude bits of a half h at bits [24:10] of the HIGH word of a
// double (sign at bit 31, low word 0).  Read as a double this is
//     d = value(h) * 2^-1008
// for zero, subnormal and normal h alike: a normal h = 1.m * 2^(e-15) becomes
// 1.m * 2^(e-1023); a subnormal h = 0.m * 2^-14 becomes the fp64 subnormal
// 0.m * 2^-1022; fp64 hardware handles subnormal operands at full speed.
// Multiplying by the pre-scaled x' = x * 2^1008 (exact: |x| <= 65504 keeps x'
// below DBL_MAX) gives d * x' == value(h) * x as real numbers, and FMA rounds
// only once, so fma(d, x', acc) == fma(double(h), double(x), acc) bit for bit.
// Inf/NaN halves (exponent 31) do NOT map correctly; they are detected on the
// otherwise idle fp16 pipe (0 * h is NaN exactly for those) and the warp
// recomputes its part with ordinary conversions.
// Three integer instructions per element instead of one F2F on the
// quarter-rate pipe: isolate the half in the top 16 bits of a word (LOP3 or a
// shift), one signed 32x32->64 multiply by 2^26 (IMAD.WIDE: high word = the
// word shifted right by 6 with the sign replicated, low word = 0 -- the
// register PAIR a double needs comes out of one instruction; building it from
// a 32-bit shift costs an extra "move zero into the low register" per
// element), and one mask on the high word.
// ---------------------------------------------------------------------------
struct HalfToDoubleScaled {
    static constexpr long long kMask =
        static_cast<long long>(0x81FFFC00FFFFFFFFull);
    // 2^1008 as a double: exponent field 1008 + 1023 = 2031
    static __device__ __forceinline__ double scale()
    {
        return __hiloint2double(2031 << 20, 0);
    }
    // t holds the half in bits [31:16] and zeros below
    static __device__ __forceinline__ double widen_top(unsigned t)
    {
        long long p;
        asm("mul.wide.s32 %0, %1, 0x4000000;" : "=l"(p) : "r"(t));
        return __longlong_as_double(p & kMask);
    }
    static __device__ __forceinline__ double low(unsigned w)
    {
        return widen_top(w << 16);
    }
    static __device__ __forceinline__ double high(unsigned w)
    {
        return widen_top(w & 0xFFFF0000u);
    }
};

template <typename Ar, typename St>
struct use_scaled_half : std::false_type {};
template <>
struct use_scaled_half<double, __half> : std::true_type {};

template <typename Ar, typename St, int ROWS, int UNROLL, bool FAST>
struct ChunkOps {
    static constexpr int VEC = vec_traits<St>::elems;
    static constexpr int CHUNK = kWarp * VEC * UNROLL;
    // independent accumulators per (row, vector): fp32 arithmetic gets two so
    // its per-lane chains are as short as the reference's (n / 512 terms)
    static constexpr int SPLIT = std::is_same<Ar, float>::value ? 2 : 1;
    static constexpr int SLOTS = UNROLL * SPLIT;

    // full chunk: no predicates.  `chk` collects NaN iff a non-finite half was
    // consumed by the FAST path.
    static __device__ __forceinline__ void full(const St* const (&row)[ROWS],
                                                const St* __restrict__ x,
                                                std::int64_t c0, int lane,
                                                Ar (&acc)[ROWS][SLOTS],
                                                __half2& chk)
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
        compute(ar, xr, acc, chk);
    }

    // raw 128-bit vectors (from registers or shared memory) -> accumulators
    static __device__ __forceinline__ void compute(
        const uint4 (&ar)[ROWS][UNROLL], const uint4 (&xr)[UNROLL],
        Ar (&acc)[ROWS][SLOTS], __half2& chk)
    {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            Ar xv[VEC];
            unpack_all<Ar, St>(xr[u], xv);
            if constexpr (FAST) {
                const double s = HalfToDoubleScaled::scale();
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    xv[i] = xv[i] * s;
                }
                const __half2 zero = __float2half2_rn(0.0f);
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const unsigned w[4] = {ar[r][u].x, ar[r][u].y, ar[r][u].z,
                                           ar[r][u].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[r][u] = fma(HalfToDoubleScaled::low(w[j]),
                                        xv[2 * j], acc[r][u]);
                        acc[r][u] = fma(HalfToDoubleScaled::high(w[j]),
                                        xv[2 * j + 1], acc[r][u]);
                        chk = __hfma2(*reinterpret_cast<const __half2*>(&w[j]),
                                      zero, chk);
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        Ar& a = acc[r][u * SPLIT + (i % SPLIT)];
                        a = fma_ar(unpack<Ar>(ar[r][u], i, St{}), xv[i], a);
                    }
                }
            }
        }
    }

    // last, partial chunk: whole vectors while they fit, then scalars
    static __device__ __forceinline__ void partial(
        const St* const (&row)[ROWS], const St* __restrict__ x,
        std::int64_t c0, std::int64_t n, int lane, Ar (&acc)[ROWS][SLOTS],
        bool aligned16 = true)
    {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const std::int64_t col = c0 + (u * kWarp + lane) * VEC;
            if (aligned16 && col + VEC <= n) {
                Ar xv[VEC];
                unpack_all<Ar, St>(ldg_cached_128(x + col), xv);
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const uint4 a = ldg_stream_128(row[r] + col);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        Ar& t = acc[r][u * SPLIT + (i % SPLIT)];
                        t = fma_ar(unpack<Ar>(a, i, St{}), xv[i], t);
                    }
                }
            } else {
                const std::int64_t c_end = (col + VEC < n) ? col + VEC : n;
                for (std::int64_t c = col; c < c_end; ++c) {
                    const Ar xv = to_ar<Ar, St>(x[c]);
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) {
                        Ar& t = acc[r][u * SPLIT];
                        t = fma_ar(to_ar<Ar, St>(row[r][c]), xv, t);
                    }
                }
            }
        }
    }
};

// fp16 -> fp64 on the conversion pipe in ONE instruction (F2F.F64.F16 with a
// half selector; `static_cast<double>(__half2float(h))` costs two)
__device__ __forceinline__ double half_bits_to_double(unsigned short bits)
{
    double d;
    asm("cvt.f64.f16 %0, %1;" : "=d"(d) : "h"(bits));
    return d;
}

// One "batch" = one 128-bit vector per lane from each of the ROWS rows plus the
// matching vector of x: 32 * VEC consecutive columns.  The streaming loop keeps
// one batch in flight while the previous one is being consumed (software
// pipeline of depth 2 in registers), so a warp always has ROWS * 512 bytes of
// the matrix stream outstanding -- also while it converts and multiplies.
// ncu on the unpipelined loop: "long scoreboard" was 4.8 of every 7.9 stall
// cycles of Acc<fp64,fp16>, issue slots 56 % busy, DRAM 66 %.
template <typename Ar, typename St, int ROWS, bool FAST, int IW, int CB = 16>
struct BatchOps {
    static constexpr int VEC = vec_traits<St>::elems;
    static constexpr int COLS = kWarp * VEC;
    static constexpr int SPLIT = std::is_same<Ar, float>::value ? 2 : 1;
    static constexpr int SLOTS = 2 * SPLIT;
    // FAST (fp16 storage, fp64 arithmetic): words [0, kIntWords) of a vector
    // are widened on the integer pipes (HalfToDoubleScaled), the remaining
    // words on the conversion pipe, so that neither path has to carry the
    // whole stream: the conversion pipe handles 16 elements/clk/SM, the HBM
    // stream delivers ~14.
    static constexpr int kIntWords = IW;

    struct Batch {
        uint4 x;
        uint4 a[ROWS];
    };

    // Orders the consumption of `b` after every volatile load issued so far
    // (volatile asm statements keep their program order); without it the
    // compiler sinks the next batch's loads below the current batch's
    // arithmetic to save registers, which undoes the pipeline.
    static __device__ __forceinline__ void pin(Batch& b)
    {
        asm volatile("" : "+r"(b.x.x), "+r"(b.x.y), "+r"(b.x.z), "+r"(b.x.w));
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            asm volatile(""
                         : "+r"(b.a[r].x), "+r"(b.a[r].y), "+r"(b.a[r].z),
                           "+r"(b.a[r].w));
        }
    }

    // `col` = first column of this lane's vector
    static __device__ __forceinline__ void load(const St* const (&row)[ROWS],
                                                const St* __restrict__ x,
                                                std::int64_t col, Batch& b)
    {
        if constexpr (CB == 16) {
            b.x = ldg_cached_128_ordered(x + col);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                b.a[r] = ldg_stream_128(row[r] + col);
            }
        } else {
            // operands that are only 8- / 4-byte aligned: the same 16 bytes
            // in 64- / 32-bit pieces
            b.x = ldg_pieces<CB, false>(x + col);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                b.a[r] = ldg_pieces<CB, true>(row[r] + col);
            }
        }
    }

    template <int PARITY>
    static __device__ __forceinline__ void compute(const Batch& b,
                                                   Ar (&acc)[ROWS][SLOTS],
                                                   __half2& chk)
    {
        if constexpr (FAST) {
            const unsigned xw[4] = {b.x.x, b.x.y, b.x.z, b.x.w};
            double xv[VEC];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                xv[2 * j] = half_bits_to_double(
                    static_cast<unsigned short>(xw[j] & 0xffffu));
                xv[2 * j + 1] =
                    half_bits_to_double(static_cast<unsigned short>(xw[j] >> 16));
            }
            const double s = HalfToDoubleScaled::scale();
#pragma unroll
            for (int i = 0; i < 2 * kIntWords; ++i) {
                xv[i] = xv[i] * s;
            }
            const __half2 zero = __float2half2_rn(0.0f);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const unsigned w[4] = {b.a[r].x, b.a[r].y, b.a[r].z, b.a[r].w};
                Ar& t = acc[r][PARITY];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j < kIntWords) {
                        t = fma(HalfToDoubleScaled::low(w[j]), xv[2 * j], t);
                        t = fma(HalfToDoubleScaled::high(w[j]), xv[2 * j + 1], t);
                        chk = __hfma2(*reinterpret_cast<const __half2*>(&w[j]),
                                      zero, chk);
                    } else {
                        t = fma(half_bits_to_double(static_cast<unsigned short>(
                                    w[j] & 0xffffu)),
                                xv[2 * j], t);
                        t = fma(half_bits_to_double(
                                    static_cast<unsigned short>(w[j] >> 16)),
                                xv[2 * j + 1], t);
                    }
                }
            }
        } else {
            Ar xv[VEC];
            unpack_all<Ar, St>(b.x, xv);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    Ar& t = acc[r][PARITY * SPLIT + (i % SPLIT)];
                    t = fma_ar(unpack<Ar>(b.a[r], i, St{}), xv[i], t);
                }
            }
        }
    }

    // all full chunks (= pairs of batches) cw, cw + COLW, ... of the row group
    template <int COLW>
    static __device__ __forceinline__ void stream(
        const St* const (&row)[ROWS], const St* __restrict__ x,
        std::int64_t full_chunks, int cw, int lane, Ar (&acc)[ROWS][SLOTS],
        __half2& chk)
    {
        constexpr std::int64_t CHUNK = 2 * COLS;
        const std::int64_t lane_col = static_cast<std::int64_t>(lane) * VEC;
        std::int64_t k = cw;
        if (k >= full_chunks) {
            return;
        }
        Batch b0, b1;
        load(row, x, k * CHUNK + lane_col, b0);
        for (;;) {
            load(row, x, k * CHUNK + COLS + lane_col, b1);
            pin(b0);  // the loads above are issued BEFORE b0 is consumed
            compute<0>(b0, acc, chk);
            k += COLW;
            const bool more = k < full_chunks;
            if (more) {
                load(row, x, k * CHUNK + lane_col, b0);
            }
            pin(b1);
            compute<1>(b1, acc, chk);
            if (!more) {
                break;
            }
        }
    }
};

// Same batches, but the stream lands in shared memory: every lane copies the
// 16 bytes per row (and of x) that it will consume ITSELF with cp.async
// (LDGSTS, L1-bypassing for the matrix) into its own slots of a per-warp ring,
// so "has my data arrived" is a per-thread cp.async.wait_group -- no barrier,
// no mbarrier, no producer lane -- and the bytes in flight per warp are
// (STAGES - 1) x (ROWS + 1) x 512, independent of the register budget.
// CB = bytes per cp.async (16, 8 or 4): operands that are only 8- or 4-byte
// aligned (a row stride that is not a multiple of 16 bytes, e.g. fp16 rows of
// the reference driver's default stride 24500) are copied in 8- or 4-byte
// pieces into the SAME 16-byte shared-memory slots, so everything after the
// copy is unchanged -- no scalar kernel, no per-row peeling.
template <typename Ar, typename St, int ROWS, bool FAST, int IW, int STAGES,
          int CB = 16>
struct AsyncOps {
    using B = BatchOps<Ar, St, ROWS, FAST, IW, CB>;
    static constexpr int VEC = B::VEC;
    static constexpr int COLS = B::COLS;
    static constexpr unsigned STAGE_BYTES = (ROWS + 1) * 512;

    static __device__ __forceinline__ void copy16_stream(unsigned dst,
                                                         const void* src)
    {
        if constexpr (CB == 16) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst),
                         "l"(src)
                         : "memory");
        } else {
            copy16_cached(dst, src);  // pieces below 16 bytes exist as .ca only
        }
    }
    static __device__ __forceinline__ void copy16_cached(unsigned dst,
                                                         const void* src)
    {
        const char* s8 = static_cast<const char*>(src);
#pragma unroll
        for (int i = 0; i < 16 / CB; ++i) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(
                             dst + i * CB),
                         "l"(s8 + i * CB), "n"(CB)
                         : "memory");
        }
    }
    static __device__ __forceinline__ void commit()
    {
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    static __device__ __forceinline__ void wait_oldest()
    {
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
    }
    static __device__ __forceinline__ uint4 lds(unsigned addr)
    {
        uint4 r;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                     : "r"(addr));
        return r;
    }

    // `ring` = shared address of this LANE's 16 bytes in stage 0, row slot 0
    template <int COLW>
    static __device__ __forceinline__ void stream(
        const St* const (&row)[ROWS], const St* __restrict__ x,
        std::int64_t full_chunks, int cw, int lane,
        Ar (&acc)[ROWS][B::SLOTS], __half2& chk, unsigned ring)
    {
        constexpr std::int64_t CHUNK = 2 * COLS;
        const std::int64_t count =
            full_chunks > cw ? (full_chunks - cw + COLW - 1) / COLW : 0;
        const std::int64_t nb = 2 * count;  // batches of this warp
        const std::int64_t lane_col = static_cast<std::int64_t>(lane) * VEC;
        auto issue = [&](std::int64_t j, int slot) {
            const std::int64_t col =
                (cw + (j >> 1) * COLW) * CHUNK + (j & 1) * COLS + lane_col;
            const unsigned dst = ring + slot * STAGE_BYTES;
            copy16_cached(dst, x + col);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                copy16_stream(dst + (r + 1) * 512, row[r] + col);
            }
        };
        auto fetch = [&](int slot, typename B::Batch& b) {
            const unsigned src = ring + slot * STAGE_BYTES;
            b.x = lds(src);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                b.a[r] = lds(src + (r + 1) * 512);
            }
        };
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < nb) {
                issue(s, s);
            }
            commit();
        }
        int slot = 0;               // stage holding batch j
        int refill = STAGES - 1;    // stage batch j + STAGES - 1 goes to
        for (std::int64_t j = 0; j < nb; j += 2) {
            typename B::Batch b;
            wait_oldest();
            fetch(slot, b);
            if (j + STAGES - 1 < nb) {
                issue(j + STAGES - 1, refill);
            }
            commit();
            B::template compute<0>(b, acc, chk);
            slot = (slot + 1 == STAGES) ? 0 : slot + 1;
            refill = (refill + 1 == STAGES) ? 0 : refill + 1;

            wait_oldest();
            fetch(slot, b);
            if (j + STAGES < nb) {
                issue(j + STAGES, refill);
            }
            commit();
            B::template compute<1>(b, acc, chk);
            slot = (slot + 1 == STAGES) ? 0 : slot + 1;
            refill = (refill + 1 == STAGES) ? 0 : refill + 1;
        }
        // nothing may still be landing in the ring when the CTA moves on
        asm volatile("cp.async.wait_group 0;" ::: "memory");
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

// One row group of R rows: the COLW warps that share it walk the column chunks,
// reduce and write.  `part` is the CTA's [RG][COLW][MAXR] staging array.
template <typename St, typename Ar, int R, int MAXR, int UNROLL, int COLW,
          int PIPE, int IW, int CB>
__device__ __forceinline__ void gemv_row_group(
    std::int64_t m, std::int64_t n, Ar alpha, const St* __restrict__ A,
    std::int64_t lda, const St* __restrict__ x, Ar beta, St* __restrict__ y,
    std::int64_t incy, std::int64_t row0, int rg, int cw, int lane,
    Ar (*part)[COLW][MAXR], unsigned ring)
{
    constexpr bool FAST = use_scaled_half<Ar, St>::value;
    using Ops = ChunkOps<Ar, St, R, UNROLL, FAST>;
    using ExactOps = ChunkOps<Ar, St, R, UNROLL, false>;
    constexpr int CHUNK = Ops::CHUNK;
    constexpr int SLOTS = Ops::SLOTS;
    const bool active = row0 < m;

    // one accumulator per (row, unroll slot): short per-lane chains keep the
    // fp32-arithmetic rounding error at the level of the reference's
    // 512-partials-per-row tree
    Ar part_acc[R][SLOTS];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int u = 0; u < SLOTS; ++u) {
            part_acc[r][u] = Ar{};
        }
    }

    if (active) {
        const St* row[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            // rows past the end re-read the last valid row; never written
            const std::int64_t ri = (row0 + r < m) ? row0 + r : m - 1;
            row[r] = A + ri * lda;
        }
        const std::int64_t full_chunks = n / CHUNK;
        __half2 chk = __float2half2_rn(0.0f);
        if constexpr (PIPE >= 2 && UNROLL == 2) {
            AsyncOps<Ar, St, R, FAST, IW, PIPE, CB>::template stream<COLW>(
                row, x, full_chunks, cw, lane, part_acc, chk, ring);
        } else if constexpr (PIPE == 1 && UNROLL == 2) {
            // same chunk ownership and the same accumulator per (row, vector
            // slot) as the unpipelined loop: results are bit-identical
            BatchOps<Ar, St, R, FAST, IW, CB>::template stream<COLW>(
                row, x, full_chunks, cw, lane, part_acc, chk);
        } else {
            for (std::int64_t k = cw; k < full_chunks; k += COLW) {
                Ops::full(row, x, k * CHUNK, lane, part_acc, chk);
            }
        }
        if (FAST) {
            // a non-finite half went through the scaled path: redo this
            // warp's chunks with ordinary conversions (warp-uniform branch)
            const float2 c = __half22float2(chk);
            if (__any_sync(0xffffffffu, (c.x != c.x) || (c.y != c.y))) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int u = 0; u < SLOTS; ++u) {
                        part_acc[r][u] = Ar{};
                    }
                }
                for (std::int64_t k = cw; k < full_chunks; k += COLW) {
                    if constexpr (CB == 16) {
                        ExactOps::full(row, x, k * CHUNK, lane, part_acc, chk);
                    } else {
                        ExactOps::partial(row, x, k * CHUNK, n, lane, part_acc,
                                          false);
                    }
                }
            }
        }
        if (full_chunks % COLW == cw && full_chunks * CHUNK < n) {
            Ops::partial(row, x, full_chunks * CHUNK, n, lane, part_acc,
                         CB == 16);
        }
    }

    Ar acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        Ar v = part_acc[r][0];
#pragma unroll
        for (int u = 1; u < SLOTS; ++u) {
            v += part_acc[r][u];
        }
        acc[r] = warp_sum(v);
    }

    if (COLW == 1) {
        if (active && lane < R && row0 + lane < m) {
            Ar mine = acc[0];
#pragma unroll
            for (int r = 1; r < R; ++r) {
                mine = (lane == r) ? acc[r] : mine;
            }
            write_row<Ar, St>(y, (row0 + lane) * incy, alpha, beta, mine);
        }
    } else {
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                part[rg][cw][r] = acc[r];
            }
        }
        __syncthreads();
        if (active && cw == 0 && lane < R && row0 + lane < m) {
            Ar sum = part[rg][0][lane];
#pragma unroll
            for (int c = 1; c < COLW; ++c) {
                sum += part[rg][c][lane];
            }
            write_row<Ar, St>(y, (row0 + lane) * incy, alpha, beta, sum);
        }
    }
}

// Requirements (checked by the launcher): A and x 16-byte aligned,
// lda * sizeof(St) a multiple of 16, incx == 1.
// CTA = RG row groups x COLW column-splitting warps.
//
// Row groups are handed out in blockIdx order, which is also the order the
// hardware dispatches CTAs in.  CTAs below b_full own ROWS rows per group, the
// next b_half own ROWS/2 and the rest ROWS/4: the CTAs that start last are
// short, so the machine drains in a fraction of the time (measured tail of the
// uniform shape at 16384^2 fp32: 11 us of a 164 us kernel).  The size class is
// uniform per CTA, so each class runs its own unpredicated instantiation.
template <typename St, typename Ar, int ROWS, int UNROLL, int RG, int COLW,
          int MINB = (ROWS * UNROLL <= 8) ? 3 : 1, int PIPE = 0, int IW = 2,
          int CB = 16>
__global__ __launch_bounds__(RG* COLW* kWarp, MINB)
void gemv_stream_kernel(
    std::int64_t m, std::int64_t n, Ar alpha, const St* __restrict__ A,
    std::int64_t lda, const St* __restrict__ x, Ar beta, St* __restrict__ y,
    std::int64_t incy, std::int64_t b_full, std::int64_t b_half, int pdl)
{
    // Programmatic dependent launch: let the NEXT kernel of the stream become
    // resident as soon as every CTA of this grid has started (it fills the SM
    // slots our last wave leaves empty), and do not touch anything a previous
    // kernel may still be writing before griddepcontrol.wait.  Only requests
    // into L2 (always coherent) are issued ahead of the wait.
    if (pdl) {
        asm volatile("griddepcontrol.launch_dependents;");
    }
    constexpr int HALF = ROWS >= 2 ? ROWS / 2 : 1;
    constexpr int QUARTER = ROWS >= 4 ? ROWS / 4 : 1;
    __shared__ Ar part[RG][COLW][ROWS];
    extern __shared__ __align__(128) unsigned char gemv_ring[];
    const int lane = threadIdx.x & (kWarp - 1);
    const int warp = threadIdx.x >> 5;
    const int rg = warp / COLW;  // row group inside the CTA
    const int cw = warp % COLW;  // column slot inside the row group
    // cp.async ring of this warp (PIPE >= 2): PIPE stages of (ROWS + 1) x 512 B
    const unsigned ring =
        PIPE >= 2 ? static_cast<unsigned>(__cvta_generic_to_shared(gemv_ring)) +
                        warp * (PIPE * (ROWS + 1) * 512) + lane * 16
                  : 0u;
#if defined(ACCBLAS_DEV_HOOKS)
    unsigned long long* const trace = g_gemv_trace;
#else
    constexpr unsigned long long* trace = nullptr;
#endif
    if (trace != nullptr && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        trace[3 * blockIdx.x] = gemv_globaltimer();
        trace[3 * blockIdx.x + 2] = smid;
    }

    const std::int64_t b = blockIdx.x;
    if (pdl) {
        // first chunk of this warp's rows -> L2 while the predecessor drains
        constexpr int VEC = vec_traits<St>::elems;
        std::int64_t row0 = (b < b_full) ? (b * RG + rg) * ROWS : m;
        if (row0 < m && static_cast<std::int64_t>(cw + 1) * kWarp * VEC * UNROLL <= n) {
            const St* p0 = A + row0 * lda +
                           static_cast<std::int64_t>(cw) * kWarp * VEC * UNROLL;
            const int r = lane >> 3;  // ROWS <= 4 rows x 8 lines of 128 bytes
            const int line = lane & 7;
            if (r < ROWS && row0 + r < m &&
                line * 128 < kWarp * 16 * UNROLL) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(
                    reinterpret_cast<const char*>(p0 + r * lda) + line * 128));
            }
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    if (b < b_full) {
        const std::int64_t row0 = (b * RG + rg) * ROWS;
        gemv_row_group<St, Ar, ROWS, ROWS, UNROLL, COLW, PIPE, IW, CB>(
            m, n, alpha, A, lda, x, beta, y, incy, row0, rg, cw, lane, part, ring);
    } else if (b < b_full + b_half) {
        const std::int64_t row0 =
            b_full * RG * ROWS + ((b - b_full) * RG + rg) * HALF;
        gemv_row_group<St, Ar, HALF, ROWS, UNROLL, COLW, PIPE, IW, CB>(
            m, n, alpha, A, lda, x, beta, y, incy, row0, rg, cw, lane, part, ring);
    } else {
        const std::int64_t row0 = b_full * RG * ROWS + b_half * RG * HALF +
                                  ((b - b_full - b_half) * RG + rg) * QUARTER;
        gemv_row_group<St, Ar, QUARTER, ROWS, UNROLL, COLW, PIPE, IW, CB>(
            m, n, alpha, A, lda, x, beta, y, incy, row0, rg, cw, lane, part, ring);
    }
    if (trace != nullptr && threadIdx.x == 0) {
        trace[3 * blockIdx.x + 1] = gemv_globaltimer();
    }
}

// ---------------------------------------------------------------------------
// Bulk-copy pipelined variant.
//
// Same work decomposition and arithmetic as gemv_stream_kernel, but the matrix
// stream does not pass through registers on its way in: every warp owns a ring
// of STAGES shared-memory buffers and one lane feeds it with 1-D TMA bulk
// copies (cp.async.bulk, SASS UBLKCP), one per row and stage, completion
// signalled on an mbarrier with a transaction count.  The consumer lanes wait on
// the barrier, read their 16 bytes with LDS.128 and re-arm the stage.  Bytes in
// flight per warp = STAGES * ROWS * 512 * UNROLL, independent of the register
// budget, so the HBM pipe stays full while a warp is busy converting.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p)
{
    return static_cast<unsigned>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar),
                 "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                     bar),
                 "r"(bytes)
                 : "memory");
}

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

__device__ __forceinline__ void bulk_copy_g2s(unsigned dst, const void* src,
                                              unsigned bytes, unsigned bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

__device__ __forceinline__ uint4 lds_128(unsigned addr)
{
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "r"(addr));
    return r;
}

// PERSISTENT: the grid is sized to the machine (CTAs per SM x SMs) and every
// CTA walks row groups blockIdx.x, blockIdx.x + gridDim.x, ...  The producer
// lane keeps issuing copies ACROSS row-group boundaries, so the HBM stream
// never drains while a CTA reduces and writes its rows -- measured on B200 the
// per-CTA ramp of the non-persistent shape costs 8 % (fp32 rows) to 20 %
// (fp16 rows) of the bandwidth.
template <typename St, typename Ar, int ROWS, int UNROLL, int RG, int COLW,
          int STAGES>
__global__ __launch_bounds__(RG* COLW* kWarp) void gemv_bulk_kernel(
    std::int64_t m, std::int64_t n, Ar alpha, const St* __restrict__ A,
    std::int64_t lda, const St* __restrict__ x, Ar beta, St* __restrict__ y,
    std::int64_t incy)
{
    constexpr bool FAST = use_scaled_half<Ar, St>::value;
    using Ops = ChunkOps<Ar, St, ROWS, UNROLL, FAST>;
    using ExactOps = ChunkOps<Ar, St, ROWS, UNROLL, false>;
    constexpr int CHUNK = Ops::CHUNK;
    constexpr int VEC = Ops::VEC;
    constexpr int SLOTS = Ops::SLOTS;
    constexpr int WARPS = RG * COLW;
    constexpr unsigned ROW_BYTES = kWarp * 16 * UNROLL;
    constexpr unsigned STAGE_BYTES = ROWS * ROW_BYTES;

    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) unsigned long long bars[WARPS][STAGES];
    __shared__ Ar part[2][RG][COLW][ROWS];

    const int lane = threadIdx.x & (kWarp - 1);
    const int warp = threadIdx.x >> 5;
    const int rg = warp / COLW;
    const int cw = warp % COLW;

    const std::int64_t num_groups = (m + ROWS - 1) / ROWS;
    const std::int64_t cta_items = (num_groups + RG - 1) / RG;  // per launch
    // items of this CTA: blockIdx.x + j * gridDim.x
    const std::int64_t my_items =
        cta_items > blockIdx.x
            ? (cta_items - blockIdx.x + gridDim.x - 1) / gridDim.x
            : 0;
    const std::int64_t full_chunks = n / CHUNK;
    // chunks cw, cw + COLW, ... of every row group belong to this warp
    const std::int64_t count =
        full_chunks > cw ? (full_chunks - cw + COLW - 1) / COLW : 0;
    const std::int64_t total = my_items * count;  // ring transactions
    const unsigned my_ring = smem_u32(ring) + warp * STAGES * STAGE_BYTES;
    const unsigned my_bars = smem_u32(&bars[warp][0]);

    auto group_of = [&](std::int64_t j) {
        return (blockIdx.x + j * gridDim.x) * RG + rg;
    };
    auto row_ptr = [&](std::int64_t group, int r) {
        std::int64_t ri = group * ROWS + r;
        ri = ri < m ? ri : m - 1;  // past the end: re-read a valid row
        return A + ri * lda;
    };

    // producer state (lane 0): next transaction to issue
    std::int64_t issue_j = 0, issue_i = 0, issued = 0;
    auto issue_next = [&](int stage) {
        const std::int64_t c0 = (cw + issue_i * COLW) * CHUNK;
        const std::int64_t g = group_of(issue_j);
        const unsigned bar = my_bars + stage * 8;
        mbar_expect_tx(bar, STAGE_BYTES);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            bulk_copy_g2s(my_ring + stage * STAGE_BYTES + r * ROW_BYTES,
                          row_ptr(g, r) + c0, ROW_BYTES, bar);
        }
        ++issued;
        if (++issue_i == count) {
            issue_i = 0;
            ++issue_j;
        }
    };

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(my_bars + s * 8, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            if (issued < total) {
                issue_next(s);
            }
        }
    }
    __syncwarp();

    int stage = 0;
    unsigned parity = 0;
    for (std::int64_t j = 0; j < my_items; ++j) {
        const std::int64_t group = group_of(j);
        const std::int64_t row0 = group * ROWS;
        const bool active = row0 < m;

        Ar part_acc[ROWS][SLOTS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
                part_acc[r][u] = Ar{};
            }
        }
        __half2 chk = __float2half2_rn(0.0f);
        for (std::int64_t i = 0; i < count; ++i) {
            const std::int64_t c0 = (cw + i * COLW) * CHUNK;
            uint4 xr[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                xr[u] = ldg_cached_128(x + c0 + (u * kWarp + lane) * VEC);
            }
            mbar_wait(my_bars + stage * 8, parity);
            uint4 ar[ROWS][UNROLL];
            const unsigned base = my_ring + stage * STAGE_BYTES + lane * 16;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    ar[r][u] = lds_128(base + r * ROW_BYTES + u * kWarp * 16);
                }
            }
            Ops::compute(ar, xr, part_acc, chk);
            // every lane has consumed its data: the stage can be refilled,
            // possibly with the first chunk of the NEXT row group
            __syncwarp();
            if (lane == 0 && issued < total) {
                issue_next(stage);
            }
            if (++stage == STAGES) {
                stage = 0;
                parity ^= 1u;
            }
        }
        if (active) {
            const St* row[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                row[r] = row_ptr(group, r);
            }
            if (FAST) {
                const float2 c = __half22float2(chk);
                if (__any_sync(0xffffffffu, (c.x != c.x) || (c.y != c.y))) {
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) {
#pragma unroll
                        for (int u = 0; u < SLOTS; ++u) {
                            part_acc[r][u] = Ar{};
                        }
                    }
                    for (std::int64_t k = cw; k < full_chunks; k += COLW) {
                        ExactOps::full(row, x, k * CHUNK, lane, part_acc, chk);
                    }
                }
            }
            if (full_chunks % COLW == cw && full_chunks * CHUNK < n) {
                Ops::partial(row, x, full_chunks * CHUNK, n, lane, part_acc);
            }
        }

        Ar acc[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            Ar v = part_acc[r][0];
#pragma unroll
            for (int u = 1; u < SLOTS; ++u) {
                v += part_acc[r][u];
            }
            acc[r] = warp_sum(v);
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
            // double-buffered by item parity: one barrier per row group
            const int pb = static_cast<int>(j & 1);
            if (lane == 0) {
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    part[pb][rg][cw][r] = acc[r];
                }
            }
            __syncthreads();
            if (active && cw == 0 && lane < ROWS && row0 + lane < m) {
                Ar sum = part[pb][rg][0][lane];
#pragma unroll
                for (int c = 1; c < COLW; ++c) {
                    sum += part[pb][rg][c][lane];
                }
                write_row<Ar, St>(y, (row0 + lane) * incy, alpha, beta, sum);
            }
        }
    }
}

template <typename St, typename Ar, int ROWS, int UNROLL, int RG, int COLW,
          int STAGES>
int launch_bulk(Handle* h, std::int64_t m, std::int64_t n, Ar alpha,
                const St* A, std::int64_t lda, const St* x, Ar beta, St* y,
                std::int64_t incy, cudaStream_t stream)
{
    auto kernel = gemv_bulk_kernel<St, Ar, ROWS, UNROLL, RG, COLW, STAGES>;
    constexpr size_t smem =
        size_t{RG} * COLW * STAGES * ROWS * kWarp * 16 * UNROLL;
    // the shared-memory opt-in is per device (and per instantiation)
    static int per_device[64] = {};
    int device = 0;
    ACCBLAS_CUDA(cudaGetDevice(&device));
    const int slot = (device >= 0 && device < 64) ? device : 0;
    if (per_device[slot] == 0 || slot != device) {
        ACCBLAS_CUDA(cudaFuncSetAttribute(
            kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
            static_cast<int>(smem)));
        int occ = 0;
        ACCBLAS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &occ, kernel, RG * COLW * kWarp, smem));
        per_device[slot] = occ > 0 ? occ : 1;
    }
    const int ctas_per_sm = per_device[slot];
    const std::int64_t groups = (m + ROWS - 1) / ROWS;
    const std::int64_t items = (groups + RG - 1) / RG;
    int per_sm = tuning().gemv_ctas_per_sm;
    if (per_sm <= 0 || per_sm > ctas_per_sm) {
        per_sm = ctas_per_sm;
    }
    std::int64_t grid = std::int64_t{h->sm_count} * per_sm;
    if (grid > items) {
        grid = items;
    }
    kernel<<<static_cast<unsigned>(grid), RG * COLW * kWarp, smem, stream>>>(
        m, n, alpha, A, lda, x, beta, y, incy);
    ACCBLAS_CUDA(cudaGetLastError());
    return ACCBLAS_OK;
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

template <typename St, typename Ar, int ROWS, int UNROLL, int RG, int COLW,
          int MINB = (ROWS * UNROLL <= 8) ? 3 : 1, int PIPE = 0, int IW = 2,
          int CB = 16>
int launch_stream(Handle* h, std::int64_t m, std::int64_t n, Ar alpha,
                  const St* A, std::int64_t lda, const St* x, Ar beta, St* y,
                  std::int64_t incy, cudaStream_t stream)
{
    auto kernel =
        gemv_stream_kernel<St, Ar, ROWS, UNROLL, RG, COLW, MINB, PIPE, IW, CB>;
    constexpr size_t smem =
        PIPE >= 2 ? size_t{RG} * COLW * PIPE * (ROWS + 1) * 512 : 0;
    static int resident_on[64] = {};  // CTAs per SM, per instantiation and device
    int device = 0;
    ACCBLAS_CUDA(cudaGetDevice(&device));
    const int slot = (device >= 0 && device < 64) ? device : 0;
    if (resident_on[slot] == 0 || slot != device) {
        if (smem > 0) {
            ACCBLAS_CUDA(cudaFuncSetAttribute(
                kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                static_cast<int>(smem)));
        }
        int occ = 0;
        ACCBLAS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &occ, kernel, RG * COLW * kWarp, smem));
        resident_on[slot] = occ > 0 ? occ : 1;
    }
    const int resident = resident_on[slot];
    // taper: the last two "waves" of CTAs own half / a quarter of the rows
    constexpr int HALF = ROWS >= 2 ? ROWS / 2 : 1;
    constexpr int QUARTER = ROWS >= 4 ? ROWS / 4 : 1;
    const std::int64_t slots = std::int64_t{h->sm_count} * resident;  // CTAs
    std::int64_t b_full, b_half, b_quarter;
    if (ROWS >= 4 && tuning().gemv_taper != 0 && m >= 4 * slots * RG * ROWS) {
        b_half = slots;
        b_full = (m - slots * RG * (HALF + QUARTER)) / (RG * ROWS);
        const std::int64_t rest = m - b_full * RG * ROWS - b_half * RG * HALF;
        b_quarter = (rest + RG * QUARTER - 1) / (RG * QUARTER);
    } else {
        b_full = (m + RG * ROWS - 1) / (RG * ROWS);
        b_half = b_quarter = 0;
    }
    const std::int64_t grid = b_full + b_half + b_quarter;
    if (grid > 0x7fffffffLL) {
        set_error("gemv: too many rows (%lld)", static_cast<long long>(m));
        return ACCBLAS_ERR_INVALID;
    }
    const int pdl = tuning().gemv_pdl;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(RG * COLW * kWarp);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    ACCBLAS_CUDA(cudaLaunchKernelEx(&cfg, kernel, m, n, alpha, A, lda, x, beta,
                                    y, incy, b_full, b_half, pdl));
    return ACCBLAS_OK;
}

// variant: 2 = CTA owns 2 rows, 8 warps split the columns
//          3 = CTA owns 1 row,  8 warps split the columns
//          4 = CTA owns 4 rows, 8 warps split the columns (default shape)
//          5 = CTA owns 8 rows (2 groups of 4), 4 warps per group
// (shapes that lost everywhere on B200 -- a warp owning whole rows, 16-row
// CTAs, 8-row groups -- are gone; profiles/r01_summary.md has their numbers)
template <typename St, typename Ar, int UNROLL>
int launch_variant(Handle* h, int variant, std::int64_t m, std::int64_t n,
                   Ar alpha,
                   const St* A, std::int64_t lda, const St* x, Ar beta, St* y,
                   std::int64_t incy, cudaStream_t stream)
{
    constexpr bool FAST = use_scaled_half<Ar, St>::value;
    if constexpr (UNROLL == 2) {
        // bulk-copy pipeline (default shape only): ring depth 2 / 3 / 4
        switch (variant == 4 ? tuning().gemv_stages : 0) {
        case 2:
            return launch_bulk<St, Ar, 4, 2, 1, 8, 2>(h, m, n, alpha, A, lda, x,
                                                      beta, y, incy, stream);
        case 3:
            return launch_bulk<St, Ar, 4, 2, 1, 8, 3>(h, m, n, alpha, A, lda, x,
                                                      beta, y, incy, stream);
        case 4:
            return launch_bulk<St, Ar, 4, 2, 1, 8, 4>(h, m, n, alpha, A, lda, x,
                                                      beta, y, incy, stream);
        default:
            break;
        }
        switch (variant) {
        case 2:
            return launch_stream<St, Ar, 2, 2, 1, 8>(h, m, n, alpha, A, lda, x,
                                                     beta, y, incy, stream);
        case 5:
            return launch_stream<St, Ar, 4, 2, 2, 4>(h, m, n, alpha, A, lda, x,
                                                     beta, y, incy, stream);
        default:
            break;
        }
    }
    if (variant == 3) {
        return launch_stream<St, Ar, 1, UNROLL, 1, 8>(h, m, n, alpha, A, lda, x,
                                                      beta, y, incy, stream);
    }
    if constexpr (UNROLL == 2) {
        int pipe = tuning().gemv_pipe;
        if (pipe < 0) {
            // measured on B200 (same box, min of 10): keeping one batch in
            // flight while the previous is consumed gains 5-7 % for
            // Acc<fp64,fp16> (instruction-heavy) and loses 1-8 % elsewhere
            pipe = FAST ? 1 : 0;
        }
        if (pipe == 3) {
            return launch_stream<St, Ar, 4, 2, 1, 8, 3, 3>(
                h, m, n, alpha, A, lda, x, beta, y, incy, stream);
        }
        if (pipe == 4) {
            return launch_stream<St, Ar, 4, 2, 1, 8, 2, 4>(
                h, m, n, alpha, A, lda, x, beta, y, incy, stream);
        }
        if (pipe != 0) {
            if constexpr (FAST) {
                switch (tuning().gemv_intwords) {
                case 0:
                    return launch_stream<St, Ar, 4, 2, 1, 8, 3, 1, 0>(
                        h, m, n, alpha, A, lda, x, beta, y, incy, stream);
                case 1:
                    return launch_stream<St, Ar, 4, 2, 1, 8, 3, 1, 1>(
                        h, m, n, alpha, A, lda, x, beta, y, incy, stream);
                case 3:
                    return launch_stream<St, Ar, 4, 2, 1, 8, 3, 1, 3>(
                        h, m, n, alpha, A, lda, x, beta, y, incy, stream);
                case 4:
                    return launch_stream<St, Ar, 4, 2, 1, 8, 3, 1, 4>(
                        h, m, n, alpha, A, lda, x, beta, y, incy, stream);
                default:
                    break;
                }
            }
            return launch_stream<St, Ar, 4, 2, 1, 8, 3, 1>(
                h, m, n, alpha, A, lda, x, beta, y, incy, stream);
        }
    }
    return launch_stream<St, Ar, 4, UNROLL, 1, 8>(h, m, n, alpha, A, lda, x,
                                                  beta, y, incy, stream);
}

// x with a stride: packed once into the workspace (n elements, a few
// microseconds at most), so that the matrix stream -- m times larger -- keeps
// its vector path
template <typename St>
__global__ __launch_bounds__(256) void gather_strided_kernel(
    const St* __restrict__ x, std::int64_t incx, St* __restrict__ out,
    std::int64_t n)
{
    const std::int64_t i = std::int64_t{blockIdx.x} * 256 + threadIdx.x;
    if (i < n) {
        out[i] = x[i * incx];
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
    if (incx != 1 && n > 0) {
        const size_t bytes = static_cast<size_t>(n) * sizeof(St);
        int rc = ensure_workspace(h, kScratchBytes + bytes, stream);
        if (rc != ACCBLAS_OK) {
            return rc;
        }
        // the region doubles as the TRSV progress vector: it has to be armed
        // again before the next solve
        h->trsv_armed_bytes = 0;
        St* packed = static_cast<St*>(trsv_region(h));
        gather_strided_kernel<St>
            <<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
                x, incx, packed, n);
        ACCBLAS_CUDA(cudaGetLastError());
        x = packed;
        incx = 1;
    }
    const bool vec_ok =
        incx == 1 &&
        ((reinterpret_cast<std::uintptr_t>(A) |
          reinterpret_cast<std::uintptr_t>(x)) & 15u) == 0 &&
        (static_cast<std::uint64_t>(lda) * sizeof(St)) % 16 == 0;
    {
        // Acc<fp64,fp16>: the register pipeline fed by 64-bit loads (finer
        // grained: half a vector per request) beats the 128-bit one by 5 % on
        // aligned data too (same box: 5826 -> 6100-6170 GB/s); every other
        // pair is 2-8 % faster with 128-bit loads.  gemv_force_pieces: 8 = all
        // pairs through the 64-bit pipeline, -1 = none (for A/B runs).
        const Tuning& t = tuning();
        const bool fast_default =
            use_scaled_half<Ar, St>::value && t.gemv_force_pieces == 0 &&
            t.gemv_pipe < 0 && t.gemv_variant == 0 && t.gemv_unroll == 2 &&
            t.gemv_stages == 0 && m >= std::int64_t{4} * 8 * h->sm_count;
        if (vec_ok && (t.gemv_force_pieces == 8 || fast_default)) {
            if constexpr (use_scaled_half<Ar, St>::value) {
                if (t.gemv_intwords == 1) {
                    return launch_stream<St, Ar, 4, 2, 1, 8, 3, 1, 1, 8>(
                        h, m, n, alpha, A, lda, x, beta, y, incy, stream);
                }
                if (t.gemv_intwords == 3) {
                    return launch_stream<St, Ar, 4, 2, 1, 8, 3, 1, 3, 8>(
                        h, m, n, alpha, A, lda, x, beta, y, incy, stream);
                }
            }
            return launch_stream<St, Ar, 4, 2, 1, 8, 3, 1, 2, 8>(
                h, m, n, alpha, A, lda, x, beta, y, incy, stream);
        }
    }
    if (vec_ok) {
        int variant = tuning().gemv_variant;
        int unroll = tuning().gemv_unroll;
        if (variant == 0) {
            // rows are cheap to split: prefer 4-row CTAs while that still
            // gives every SM several CTAs, then 2-row, then 1-row CTAs
            const std::int64_t sms = h->sm_count;
            variant = (m >= 4 * 8 * sms) ? 4 : (m >= 2 * 4 * sms) ? 2 : 3;
        }
        if (unroll == 0) {
            unroll = 2;
        }
        switch (unroll) {
        case 1:
            return launch_variant<St, Ar, 1>(h, variant, m, n, alpha, A, lda, x,
                                             beta, y, incy, stream);
        case 4:
            return launch_variant<St, Ar, 4>(h, variant, m, n, alpha, A, lda, x,
                                             beta, y, incy, stream);
        default:
            return launch_variant<St, Ar, 2>(h, variant, m, n, alpha, A, lda, x,
                                             beta, y, incy, stream);
        }
    }
    if (incx == 1) {
        // 8- or 4-byte aligned operands: the cp.async ring with smaller pieces
        const std::uintptr_t bits =
            reinterpret_cast<std::uintptr_t>(A) |
            reinterpret_cast<std::uintptr_t>(x) |
            static_cast<std::uintptr_t>(static_cast<std::uint64_t>(lda) *
                                        sizeof(St));
        const bool ring = tuning().gemv_pipe >= 2;
        if ((bits & 7u) == 0) {
            return ring ? launch_stream<St, Ar, 4, 2, 1, 8, 3, 3, 2, 8>(
                              h, m, n, alpha, A, lda, x, beta, y, incy, stream)
                        : launch_stream<St, Ar, 4, 2, 1, 8, 3, 1, 2, 8>(
                              h, m, n, alpha, A, lda, x, beta, y, incy, stream);
        }
        if ((bits & 3u) == 0) {
            return ring ? launch_stream<St, Ar, 4, 2, 1, 8, 3, 3, 2, 4>(
                              h, m, n, alpha, A, lda, x, beta, y, incy, stream)
                        : launch_stream<St, Ar, 4, 2, 1, 8, 3, 1, 2, 4>(
                              h, m, n, alpha, A, lda, x, beta, y, incy, stream);
        }
        if constexpr (sizeof(St) == 2) {
            // fp16 with an odd stride or base (element aligned only): the same
            // pipeline, every 16 bytes fetched as 32-bit words (plus two
            // half-words when they start two bytes past a 4-byte boundary)
            return launch_stream<St, Ar, 4, 2, 1, 8, 3, 1, 2, 2>(
                h, m, n, alpha, A, lda, x, beta, y, incy, stream);
        }
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
