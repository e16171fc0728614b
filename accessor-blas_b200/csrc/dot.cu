// DOT over reduced-precision storage, deterministic two-pass reduction.
//
// Replaces kernel::acc_dot / kernel::dot + atomic_add + init_res + cast_result
// (/root/reference/cuda/dot_kernels.cuh:78-173, cuda/atomics.cuh:80-110) with
// ONE launch:
//   pass 1  every CTA streams whole tiles of x and y with 128-bit
//           L1-bypassing loads (UNROLL vectors of each operand in flight per
//           thread), converts storage -> arithmetic in registers, accumulates
//           with FMA, reduces with a fixed warp/CTA tree and writes its partial
//           to the workspace;
//   pass 2  the CTA that finishes last (completion counter) folds the partials
//           in a fixed order with the same tree, converts to the result type
//           and re-arms the counter.
// The summation order depends only on (n, dtype pair, launch shape), never on
// which CTA happens to be last, so the result is bit-reproducible run to run.
//
// Operand layouts on the vector path: contiguous operands with the same
// misalignment are peeled to the first 16-byte boundary; contiguous operands
// with DIFFERENT misalignments stream the better aligned one with 128-bit
// loads and fetch the other one's 16 bytes in 8- / 4- / 2-byte pieces (same
// registers afterwards, so the arithmetic and its order are those of the
// aligned kernel).  Strided operands use scalar loads, four of each operand in
// flight per thread.
#include "common.cuh"
#include "tuning.h"

namespace accblas {
namespace {

// acc += x * y in Ar.  For fp16 storage with fp64 arithmetic the product of
// two halves is exact in fp32 (22 significant bits, |p| in [2^-48, 2^32)), so
// it is formed with one FMUL and widened once: bit-identical to
// fma(double(x), double(y), acc) at half the 64-bit conversions.
template <typename Ar, typename St>
struct pair_fma {
    static __device__ __forceinline__ Ar apply(St a, St b, Ar acc)
    {
        return fma_ar(to_ar<Ar, St>(a), to_ar<Ar, St>(b), acc);
    }
};
template <>
struct pair_fma<double, __half> {
    static __device__ __forceinline__ double apply(__half a, __half b,
                                                   double acc)
    {
        const float p = __fmul_rn(__half2float(a), __half2float(b));
        return __dadd_rn(acc, static_cast<double>(p));
    }
};

// fp32 -> fp64 off the conversion pipe (experiment, `dot_intmix`).  The 32
// bits of a float f, shifted right by three into the high word of a double
// (sign kept at bit 31, the three replicated sign bits cleared), read as
//     d = value(f) * 2^-896
// for zero, subnormal and normal f alike: the 8-bit exponent lands in the low
// 8 bits of the 11-bit field.  One signed 32x32->64 multiply by 2^29 builds the
// register pair, one mask cleans the high word.  x goes this way, y through
// F2F; the accumulator then holds sum * 2^-896 and is scaled back once per
// thread (exact).  Inf/NaN inputs do not map: 0 * f on the fp32 pipe is NaN
// exactly for those, and the CTA falls back to ordinary conversions.
struct FloatToDoubleScaled {
    static __device__ __forceinline__ double widen(unsigned f)
    {
        long long p;
        asm("mul.wide.s32 %0, %1, 0x20000000;" : "=l"(p) : "r"(f));
        return __longlong_as_double(
            p & static_cast<long long>(0x8FFFFFFFFFFFFFFFull));
    }
    static __device__ __forceinline__ double unscale()
    {
        return __hiloint2double((896 + 1023) << 20, 0);  // 2^896
    }
};

template <typename St>
__device__ __forceinline__ St raw_elem(const uint4& raw, int i);
template <>
__device__ __forceinline__ double raw_elem<double>(const uint4& raw, int i)
{
    return (i == 0) ? __hiloint2double(raw.y, raw.x)
                    : __hiloint2double(raw.w, raw.z);
}
template <>
__device__ __forceinline__ float raw_elem<float>(const uint4& raw, int i)
{
    return __uint_as_float((i == 0)   ? raw.x
                           : (i == 1) ? raw.y
                           : (i == 2) ? raw.z
                                      : raw.w);
}
template <>
__device__ __forceinline__ __half raw_elem<__half>(const uint4& raw, int i)
{
    const int word = i >> 1;
    const unsigned w = (word == 0)   ? raw.x
                       : (word == 1) ? raw.y
                       : (word == 2) ? raw.z
                                     : raw.w;
    return __ushort_as_half(
        static_cast<unsigned short>((i & 1) ? (w >> 16) : (w & 0xffffu)));
}

template <typename Ar>
__device__ __forceinline__ void store_result(void* result, int res_dtype, Ar v)
{
    switch (res_dtype) {
    case ACCBLAS_F64:
        *static_cast<double*>(result) = static_cast<double>(v);
        break;
    case ACCBLAS_F32:
        *static_cast<float*>(result) = static_cast<float>(v);
        break;
    default:
        *static_cast<__half*>(result) = to_st<__half, Ar>(v);
        break;
    }
}

__device__ __forceinline__ unsigned long long dot_globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Fused all-reduce of the per-GPU partials (one thread per peer): store my
// partial into every peer's mailbox (P2P stores over NVLink), then wait for the
// peer's entry in mine.  Returns the sum in rank order (valid in thread 0).
// A peer that does not show up within px.timeout_ns yields NaN on this rank
// AND raises the handle's sticky failure word (mapped host memory), so the
// next call on the handle returns ACCBLAS_ERR_PEER instead of launching.
template <typename Ar>
__device__ __forceinline__ Ar peer_allreduce(Ar mine, const PeerExchange& px)
{
    __shared__ double peer_vals[kMaxPeers];
    __shared__ Ar mine_s;
    if (threadIdx.x == 0) {
        mine_s = mine;
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t < px.world) {
        const int half = static_cast<int>(px.epoch & 1ull) * kMaxPeers;
        const double payload = static_cast<double>(mine_s);  // exact for float
        volatile unsigned long long* out =
            static_cast<volatile unsigned long long*>(px.mailbox[t]) +
            2 * (half + px.rank);
        out[0] = static_cast<unsigned long long>(__double_as_longlong(payload));
        __threadfence_system();  // value before flag, system scope (peer GPU)
        out[1] = px.epoch;
        volatile unsigned long long* in =
            static_cast<volatile unsigned long long*>(px.mailbox[px.rank]) +
            2 * (half + t);
        const unsigned long long t0 = dot_globaltimer_ns();
        bool ok = true;
        while (in[1] != px.epoch) {
            if (px.timeout_ns != 0 &&
                dot_globaltimer_ns() - t0 > px.timeout_ns) {
                ok = false;
                break;
            }
        }
        __threadfence_system();  // flag before value
        peer_vals[t] = ok ? __longlong_as_double(static_cast<long long>(in[0]))
                          : __longlong_as_double(0x7ff8000000000000LL);
        if (!ok && px.status != nullptr) {
            *reinterpret_cast<volatile unsigned long long*>(px.status) =
                px.epoch;
            __threadfence_system();
        }
    }
    __syncthreads();
    Ar total = Ar{};
    if (t == 0) {
        for (int r = 0; r < px.world; ++r) {
            total += static_cast<Ar>(peer_vals[r]);
        }
    }
    return total;
}

// Second pass + epilogue shared by all first-pass kernels.
template <typename Ar, int BLOCK>
__device__ __forceinline__ void finish_dot(Ar local, Ar* partials,
                                           unsigned* counter, void* result,
                                           int res_dtype, Ar* scratch,
                                           const PeerExchange& px,
                                           unsigned num_partials = 0)
{
    // partials[0 .. gridDim.x) = one per CTA; the streaming kernel appends one
    // per dynamically assigned chunk (num_partials > gridDim.x)
    num_partials = num_partials ? num_partials : gridDim.x;
    __shared__ bool is_last;
    const Ar total = block_sum(local, scratch);
    if (threadIdx.x == 0) {
        // volatile store + fence: the partial must be visible device-wide
        // before the counter increment is
        *reinterpret_cast<volatile Ar*>(partials + blockIdx.x) = total;
        __threadfence();
        const unsigned ticket = atomicAdd(counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) {
        return;
    }
    __threadfence();
    // The fold of the (up to ~5000) partials runs in fp64 also for fp32
    // arithmetic and is rounded to Ar once: one CTA, a few values per thread --
    // free, and the partials of the dynamic pool do not add fp32 rounding
    // steps to the result.
    __shared__ double wide_scratch[kWarp];
    // (eight loads in flight per thread: one after the other, the ~16 L2 round
    // trips of a 4000-entry fold were ~1 % of a 600 us call; the order of the
    // additions is unchanged)
    double v = 0.0;
    constexpr unsigned kFoldLoads = 8;
    for (unsigned base = threadIdx.x; base < num_partials;
         base += BLOCK * kFoldLoads) {
        Ar t[kFoldLoads];
#pragma unroll
        for (unsigned j = 0; j < kFoldLoads; ++j) {
            const unsigned i = base + j * BLOCK;
            t[j] = i < num_partials ? __ldcg(partials + i) : Ar{};
        }
#pragma unroll
        for (unsigned j = 0; j < kFoldLoads; ++j) {
            if (base + j * BLOCK < num_partials) {
                v += static_cast<double>(t[j]);
            }
        }
    }
    Ar sum = static_cast<Ar>(block_sum(v, wide_scratch));
    if (px.world > 1) {
        sum = peer_allreduce(sum, px);
    }
    if (threadIdx.x == 0) {
        store_result(result, res_dtype, sum);
        *counter = 0u;  // re-arm for the next call on this handle
        counter[kCtlDotChunk - kCtlDotCounter] = 0u;
    }
}

// Contiguous operands.  x is 16-byte aligned (after peeling `head` elements);
// y is aligned to CBY bytes (16 = the same alignment as x).
// MIX (fp32 storage, fp64 arithmetic only): x widened on the integer pipes.
template <typename St, typename Ar, int BLOCK, int UNROLL, int CBY, bool MIX>
__global__ __launch_bounds__(BLOCK) void dot_stream_kernel(
    const St* __restrict__ x, const St* __restrict__ y, std::int64_t n,
    Ar* __restrict__ partials, unsigned* __restrict__ counter,
    void* __restrict__ result, int res_dtype, const PeerExchange px, int pdl,
    int head, std::int64_t static_tiles, int chunk_tiles, unsigned num_chunks)
{
    // x and y point at the element that falls on x's first 16-byte boundary;
    // `head` (< 16 / sizeof(St)) elements in front of it belong to the operands
    // as well (e.g. x[1:] . y[1:])
    constexpr int VEC = vec_traits<St>::elems;
    constexpr std::int64_t TILE = std::int64_t{BLOCK} * VEC * UNROLL;
    __shared__ Ar scratch[kWarp];
    if (pdl) {
        // programmatic dependent launch (see gemv.cu): the next kernel of the
        // stream may become resident while this grid drains; nothing but L2
        // prefetches of the first tile happens before the predecessor is done
        asm volatile("griddepcontrol.launch_dependents;");
        if (static_cast<std::int64_t>(blockIdx.x) < n / TILE) {
            const std::int64_t first =
                blockIdx.x * TILE + std::int64_t{threadIdx.x} * VEC;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if ((threadIdx.x & 7) == 0) {  // one request per 128-byte line
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(
                        x + first + std::int64_t{u} * BLOCK * VEC));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(
                        y + first + std::int64_t{u} * BLOCK * VEC));
                }
            }
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }

    const std::int64_t num_tiles = n / TILE;
    // fp32 arithmetic: one accumulator per (vector, element) slot, so a chain
    // is only as long as the number of tiles this CTA walks and the error
    // stays below the reference's (thread chains of ~55 at n = 2^28).  fp64
    // arithmetic: one per vector is plenty.
    constexpr int SLOTS = std::is_same<Ar, float>::value ? (VEC < 4 ? VEC : 4) : 1;
    Ar acc[UNROLL][SLOTS];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) {
            acc[u][i] = Ar{};
        }
    }
    float chk = 0.0f;  // MIX: NaN iff an Inf/NaN went through the scaled path

    auto consume_tile = [&](std::int64_t tile) {
        const std::int64_t base = tile * TILE + std::int64_t{threadIdx.x} * VEC;
        uint4 xr[UNROLL];
        uint4 yr[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            xr[u] = ldg_stream_128(x + base + std::int64_t{u} * BLOCK * VEC);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            yr[u] = ldg_pieces<CBY, true>(y + base +
                                          std::int64_t{u} * BLOCK * VEC);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if constexpr (MIX) {
                const unsigned xw[4] = {xr[u].x, xr[u].y, xr[u].z, xr[u].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double xs = FloatToDoubleScaled::widen(xw[i]);
                    acc[u][0] = fma(
                        xs, static_cast<double>(raw_elem<float>(yr[u], i)),
                        acc[u][0]);
                    chk = fmaf(__uint_as_float(xw[i]), 0.0f, chk);
                }
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    acc[u][i % SLOTS] = pair_fma<Ar, St>::apply(
                        raw_elem<St>(xr[u], i), raw_elem<St>(yr[u], i),
                        acc[u][i % SLOTS]);
                }
            }
        }
    };
    // Static part: tiles [0, static_tiles) in grid-stride order; the rest of
    // the tiles is handed out dynamically further down.
    for (std::int64_t tile = blockIdx.x; tile < static_tiles;
         tile += gridDim.x) {
        consume_tile(tile);
    }
    if constexpr (MIX) {
        // (block-uniform) an Inf/NaN was consumed by the scaled conversion:
        // redo this CTA's tiles with ordinary conversions
        const bool bad = __syncthreads_or(chk != chk);
        if (bad) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                acc[u][0] = 0.0;
            }
            for (std::int64_t tile = blockIdx.x; tile < static_tiles;
                 tile += gridDim.x) {
                const std::int64_t base =
                    tile * TILE + std::int64_t{threadIdx.x} * VEC;
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const uint4 xv =
                        ldg_stream_128(x + base + std::int64_t{u} * BLOCK * VEC);
                    const uint4 yv = ldg_pieces<CBY, true>(
                        y + base + std::int64_t{u} * BLOCK * VEC);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        acc[u][0] = pair_fma<Ar, St>::apply(
                            raw_elem<St>(xv, i), raw_elem<St>(yv, i), acc[u][0]);
                    }
                }
            }
        } else {
            const double s = FloatToDoubleScaled::unscale();
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                acc[u][0] = acc[u][0] * s;
            }
        }
    }

    // ragged tail (< TILE elements), spread over the whole grid
    Ar tail = Ar{};
    if (blockIdx.x == 0 && static_cast<int>(threadIdx.x) < head) {
        const std::int64_t i = static_cast<std::int64_t>(threadIdx.x) - head;
        tail = pair_fma<Ar, St>::apply(x[i], y[i], tail);
    }
    for (std::int64_t i = num_tiles * TILE +
                          std::int64_t{blockIdx.x} * BLOCK + threadIdx.x;
         i < n; i += std::int64_t{gridDim.x} * BLOCK) {
        tail = pair_fma<Ar, St>::apply(x[i], y[i], tail);
    }

    // fixed-order fold of the per-thread accumulators
    auto fold_acc = [&]() {
        Ar sum = Ar{};
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            // pairwise fold of the element slots, then the vectors in order
            Ar v[SLOTS];
#pragma unroll
            for (int i = 0; i < SLOTS; ++i) {
                v[i] = acc[u][i];
            }
#pragma unroll
            for (int width = SLOTS / 2; width > 0; width /= 2) {
#pragma unroll
                for (int i = 0; i < width; ++i) {
                    v[i] += v[i + width];
                }
            }
            sum += v[0];
        }
        return sum;
    };
    Ar local = fold_acc();
    local += tail;

    // Dynamic part.  Equal shares are as slow as the slowest SM: on B200 the
    // SMs do not all get the same share of the memory system (a static
    // partition measured 6970 GB/s for fp64 storage where the reference's
    // 32-waves-of-small-blocks kernel, balanced by the block scheduler, reached
    // 7360 on the same box; profiles/r02_dot_dynamic_ab.txt).  The last ~12 %
    // of the tiles are therefore handed out in small chunks by an atomic
    // counter: SMs that finish their static share early take more of them.
    // Every chunk has its OWN partial sum, folded in chunk order at the end, so
    // the result does not depend on which CTA happened to take which chunk.
    if (chunk_tiles > 0) {
        // One barrier per chunk: the id of the NEXT chunk is fetched while the
        // current one is streamed (the atomic's round trip stays off the
        // path), and the warps' sums of a chunk are added up by thread 0 after
        // the barrier that publishes that id.
        constexpr int NW = BLOCK / kWarp;
        __shared__ unsigned chunk_id[2];
        __shared__ Ar warp_part[2][NW];
        unsigned* const chunk_counter =
            counter + (kCtlDotChunk - kCtlDotCounter);
        const int lane = threadIdx.x & (kWarp - 1);
        const int warp = threadIdx.x >> 5;
        if (threadIdx.x == 0) {
            chunk_id[0] = atomicAdd(chunk_counter, 1u);
        }
        __syncthreads();
        for (int k = 0;; ++k) {
            const int slot = k & 1;
            const unsigned c = chunk_id[slot];
            if (c >= num_chunks) {
                break;
            }
            if (threadIdx.x == 0) {
                chunk_id[slot ^ 1] = atomicAdd(chunk_counter, 1u);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
                for (int i = 0; i < SLOTS; ++i) {
                    acc[u][i] = Ar{};
                }
            }
            // (chunks stay SMALL: the static grid-stride order interleaves
            // the CTAs tile by tile, and so do small chunks; handing out the
            // static share in chunks of 16-64 tiles instead -- every CTA alone
            // in its own 0.25-1 MB -- cost fp32 / fp16 storage 5-40 %,
            // profiles/r02_dot_pool_sweep.txt)
            const std::int64_t t0 = static_tiles + std::int64_t{c} * chunk_tiles;
            const std::int64_t t1 =
                t0 + chunk_tiles < num_tiles ? t0 + chunk_tiles : num_tiles;
            for (std::int64_t tile = t0; tile < t1; ++tile) {
                consume_tile(tile);
            }
            const Ar wsum = warp_sum(fold_acc());
            if (lane == 0) {
                warp_part[slot][warp] = wsum;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                Ar part = Ar{};
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    part += warp_part[slot][w];
                }
                *reinterpret_cast<volatile Ar*>(partials + gridDim.x + c) = part;
            }
        }
    }
    finish_dot<Ar, BLOCK>(local, partials, counter, result, res_dtype, scratch,
                          px, gridDim.x + num_chunks);
}

// Any stride: scalar loads, four of each operand in flight per thread, 64-bit
// indices (the reference's plain kernel is limited to int32,
// cuda/dot_kernels.cuh:89-97).
template <typename St, typename Ar, int BLOCK>
__global__ __launch_bounds__(BLOCK) void dot_strided_kernel(
    const St* __restrict__ x, std::int64_t incx, const St* __restrict__ y,
    std::int64_t incy, std::int64_t n, Ar* __restrict__ partials,
    unsigned* __restrict__ counter, void* __restrict__ result, int res_dtype,
    const PeerExchange px)
{
    constexpr int U = 4;
    __shared__ Ar scratch[kWarp];
    Ar acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        acc[u] = Ar{};
    }
    const std::int64_t step = std::int64_t{gridDim.x} * BLOCK;
    std::int64_t i = std::int64_t{blockIdx.x} * BLOCK + threadIdx.x;
    for (; i + (U - 1) * step < n; i += U * step) {
        St xa[U], ya[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            xa[u] = x[(i + u * step) * incx];
            ya[u] = y[(i + u * step) * incy];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            acc[u] = pair_fma<Ar, St>::apply(xa[u], ya[u], acc[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (i + u * step < n) {
            acc[u] = pair_fma<Ar, St>::apply(x[(i + u * step) * incx],
                                             y[(i + u * step) * incy], acc[u]);
        }
    }
    finish_dot<Ar, BLOCK>((acc[0] + acc[1]) + (acc[2] + acc[3]), partials,
                          counter, result, res_dtype, scratch, px);
}

template <typename St, typename Ar, int BLOCK, int UNROLL, int CBY, bool MIX>
int launch_stream(Handle* h, std::int64_t n, const void* x, const void* y,
                  void* result, int res, int ctas_per_sm, cudaStream_t stream,
                  const PeerExchange& px, int head)
{
    constexpr int VEC = vec_traits<St>::elems;
    constexpr std::int64_t TILE = std::int64_t{BLOCK} * VEC * UNROLL;
    auto kernel = dot_stream_kernel<St, Ar, BLOCK, UNROLL, CBY, MIX>;
    // one resident wave: every CTA of the grid-stride loop is on the machine
    // from start to end (a partial second wave would leave a tail).  The
    // occupancy is a property of (instantiation, device).
    static int resident_on[64] = {};
    const int slot = (h->device >= 0 && h->device < 64) ? h->device : 0;
    int resident = resident_on[slot];
    if (resident == 0 || slot != h->device) {
        int occ = 0;
        // (of the conversion-pipe instantiation also for MIX: the integer-
        // widening variant promises the bits of the same shape, and the grid
        // is part of the shape)
        ACCBLAS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &occ, dot_stream_kernel<St, Ar, BLOCK, UNROLL, CBY, false>, BLOCK,
            0));
        resident = occ > 0 ? occ : 1;
        if (slot == h->device) {
            resident_on[slot] = resident;
        }
    }
    // (several waves of shorter CTAs -- the reference's launch shape -- and
    // loads that allocate in L1 were measured too: no gain, profiles/r02_sweep.txt)
    if (ctas_per_sm <= 0 || ctas_per_sm > resident) {
        ctas_per_sm = resident;
    }
    std::int64_t tiles = n / TILE;
    std::int64_t grid = std::int64_t{h->sm_count} * ctas_per_sm;
    if (tiles < grid) {
        grid = tiles > 0 ? tiles : 1;
    }
    const std::int64_t max_grid = kScratchBytes / sizeof(Ar);
    if (grid > max_grid) {
        grid = max_grid;
    }
    int rc = ensure_workspace(h, kScratchBytes, stream);
    if (rc != ACCBLAS_OK) {
        return rc;
    }
    const int pdl = tuning().dot_pdl;
    // static share + dynamically assigned pool (see the kernel): the split is
    // a function of (n, grid) only, so repeated calls add in the same order
    std::int64_t static_tiles = tiles;
    int chunk_tiles = 0;
    std::int64_t num_chunks = 0;
    const int pool_pct = tuning().dot_pool_pct;
    if (!MIX && pool_pct > 0 && tiles >= 32 * grid) {
        std::int64_t pool = tiles * pool_pct / 100;
        static_tiles = (tiles - pool) / grid * grid;
        pool = tiles - static_tiles;
        // the chunk partials share the 64 KiB scratch with the CTA partials
        const std::int64_t room = max_grid - grid;
        const std::int64_t max_chunks = room < 4096 ? room : 4096;
        // at least ~8 chunks per CTA, so that the pool still balances when
        // it is small (n = 2^26: a 4-tile chunk is a tenth of a CTA's work)
        std::int64_t ct = tuning().dot_chunk_tiles;
        const std::int64_t fine = pool / (8 * grid);
        ct = ct < fine ? ct : (fine > 1 ? fine : 1);
        if (max_chunks > 0 && pool > 0) {
            const std::int64_t need = (pool + max_chunks - 1) / max_chunks;
            ct = ct > need ? ct : need;
            chunk_tiles = static_cast<int>(ct);
            num_chunks = (pool + ct - 1) / ct;
        } else {
            static_tiles = tiles;
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(BLOCK);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    ACCBLAS_CUDA(cudaLaunchKernelEx(
        &cfg, kernel, static_cast<const St*>(x), static_cast<const St*>(y), n,
        static_cast<Ar*>(payload(h)), control_words(h) + kCtlDotCounter, result,
        res, px, pdl, head, static_tiles, chunk_tiles,
        static_cast<unsigned>(num_chunks)));
    return ACCBLAS_OK;
}

// launch shape of the aligned kernel: CTA size and vectors in flight, from the
// tuning knobs (0 = the per-pair default measured on B200)
template <typename St, typename Ar, bool MIX, int CBY = 16>
int launch_shape(Handle* h, std::int64_t n, const void* x, const void* y,
                 void* result, int res, cudaStream_t stream,
                 const PeerExchange& px, int head)
{
    const Tuning& t = tuning();
    int unroll = t.dot_unroll;
    int block = t.dot_block;
    // same-box sweep on B200 (tools/sweep_dot_fill.py, profiles/r02_sweep.txt):
    // four vectors of each operand in flight win or tie for every pair
    // (fp16 storage: +5 % over two); 1024-thread CTAs are worth +3 % for
    // Acc<fp64,fp64> and cost the other pairs 1-7 %
    if (unroll == 0) {
        unroll = 4;
    }
    if (block == 0) {
        block = (sizeof(St) == 8 && sizeof(Ar) == 8) ? 1024 : 256;
    }
    const int cps = t.dot_ctas_per_sm;
#define ACCBLAS_DOT_SHAPE(B, U)                                               \
    if (block == B && unroll == U) {                                          \
        return launch_stream<St, Ar, B, U, CBY, MIX>(h, n, x, y, result, res, \
                                                     cps, stream, px, head);  \
    }
    ACCBLAS_DOT_SHAPE(256, 1)
    ACCBLAS_DOT_SHAPE(256, 2)
    ACCBLAS_DOT_SHAPE(256, 4)
    ACCBLAS_DOT_SHAPE(512, 1)
    ACCBLAS_DOT_SHAPE(1024, 1)
    ACCBLAS_DOT_SHAPE(512, 2)
    ACCBLAS_DOT_SHAPE(512, 4)
    ACCBLAS_DOT_SHAPE(1024, 2)
    ACCBLAS_DOT_SHAPE(1024, 4)
#undef ACCBLAS_DOT_SHAPE
    set_error("dot: no kernel for block=%d unroll=%d", block, unroll);
    return ACCBLAS_ERR_INVALID;
}

template <typename St, typename Ar>
int launch_dot(Handle* h, std::int64_t n, const void* x, std::int64_t incx,
               const void* y, std::int64_t incy, void* result, int res,
               cudaStream_t stream, const PeerExchange& px)
{
    if (incx == 1 && incy == 1) {
        std::uintptr_t xa = reinterpret_cast<std::uintptr_t>(x);
        std::uintptr_t ya = reinterpret_cast<std::uintptr_t>(y);
        if (xa % sizeof(St) == 0 && ya % sizeof(St) == 0) {
            const auto low = [](std::uintptr_t a) {
                return static_cast<unsigned>(a & 15u);
            };
            // peel so that x sits on a 16-byte boundary
            int head = 0;
            if (low(xa) != 0) {
                head = static_cast<int>((16u - low(xa)) / sizeof(St));
            }
            if (head <= n) {
                const St* xp = static_cast<const St*>(x) + head;
                const St* yp = static_cast<const St*>(y) + head;
                const unsigned dy =
                    low(reinterpret_cast<std::uintptr_t>(yp));
                const std::int64_t rest = n - head;
                if (dy == 0) {
                    if constexpr (std::is_same<St, float>::value &&
                                  std::is_same<Ar, double>::value) {
                        if (tuning().dot_intmix != 0) {
                            return launch_shape<St, Ar, true>(
                                h, rest, xp, yp, result, res, stream, px, head);
                        }
                    }
                    return launch_shape<St, Ar, false>(h, rest, xp, yp, result,
                                                       res, stream, px, head);
                }
                // different misalignments: y in pieces of its own alignment,
                // same launch shape (hence the same bits) as the aligned call
                if (dy % 8 == 0) {
                    return launch_shape<St, Ar, false, 8>(h, rest, xp, yp, result,
                                                          res, stream, px, head);
                }
                if constexpr (sizeof(St) <= 4) {
                    if (dy % 4 == 0) {
                        return launch_shape<St, Ar, false, 4>(
                            h, rest, xp, yp, result, res, stream, px, head);
                    }
                }
                if constexpr (sizeof(St) == 2) {
                    return launch_shape<St, Ar, false, 2>(h, rest, xp, yp, result,
                                                          res, stream, px, head);
                }
            }
        }
    }
    constexpr int BLOCK = 256;
    std::int64_t grid = std::int64_t{h->sm_count} * 8;
    const std::int64_t need = (n + BLOCK - 1) / BLOCK;
    if (need < grid) {
        grid = need > 0 ? need : 1;
    }
    const std::int64_t max_grid = kScratchBytes / sizeof(Ar);
    if (grid > max_grid) {
        grid = max_grid;
    }
    int rc = ensure_workspace(h, kScratchBytes, stream);
    if (rc != ACCBLAS_OK) {
        return rc;
    }
    dot_strided_kernel<St, Ar, BLOCK>
        <<<static_cast<unsigned>(grid), BLOCK, 0, stream>>>(
            static_cast<const St*>(x), incx, static_cast<const St*>(y), incy, n,
            static_cast<Ar*>(payload(h)), control_words(h) + kCtlDotCounter,
            result, res, px);
    ACCBLAS_CUDA(cudaGetLastError());
    return ACCBLAS_OK;
}

}  // namespace

int dot_impl(Handle* h, int ar, int st, int res, std::int64_t n, const void* x,
             std::int64_t incx, const void* y, std::int64_t incy, void* result,
             cudaStream_t stream, const PeerExchange* px_or_null)
{
    const PeerExchange px = px_or_null ? *px_or_null : PeerExchange{};
    return dispatch_ar_st(ar, st, [&](auto st_tag, auto ar_tag) {
        using St = decltype(st_tag);
        using Ar = decltype(ar_tag);
        return launch_dot<St, Ar>(h, n, x, incx, y, incy, result, res, stream,
                                  px);
    });
}

}  // namespace accblas
