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
// The summation order depends only on (n, dtype pair, grid), never on which
// CTA happens to be last, so the result is bit-reproducible run to run.
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
// A peer that never shows up (2 s) yields NaN instead of a hung GPU.
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
            if (dot_globaltimer_ns() - t0 > 2000000000ull) {
                ok = false;
                break;
            }
        }
        __threadfence_system();  // flag before value
        peer_vals[t] = ok ? __longlong_as_double(static_cast<long long>(in[0]))
                          : __longlong_as_double(0x7ff8000000000000LL);
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

// Second pass + epilogue shared by both first-pass kernels.
template <typename Ar, int BLOCK>
__device__ __forceinline__ void finish_dot(Ar local, Ar* partials,
                                           unsigned* counter, void* result,
                                           int res_dtype, Ar* scratch,
                                           const PeerExchange& px)
{
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
    Ar v = Ar{};
    for (unsigned i = threadIdx.x; i < gridDim.x; i += BLOCK) {
        v += __ldcg(partials + i);
    }
    Ar sum = block_sum(v, scratch);
    if (px.world > 1) {
        sum = peer_allreduce(sum, px);
    }
    if (threadIdx.x == 0) {
        store_result(result, res_dtype, sum);
        *counter = 0u;  // re-arm for the next call on this handle
    }
}

// Contiguous, 16-byte aligned operands.
template <typename St, typename Ar, int BLOCK, int UNROLL>
__global__ __launch_bounds__(BLOCK) void dot_stream_kernel(
    const St* __restrict__ x, const St* __restrict__ y, std::int64_t n,
    Ar* __restrict__ partials, unsigned* __restrict__ counter,
    void* __restrict__ result, int res_dtype, const PeerExchange px, int pdl,
    int head)
{
    // x and y point at the first 16-byte aligned element; `head` (< 16 /
    // sizeof(St)) elements in front of it belong to the operands as well
    // (both vectors misaligned by the same amount, e.g. x[1:] . y[1:])
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

    for (std::int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const std::int64_t base = tile * TILE + std::int64_t{threadIdx.x} * VEC;
        uint4 xr[UNROLL];
        uint4 yr[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            xr[u] = ldg_stream_128(x + base + std::int64_t{u} * BLOCK * VEC);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            yr[u] = ldg_stream_128(y + base + std::int64_t{u} * BLOCK * VEC);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                acc[u][i % SLOTS] = pair_fma<Ar, St>::apply(
                    raw_elem<St>(xr[u], i), raw_elem<St>(yr[u], i),
                    acc[u][i % SLOTS]);
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
    Ar local = Ar{};
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
        local += v[0];
    }
    local += tail;
    finish_dot<Ar, BLOCK>(local, partials, counter, result, res_dtype, scratch,
                          px);
}

// Any stride / alignment: scalar loads, 64-bit indices (the reference's plain
// kernel is limited to int32, cuda/dot_kernels.cuh:89-97).
template <typename St, typename Ar, int BLOCK>
__global__ __launch_bounds__(BLOCK) void dot_strided_kernel(
    const St* __restrict__ x, std::int64_t incx, const St* __restrict__ y,
    std::int64_t incy, std::int64_t n, Ar* __restrict__ partials,
    unsigned* __restrict__ counter, void* __restrict__ result, int res_dtype,
    const PeerExchange px)
{
    __shared__ Ar scratch[kWarp];
    Ar acc0 = Ar{}, acc1 = Ar{};
    const std::int64_t step = std::int64_t{gridDim.x} * BLOCK;
    std::int64_t i = std::int64_t{blockIdx.x} * BLOCK + threadIdx.x;
    for (; i + step < n; i += 2 * step) {
        const St xa = x[i * incx], ya = y[i * incy];
        const St xb = x[(i + step) * incx], yb = y[(i + step) * incy];
        acc0 = pair_fma<Ar, St>::apply(xa, ya, acc0);
        acc1 = pair_fma<Ar, St>::apply(xb, yb, acc1);
    }
    if (i < n) {
        acc0 = pair_fma<Ar, St>::apply(x[i * incx], y[i * incy], acc0);
    }
    finish_dot<Ar, BLOCK>(acc0 + acc1, partials, counter, result, res_dtype,
                          scratch, px);
}

template <typename St, typename Ar, int BLOCK, int UNROLL>
int launch_stream(Handle* h, std::int64_t n, const void* x, const void* y,
                  void* result, int res, int ctas_per_sm, cudaStream_t stream,
                  const PeerExchange& px, int head)
{
    constexpr int VEC = vec_traits<St>::elems;
    constexpr std::int64_t TILE = std::int64_t{BLOCK} * VEC * UNROLL;
    auto kernel = dot_stream_kernel<St, Ar, BLOCK, UNROLL>;
    // one resident wave: every CTA of the grid-stride loop is on the machine
    // from start to end (a partial second wave would leave a tail)
    static int resident = 0;  // per instantiation
    if (resident == 0) {
        int occ = 0;
        ACCBLAS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &occ, kernel, BLOCK, 0));
        resident = occ > 0 ? occ : 1;
    }
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
        res, px, pdl, head));
    return ACCBLAS_OK;
}

template <typename St, typename Ar>
int launch_dot(Handle* h, std::int64_t n, const void* x, std::int64_t incx,
               const void* y, std::int64_t incy, void* result, int res,
               cudaStream_t stream, const PeerExchange& px)
{
    // contiguous operands with the SAME misalignment: peel the elements in
    // front of the first 16-byte boundary, stream the rest
    const std::uintptr_t xa = reinterpret_cast<std::uintptr_t>(x);
    const std::uintptr_t ya = reinterpret_cast<std::uintptr_t>(y);
    int head = 0;
    bool aligned = ((xa | ya) & 15u) == 0;
    if (!aligned && incx == 1 && incy == 1 && (xa & 15u) == (ya & 15u) &&
        (xa % sizeof(St)) == 0) {
        head = static_cast<int>((16u - (xa & 15u)) / sizeof(St));
        if (head <= n) {
            aligned = true;
            x = static_cast<const St*>(x) + head;
            y = static_cast<const St*>(y) + head;
            n -= head;
        } else {
            head = 0;
        }
    }
    if (incx == 1 && incy == 1 && aligned) {
        const int unroll = tuning().dot_unroll;
        const int cps = tuning().dot_ctas_per_sm;
        switch (unroll) {
        case 2:
            return launch_stream<St, Ar, 256, 2>(h, n, x, y, result, res, cps,
                                                 stream, px, head);
        case 8:
            return launch_stream<St, Ar, 256, 8>(h, n, x, y, result, res, cps,
                                                 stream, px, head);
        default:
            return launch_stream<St, Ar, 256, 4>(h, n, x, y, result, res, cps,
                                                 stream, px, head);
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
