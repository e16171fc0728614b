// Shared device/host helpers for the sm_100a accessor-BLAS kernels.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <type_traits>

#include "accblas.h"

namespace accblas {

constexpr int kWarp = 32;

// ---------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------
struct Handle {
    int device = 0;
    int sm_count = 0;
    // device workspace: [0, 256) bytes control words, then payload
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // staging buffers of the *_host entry points
    void* stage[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t stage_bytes[4] = {0, 0, 0, 0};
    // bytes of the TRSV progress region currently known to hold the sentinel
    size_t trsv_armed_bytes = 0;
    // peer exchange (multi-GPU DOT): this GPU's mailbox, the peers' mailboxes
    // as mapped into this process, and the call counter all ranks share
    void* mailbox = nullptr;
    void* peer_mailbox[8] = {};
    bool peer_is_ipc[8] = {};
    int peer_world = 0;
    int peer_rank = 0;
    unsigned long long peer_epoch = 0;
    // sticky failure word of the exchange: mapped pinned host memory the
    // kernel writes (system scope) when a peer did not arrive in time; the
    // host reads it at the start of every call without synchronising
    unsigned long long* peer_status_host = nullptr;
    unsigned long long* peer_status_dev = nullptr;
    unsigned long long peer_timeout_ns = 30000000000ull;  // 30 s
    // per-device launch constants of the DOT kernels (queried once)
    int dot_regs_checked = 0;
};

// What the last CTA of a DOT needs to combine the per-GPU partials itself:
// every rank writes {value, epoch} into slot [epoch & 1][rank] of EVERY rank's
// mailbox over NVLink and waits until its own mailbox holds all `world`
// entries of this epoch; the sum is formed in rank order on every GPU, so all
// ranks get the same bits whatever the arrival order.  world == 0: disabled.
constexpr int kMaxPeers = 8;
constexpr size_t kMailboxBytes = 2 * kMaxPeers * 16;
struct PeerExchange {
    int world = 0;
    int rank = 0;
    unsigned long long epoch = 0;
    unsigned long long timeout_ns = 0;       // 0 = wait for ever
    unsigned long long* status = nullptr;    // device alias of the failure word
    void* mailbox[kMaxPeers] = {};
};

// thread-local error message (accblas_last_error)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what);

#define ACCBLAS_CUDA(call)                                  \
    do {                                                    \
        cudaError_t err__ = (call);                         \
        if (err__ != cudaSuccess) {                         \
            return ::accblas::cuda_fail(err__, #call);      \
        }                                                   \
    } while (0)

// Makes sure the handle workspace holds at least `bytes` (payload, after the
// 256-byte control block).  Control block is zeroed on (re)allocation.
int ensure_workspace(Handle* h, size_t bytes, cudaStream_t stream);

constexpr size_t kControlBytes = 256;
// control words (unsigned, inside the control block)
constexpr int kCtlDotCounter = 0;    // DOT "blocks done" counter (self-resetting)
constexpr int kCtlTrsvTicket = 1;    // TRSV block-row ticket (self-resetting)
constexpr int kCtlFillFlag = 2;      // fill_uniform non-normal counter
constexpr int kCtlErrCounter = 3;    // l1_error "blocks done" counter
constexpr int kCtlDotChunk = 4;      // DOT: next chunk of the dynamically assigned pool (self-resetting)

inline unsigned* control_words(Handle* h)
{
    return reinterpret_cast<unsigned*>(h->ws);
}
// payload = [scratch for reduction partials | TRSV progress vector]
constexpr size_t kScratchBytes = size_t{64} << 10;
inline void* payload(Handle* h)
{
    return static_cast<char*>(h->ws) + kControlBytes;
}
inline void* trsv_region(Handle* h)
{
    return static_cast<char*>(h->ws) + kControlBytes + kScratchBytes;
}

// ---------------------------------------------------------------------------
// dtype dispatch
// ---------------------------------------------------------------------------
template <accblas_dtype T>
struct dtype_to_type;
template <>
struct dtype_to_type<ACCBLAS_F64> {
    using type = double;
};
template <>
struct dtype_to_type<ACCBLAS_F32> {
    using type = float;
};
template <>
struct dtype_to_type<ACCBLAS_F16> {
    using type = __half;
};

inline bool valid_dtype(int t) { return t >= 0 && t <= 2; }

// Calls f(St{}, Ar{}) for the (arithmetic, storage) pair; arithmetic must be
// fp64 or fp32.
template <typename F>
int dispatch_ar_st(int ar, int st, F&& f)
{
    if (!valid_dtype(ar) || !valid_dtype(st)) {
        set_error("invalid dtype (ar=%d, st=%d)", ar, st);
        return ACCBLAS_ERR_INVALID;
    }
    if (ar == ACCBLAS_F16) {
        set_error("fp16 arithmetic is not supported (storage only)");
        return ACCBLAS_ERR_UNSUPPORTED;
    }
    if (ar == ACCBLAS_F64) {
        switch (st) {
        case ACCBLAS_F64:
            return f(double{}, double{});
        case ACCBLAS_F32:
            return f(float{}, double{});
        default:
            return f(__half{}, double{});
        }
    } else {
        switch (st) {
        case ACCBLAS_F64:
            return f(double{}, float{});
        case ACCBLAS_F32:
            return f(float{}, float{});
        default:
            return f(__half{}, float{});
        }
    }
}

// ---------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------
#if defined(__CUDACC__)

// storage -> arithmetic (exact widening, or the narrowing static_cast does
// when St is wider than Ar)
template <typename Ar, typename St>
__device__ __forceinline__ Ar to_ar(St v)
{
    return static_cast<Ar>(v);
}
template <>
__device__ __forceinline__ float to_ar<float, __half>(__half v)
{
    return __half2float(v);
}
template <>
__device__ __forceinline__ double to_ar<double, __half>(__half v)
{
    // one F2F.F64.F16 (exact) instead of half -> float -> double
    double d;
    asm("cvt.f64.f16 %0, %1;" : "=d"(d) : "h"(__half_as_ushort(v)));
    return d;
}

// arithmetic -> storage: ONE round-to-nearest-even
template <typename St, typename Ar>
__device__ __forceinline__ St to_st(Ar v)
{
    return static_cast<St>(v);
}
template <>
__device__ __forceinline__ __half to_st<__half, float>(float v)
{
    return __float2half_rn(v);
}
template <>
__device__ __forceinline__ __half to_st<__half, double>(double v)
{
    return __double2half(v);
}
template <>
__device__ __forceinline__ __half to_st<__half, __half>(__half v)
{
    return v;
}

// fused multiply-add in the arithmetic type (what nvcc contracts
// `acc += a * b` to in the reference kernels)
__device__ __forceinline__ double fma_ar(double a, double b, double c)
{
    return fma(a, b, c);
}
__device__ __forceinline__ float fma_ar(float a, float b, float c)
{
    return fmaf(a, b, c);
}

// 128-bit streaming load: read-only path, do not allocate in L1 (the matrix /
// vector stream is touched exactly once)
__device__ __forceinline__ uint4 ldg_stream_128(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// 128-bit cached load (x vector in GEMV: re-read by every warp of the SM)
__device__ __forceinline__ uint4 ldg_cached_128(const void* p)
{
    return __ldg(reinterpret_cast<const uint4*>(p));
}

// same, as a volatile asm so that it keeps its place among the streaming loads
__device__ __forceinline__ uint4 ldg_cached_128_ordered(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// 16 bytes as 64- or 32-bit loads (volatile asm: keeps its place among the
// other streaming loads); STREAM = do not allocate in L1
template <int CB, bool STREAM>
__device__ __forceinline__ uint4 ldg_pieces(const void* p)
{
    const char* c = static_cast<const char*>(p);
    uint4 r;
    if constexpr (CB == 16) {
        return STREAM ? ldg_stream_128(p) : ldg_cached_128_ordered(p);
    } else if constexpr (CB == 8) {
        if constexpr (STREAM) {
            asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];"
                         : "=r"(r.x), "=r"(r.y) : "l"(c));
            asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];"
                         : "=r"(r.z), "=r"(r.w) : "l"(c + 8));
        } else {
            asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];"
                         : "=r"(r.x), "=r"(r.y) : "l"(c));
            asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];"
                         : "=r"(r.z), "=r"(r.w) : "l"(c + 8));
        }
    } else if constexpr (CB == 4) {
        unsigned w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if constexpr (STREAM) {
                asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];"
                             : "=r"(w[i]) : "l"(c + 4 * i));
            } else {
                asm volatile("ld.global.nc.u32 %0, [%1];"
                             : "=r"(w[i]) : "l"(c + 4 * i));
            }
        }
        r = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        // 2-byte aligned (fp16 rows with an odd stride or an odd base).  The
        // 16 bytes start either on a 4-byte boundary -- four 32-bit loads --
        // or two bytes past one: a 16-bit load, three 32-bit loads, a 16-bit
        // load (every piece naturally aligned and inside the 16 bytes), put
        // back together with funnel shifts.  The case is the same for all
        // lanes of a warp (they read one row, 16 bytes apart).  Five loads at
        // most instead of the eight 16-bit loads this path started with.
        auto ld32 = [](const char* q) {
            unsigned v;
            if constexpr (STREAM) {
                asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];"
                             : "=r"(v) : "l"(q));
            } else {
                asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(q));
            }
            return v;
        };
        auto ld16 = [](const char* q) {
            unsigned short v;
            if constexpr (STREAM) {
                asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];"
                             : "=h"(v) : "l"(q));
            } else {
                asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(q));
            }
            return static_cast<unsigned>(v);
        };
        if ((reinterpret_cast<std::uintptr_t>(c) & 2u) == 0) {
            r = make_uint4(ld32(c), ld32(c + 4), ld32(c + 8), ld32(c + 12));
        } else {
            const unsigned h0 = ld16(c);
            const unsigned m0 = ld32(c + 2);
            const unsigned m1 = ld32(c + 6);
            const unsigned m2 = ld32(c + 10);
            const unsigned h7 = ld16(c + 14);
            r = make_uint4(h0 | (m0 << 16), __funnelshift_r(m0, m1, 16),
                           __funnelshift_r(m1, m2, 16), (m2 >> 16) | (h7 << 16));
        }
    }
    return r;
}

// number of St elements in one 128-bit vector
template <typename St>
struct vec_traits {
    static constexpr int elems = 16 / sizeof(St);
};

// unpack element `i` (compile-time after unrolling) of a raw 128-bit vector
template <typename Ar>
__device__ __forceinline__ Ar unpack(const uint4& raw, int i, double)
{
    const unsigned lo = (i == 0) ? raw.x : raw.z;
    const unsigned hi = (i == 0) ? raw.y : raw.w;
    return to_ar<Ar, double>(__hiloint2double(hi, lo));
}
template <typename Ar>
__device__ __forceinline__ Ar unpack(const uint4& raw, int i, float)
{
    const unsigned w = (i == 0) ? raw.x : (i == 1) ? raw.y : (i == 2) ? raw.z : raw.w;
    return to_ar<Ar, float>(__uint_as_float(w));
}
template <typename Ar>
__device__ __forceinline__ Ar unpack(const uint4& raw, int i, __half)
{
    const int word = i >> 1;
    const unsigned w = (word == 0) ? raw.x : (word == 1) ? raw.y : (word == 2) ? raw.z : raw.w;
    const unsigned short bits =
        static_cast<unsigned short>((i & 1) ? (w >> 16) : (w & 0xffffu));
    return to_ar<Ar, __half>(__ushort_as_half(bits));
}

// Converts a raw 128-bit vector of St into vec_traits<St>::elems values of Ar.
template <typename Ar, typename St>
__device__ __forceinline__ void unpack_all(const uint4& raw, Ar* out)
{
#pragma unroll
    for (int i = 0; i < vec_traits<St>::elems; ++i) {
        out[i] = unpack<Ar>(raw, i, St{});
    }
}

// warp butterfly sum: fixed order, every lane ends with the same value
template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int mask = kWarp / 2; mask > 0; mask >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, mask);
    }
    return v;
}

// CTA-wide sum (fixed order): warp butterflies, one smem slot per warp, then
// warp 0 folds the slots with a second butterfly.  Result valid in thread 0.
// `scratch` must hold 32 values of T.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch)
{
    const int lane = threadIdx.x & (kWarp - 1);
    const int warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + kWarp - 1) >> 5;
    v = warp_sum(v);
    __syncthreads();  // scratch may still be in use by a previous call
    if (lane == 0) {
        scratch[warp] = v;
    }
    __syncthreads();
    T total = T{};
    if (warp == 0) {
        total = (lane < nwarps) ? scratch[lane] : T{};
        total = warp_sum(total);
    }
    return total;
}

#endif  // __CUDACC__

}  // namespace accblas
