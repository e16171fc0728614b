// Helpers shared by the two TRSV kernels (trsv.cu: one CTA per block row,
// hand-off through L2; trsv_cluster.cu: thread-block clusters, hand-off
// through distributed shared memory).
#pragma once

#include "common.cuh"

namespace accblas {
namespace trsv_detail {

constexpr int kB = 128;       // rows/cols per block row
constexpr int kSB = 32;       // diagonal sub-block
constexpr int kNSB = kB / kSB;
// Leading dimension of the smem tile: 136 = 8 * 17, so that 16-byte (two
// double) loads by 2 rows x 4 column slots per quarter-warp hit eight distinct
// 16-byte bank groups (row stride 136 doubles = 68 groups = 4 mod 8).
constexpr int kLD = kB + 8;
constexpr int kThreads = 512;
// x blocks of the single-CTA kernel: fetched kXAhead block iterations ahead of
// their use into a ring of kXAhead + 1 shared-memory buffers.  One block ahead
// is enough: three ahead measured the same (319.5 vs 317.5 us Acc<fp64,fp32>,
// fp32 arithmetic 2 % slower; session r02x) -- the copy's round trip is not
// what warp 0 waits for at the top of an iteration.
#ifndef ACCBLAS_TRSV_XAHEAD
#define ACCBLAS_TRSV_XAHEAD 1
#endif
constexpr int kXAhead = ACCBLAS_TRSV_XAHEAD;
constexpr int kXRing = kXAhead + 1;
constexpr int kWarps = kThreads / kWarp;
constexpr int kEPL = 4;                    // elements per lane per row
// Widen a panel BEFORE its x block is looked at (the conversions are then off
// the critical path of a CTA that has caught up with the chain), or let the
// compiler interleave the conversions with the FMAs.  Forcing the widened
// panel (64 registers of doubles) to stay live across the barrier costs more
// than it saves for fp32 storage with fp64 arithmetic: the CTAs are almost
// always BEHIND the chain (x already there), and the kernel sits at the
// 128-register limit.  Same box: 415.6 -> 328.7 us without it; the other
// pairs are within 1-3 % either way and keep it.
// Default of the L2 look-ahead (trsv_l2_ahead = -1), bytes per row.  Measured
// on B200 at n = 16384 (profiles/r02_trsv_l2_ahead_ab.txt): it pays for fp64
// arithmetic on fp32 / fp64 storage (the widest panels and the slowest block
// iterations), and costs 2-5 % for fp16 storage and for fp32 arithmetic,
// whose block iterations are short enough for the prefetch instructions
// themselves to show.
template <typename St, typename Ar>
constexpr int trsv_default_l2_ahead()
{
    return (sizeof(Ar) == 8 && sizeof(St) >= 4) ? 1024 : 0;
}

template <typename St, typename Ar>
struct pre_convert : std::true_type {};
template <>
struct pre_convert<float, double> : std::false_type {};

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

// Four consecutive storage elements as RAW 32-bit words.  The words are only
// taken apart by get<Ar>() in the conversion phase; pin() orders every use of
// the words after the loads issued so far.  (With a typed struct the compiler
// hoisted the half-word shuffles of fp16 storage to right behind each load and
// reused ONE destination register for all eight loads of a panel: eight
// serialised L2 round trips per block iteration -- ncu: 58 % of all stall
// samples on those shuffles.)
template <typename St>
struct Quad {
    static constexpr int kWords = kEPL * static_cast<int>(sizeof(St)) / 4;
    unsigned w[kWords];

    __device__ __forceinline__ void pin()
    {
#pragma unroll
        for (int i = 0; i < kWords; ++i) {
            asm volatile("" : "+r"(w[i]));
        }
    }
    __device__ __forceinline__ void set(int e, St value);
    template <typename Ar>
    __device__ __forceinline__ Ar get(int e) const;
};

template <>
__device__ __forceinline__ void Quad<double>::set(int e, double value)
{
    w[2 * e] = static_cast<unsigned>(__double2loint(value));
    w[2 * e + 1] = static_cast<unsigned>(__double2hiint(value));
}
template <>
__device__ __forceinline__ void Quad<float>::set(int e, float value)
{
    w[e] = __float_as_uint(value);
}
template <>
__device__ __forceinline__ void Quad<__half>::set(int e, __half value)
{
    const unsigned bits = __half_as_ushort(value);
    const int word = e >> 1;
    w[word] = (e & 1) ? ((w[word] & 0x0000ffffu) | (bits << 16))
                      : ((w[word] & 0xffff0000u) | bits);
}
template <>
template <typename Ar>
__device__ __forceinline__ Ar Quad<double>::get(int e) const
{
    return to_ar<Ar, double>(__hiloint2double(static_cast<int>(w[2 * e + 1]),
                                              static_cast<int>(w[2 * e])));
}
template <>
template <typename Ar>
__device__ __forceinline__ Ar Quad<float>::get(int e) const
{
    return to_ar<Ar, float>(__uint_as_float(w[e]));
}
template <>
template <typename Ar>
__device__ __forceinline__ Ar Quad<__half>::get(int e) const
{
    const unsigned short bits =
        static_cast<unsigned short>((e & 1) ? (w[e >> 1] >> 16)
                                            : (w[e >> 1] & 0xffffu));
    return to_ar<Ar, __half>(__ushort_as_half(bits));
}

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
// VW = vector width of the row loads in bytes: 16 (rows 16-byte aligned),
// 8 (8-byte aligned: odd lda of fp64, lda = 2 mod 4 of fp32; an fp16 quad is
// 8 bytes anyway) or 0 (scalar loads).  Compile time: a run-time flag in this
// loop cost fp64 storage 5 %.
template <typename St, int VW>
__device__ __forceinline__ Quad<St> load_quad(const St* p, int valid)
{
    constexpr bool VECTOR = VW != 0;
    Quad<St> q;
    if (VW == 8 && valid == kEPL && sizeof(St) > 2) {
#pragma unroll
        for (int i = 0; i < Quad<St>::kWords / 2; ++i) {
            const uint2 a = ldg_stream_64(reinterpret_cast<const char*>(p) + 8 * i);
            q.w[2 * i] = a.x;
            q.w[2 * i + 1] = a.y;
        }
    } else if (VECTOR && valid == kEPL) {
        if (sizeof(St) == 8) {
            const uint4 a = ldg_stream_128(p);
            const uint4 b = ldg_stream_128(p + 2);
            q.w[0] = a.x; q.w[1] = a.y; q.w[2] = a.z; q.w[3] = a.w;
            q.w[4 % Quad<St>::kWords] = b.x;
            q.w[5 % Quad<St>::kWords] = b.y;
            q.w[6 % Quad<St>::kWords] = b.z;
            q.w[7 % Quad<St>::kWords] = b.w;
        } else if (sizeof(St) == 4) {
            const uint4 a = ldg_stream_128(p);
            q.w[0] = a.x; q.w[1] = a.y;
            q.w[2 % Quad<St>::kWords] = a.z;
            q.w[3 % Quad<St>::kWords] = a.w;
        } else {
            const uint2 a = ldg_stream_64(p);
            q.w[0] = a.x; q.w[1] = a.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < Quad<St>::kWords; ++i) {
            q.w[i] = 0u;
        }
#pragma unroll
        for (int e = 0; e < kEPL; ++e) {
            if (e < valid) {
                q.set(e, p[e]);
            }
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


}  // namespace trsv_detail
}  // namespace accblas
