// Clean-room, API-compatible stand-in for the part of Ginkgo's header-only
// `accessor/` directory that accessor-BLAS uses.  The upstream headers are NOT
// vendored by the reference (it clones ginkgo `develop` at configure time,
// /root/reference/CMakeLists.txt:18-33) and are not available offline, so this
// file was written from the *call sites* in the reference
// (cuda/gemv_kernels.cuh:178-189, cuda/dot_kernels.cuh:234-243,
// cuda/trsv_kernels.cuh:924-933) and the C++ semantics they rely on:
//   read  = static_cast<arithmetic_type>(storage value)
//   write = static_cast<storage_type>(arithmetic value)
//   index = row * stride + col
//
// Extension over upstream: `__half` storage (single round-to-nearest-even from
// the arithmetic type, exact widening on read).
#pragma once

#include <cstdint>
#include <type_traits>

#if defined(__CUDACC__)
#include <cuda_fp16.h>
#define GKO_ACC_ATTRIBUTES __host__ __device__
#define GKO_ACC_INLINE __forceinline__
#define GKO_ACC_HAS_HALF 1
#else
#define GKO_ACC_ATTRIBUTES
#define GKO_ACC_INLINE inline
#endif

namespace gko {
namespace acc {

// Upstream uses a signed 64-bit size type; the reference relies on that when it
// passes `matrix_info::size` (std::array<std::int64_t, 2>) straight into the
// range constructor (cuda/gemv_kernels.cuh:187 with cuda/utils.cuh:19-21).
using size_type = std::int64_t;

namespace detail {

// Storage <-> arithmetic conversion with the exact semantics of static_cast
// for float/double and of the cuda_fp16.h conversion functions for __half.
template <typename To, typename From>
struct storage_cast {
    static constexpr GKO_ACC_ATTRIBUTES GKO_ACC_INLINE To apply(From v)
    {
        return static_cast<To>(v);
    }
};

#if defined(GKO_ACC_HAS_HALF)
// half -> float is exact, float -> double is exact.
template <>
struct storage_cast<float, __half> {
    static GKO_ACC_ATTRIBUTES GKO_ACC_INLINE float apply(__half v)
    {
        return __half2float(v);
    }
};
template <>
struct storage_cast<double, __half> {
    static GKO_ACC_ATTRIBUTES GKO_ACC_INLINE double apply(__half v)
    {
        return static_cast<double>(__half2float(v));
    }
};
// ONE rounding step (round-to-nearest-even) from the arithmetic type.
template <>
struct storage_cast<__half, float> {
    static GKO_ACC_ATTRIBUTES GKO_ACC_INLINE __half apply(float v)
    {
        return __float2half_rn(v);
    }
};
template <>
struct storage_cast<__half, double> {
    static GKO_ACC_ATTRIBUTES GKO_ACC_INLINE __half apply(double v)
    {
        return __double2half(v);
    }
};
template <>
struct storage_cast<__half, __half> {
    static GKO_ACC_ATTRIBUTES GKO_ACC_INLINE __half apply(__half v)
    {
        return v;
    }
};
#endif

}  // namespace detail
}  // namespace acc
}  // namespace gko
