// TEST INFRASTRUCTURE ONLY -- shared pieces of the CPU oracle (see oracle.cpp).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

#if defined(_OPENMP)
#include <omp.h>
#endif

#include <accessor/range.hpp>
#include <accessor/reduced_row_major.hpp>

namespace oracle {

// ---------------------------------------------------------------------------
// IEEE binary16 in software (storage only)
// ---------------------------------------------------------------------------
inline std::uint16_t f64_to_f16_bits(double d)
{
    // ONE round-to-nearest-even from the exact fp64 value, the semantics of
    // static_cast<__half>(double) == __double2half (cuda_fp16.h).
    std::uint64_t b;
    std::memcpy(&b, &d, sizeof(b));
    const std::uint16_t sign = static_cast<std::uint16_t>((b >> 48) & 0x8000u);
    const int e = static_cast<int>((b >> 52) & 0x7ff);
    const std::uint64_t m = b & 0xfffffffffffffULL;
    if (e == 0x7ff) {
        return m ? std::uint16_t{0x7fff}
                 : static_cast<std::uint16_t>(sign | 0x7c00u);
    }
    if (e == 0) {
        return sign;  // zero or fp64 subnormal: far below half's 2^-25
    }
    const int exp = e - 1023;
    const std::uint64_t sig = (std::uint64_t{1} << 52) | m;
    int shift = 42;
    int field = exp + 15 - 1;  // the implicit bit carries into the field
    if (exp < -14) {
        shift += -14 - exp;
        field = 0;
        if (shift > 54) {
            return sign;
        }
    } else if (exp > 15) {
        return static_cast<std::uint16_t>(sign | 0x7c00u);
    }
    std::uint64_t q = sig >> shift;
    const std::uint64_t rem = sig & ((std::uint64_t{1} << shift) - 1);
    const std::uint64_t half = std::uint64_t{1} << (shift - 1);
    if (rem > half || (rem == half && (q & 1))) {
        ++q;
    }
    std::uint64_t bits = (static_cast<std::uint64_t>(field) << 10) + q;
    if (exp < -14) {
        bits = q;  // subnormal (or the smallest normal after a carry)
    }
    if (bits >= 0x7c00u) {
        bits = 0x7c00u;
    }
    return static_cast<std::uint16_t>(sign | bits);
}

inline float f16_bits_to_f32(std::uint16_t h)
{
    const std::uint32_t sign = static_cast<std::uint32_t>(h & 0x8000u) << 16;
    const int e = (h >> 10) & 0x1f;
    const std::uint32_t m = h & 0x3ffu;
    float out;
    if (e == 0) {
        // subnormal: m * 2^-24 (exact in fp32)
        out = std::ldexp(static_cast<float>(m), -24);
        std::uint32_t b;
        std::memcpy(&b, &out, 4);
        b |= sign;
        std::memcpy(&out, &b, 4);
        return out;
    }
    std::uint32_t b;
    if (e == 31) {
        b = sign | 0x7f800000u | (m << 13);
    } else {
        b = sign | (static_cast<std::uint32_t>(e + 112) << 23) | (m << 13);
    }
    std::memcpy(&out, &b, 4);
    return out;
}

struct f16 {
    std::uint16_t bits;
    f16() : bits(0) {}
    explicit f16(double v) : bits(f64_to_f16_bits(v)) {}
    // float -> double is exact, so one rounding happens in total
    explicit f16(float v) : bits(f64_to_f16_bits(static_cast<double>(v))) {}
    explicit operator float() const { return f16_bits_to_f32(bits); }
    explicit operator double() const
    {
        return static_cast<double>(f16_bits_to_f32(bits));
    }
};

}  // namespace oracle

// the compat accessor converts through storage_cast; teach it the host half
namespace gko {
namespace acc {
namespace detail {
using oracle::f16;
template <>
struct storage_cast<float, f16> {
    static float apply(f16 v) { return static_cast<float>(v); }
};
template <>
struct storage_cast<double, f16> {
    static double apply(f16 v) { return static_cast<double>(v); }
};
template <>
struct storage_cast<f16, float> {
    static f16 apply(float v) { return f16(v); }
};
template <>
struct storage_cast<f16, double> {
    static f16 apply(double v) { return f16(v); }
};
}  // namespace detail
}  // namespace acc
}  // namespace gko

namespace oracle {

enum { F64 = 0, F32 = 1, F16 = 2 };

template <typename Ar, typename St>
Ar widen(St v)
{
    return gko::acc::detail::storage_cast<Ar, St>::apply(v);
}
template <typename St, typename Ar>
St narrow(Ar v)
{
    return gko::acc::detail::storage_cast<St, Ar>::apply(v);
}

template <typename F>
int dispatch_st(int st, F&& f)
{
    switch (st) {
    case F64:
        return f(double{});
    case F32:
        return f(float{});
    case F16:
        return f(f16{});
    default:
        return 1;
    }
}

template <typename F>
int dispatch_ar_st(int ar, int st, F&& f)
{
    if (ar == F64) {
        return dispatch_st(st, [&](auto s) { return f(double{}, s); });
    }
    if (ar == F32) {
        return dispatch_st(st, [&](auto s) { return f(float{}, s); });
    }
    return 1;
}

using range_size = std::array<gko::acc::size_type, 2>;
using range_stride = std::array<gko::acc::size_type, 1>;

template <typename Ar, typename St>
using const_range = gko::acc::range<
    typename gko::acc::reduced_row_major<2, Ar, St>::const_accessor>;
template <typename Ar, typename St>
using mut_range = gko::acc::range<gko::acc::reduced_row_major<2, Ar, St>>;


inline double fma_t(double a, double b, double c) { return std::fma(a, b, c); }
inline float fma_t(float a, float b, float c) { return std::fmaf(a, b, c); }

}  // namespace oracle
