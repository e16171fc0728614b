// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// CPU baseline ("port") of the accessor-BLAS hot path: the bodies of
// kernel::acc_gemv / acc_dot / acc_{lower,upper}_trsv
// (/root/reference/cuda/gemv_kernels.cuh:79-113, cuda/dot_kernels.cuh:131-161,
// cuda/trsv_kernels.cuh:527-893) written as host loops over the same
// gko::acc::range<reduced_row_major<2, Ar, St>> objects, rows / index ranges
// split over the host cores with OpenMP.  The reference ships no CPU build of
// its kernels, so this is a reported baseline (bench.py cpu_baseline and
// --impl reference), never a target and never on the product path.
// Compiled with -O3 -mavx2 -mfma -ffp-contract=fast (see oracle/Makefile).
#include "oracle_common.hpp"

namespace {

using namespace oracle;

// ---------------------------------------------------------------------------
// CPU baseline ("port"): the accessor kernel bodies as host loops over the
// same range objects, rows / index ranges split over the host cores.
// ---------------------------------------------------------------------------
template <typename Ar, typename St>
int cpu_gemv(std::int64_t m, std::int64_t n, double alpha_d, const void* A_v,
             std::int64_t lda, const void* x_v, std::int64_t incx,
             double beta_d, void* y_v, std::int64_t incy)
{
    const Ar alpha = static_cast<Ar>(alpha_d), beta = static_cast<Ar>(beta_d);
    const_range<Ar, St> mtx(range_size{m, n}, static_cast<const St*>(A_v),
                            range_stride{lda});
    const_range<Ar, St> x(range_size{n, 1}, static_cast<const St*>(x_v),
                          range_stride{incx});
    mut_range<Ar, St> res(range_size{m, 1}, static_cast<St*>(y_v),
                          range_stride{incy});
#pragma omp parallel for schedule(static)
    for (std::int64_t row = 0; row < m; ++row) {
        Ar acc[4] = {Ar{}, Ar{}, Ar{}, Ar{}};
        std::int64_t col = 0;
        for (; col + 4 <= n; col += 4) {
            for (int k = 0; k < 4; ++k) {
                acc[k] += mtx(row, col + k) * x(col + k, 0);
            }
        }
        for (; col < n; ++col) {
            acc[0] += mtx(row, col) * x(col, 0);
        }
        const Ar s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        if (beta == Ar{0}) {
            res(row, 0) = alpha * s;
        } else {
            res(row, 0) = alpha * s + beta * res(row, 0);
        }
    }
    return 0;
}

template <typename Ar, typename St>
int cpu_dot(std::int64_t n, const void* x_v, std::int64_t incx,
            const void* y_v, std::int64_t incy, int res_t, void* result)
{
    const_range<Ar, St> x(range_size{n, 1}, static_cast<const St*>(x_v),
                          range_stride{incx});
    const_range<Ar, St> y(range_size{n, 1}, static_cast<const St*>(y_v),
                          range_stride{incy});
    Ar total{};
#pragma omp parallel
    {
        Ar acc[4] = {Ar{}, Ar{}, Ar{}, Ar{}};
#pragma omp for schedule(static) nowait
        for (std::int64_t i4 = 0; i4 < n / 4; ++i4) {
            for (int k = 0; k < 4; ++k) {
                acc[k] += x(4 * i4 + k, 0) * y(4 * i4 + k, 0);
            }
        }
        const Ar local = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma omp critical
        total += local;
    }
    for (std::int64_t i = n - n % 4; i < n; ++i) {
        total += x(i, 0) * y(i, 0);
    }
    switch (res_t) {
    case F64:
        *static_cast<double*>(result) = static_cast<double>(total);
        break;
    case F32:
        *static_cast<float*>(result) = static_cast<float>(total);
        break;
    default:
        *static_cast<f16*>(result) = narrow<f16, Ar>(total);
        break;
    }
    return 0;
}

// forward / backward substitution through the accessors, single thread (the
// dependency chain does not split over cores without blocking)
template <typename Ar, typename St>
int cpu_trsv(int upper, int unit, std::int64_t n, const void* A_v,
             std::int64_t lda, void* x_v, std::int64_t incx)
{
    const_range<Ar, St> mtx(range_size{n, n}, static_cast<const St*>(A_v),
                            range_stride{lda});
    mut_range<Ar, St> x(range_size{n, 1}, static_cast<St*>(x_v),
                        range_stride{incx});
    for (std::int64_t k = 0; k < n; ++k) {
        const std::int64_t r = upper ? n - 1 - k : k;
        Ar acc = x(r, 0);
        const std::int64_t c_begin = upper ? r + 1 : 0;
        const std::int64_t c_end = upper ? n : r;
        Ar part[4] = {Ar{}, Ar{}, Ar{}, Ar{}};
        std::int64_t c = c_begin;
        for (; c + 4 <= c_end; c += 4) {
            for (int q = 0; q < 4; ++q) {
                part[q] += mtx(r, c + q) * x(c + q, 0);
            }
        }
        for (; c < c_end; ++c) {
            part[0] += mtx(r, c) * x(c, 0);
        }
        acc -= (part[0] + part[1]) + (part[2] + part[3]);
        if (!unit) {
            acc /= mtx(r, r);
        }
        x(r, 0) = acc;
    }
    return 0;
}

}  // namespace

extern "C" {

// --- CPU baseline ----------------------------------------------------------------
int oracle_cpu_gemv(int ar, int st, std::int64_t m, std::int64_t n,
                    double alpha, const void* A, std::int64_t lda,
                    const void* x, std::int64_t incx, double beta, void* y,
                    std::int64_t incy)
{
    return dispatch_ar_st(ar, st, [&](auto a, auto s) {
        return cpu_gemv<decltype(a), decltype(s)>(m, n, alpha, A, lda, x, incx,
                                                  beta, y, incy);
    });
}

int oracle_cpu_dot(int ar, int st, int res_t, std::int64_t n, const void* x,
                   std::int64_t incx, const void* y, std::int64_t incy,
                   void* result)
{
    return dispatch_ar_st(ar, st, [&](auto a, auto s) {
        return cpu_dot<decltype(a), decltype(s)>(n, x, incx, y, incy, res_t,
                                                 result);
    });
}

int oracle_cpu_trsv(int ar, int st, int upper, int unit, std::int64_t n,
                    const void* A, std::int64_t lda, void* x,
                    std::int64_t incx)
{
    return dispatch_ar_st(ar, st, [&](auto a, auto s) {
        return cpu_trsv<decltype(a), decltype(s)>(upper, unit, n, A, lda, x,
                                                  incx);
    });
}

}  // extern "C"
