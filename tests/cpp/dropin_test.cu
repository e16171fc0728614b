// Compiles against include/accblas/*.cuh exactly the way a user of the
// reference compiles against its cuda/*.cuh, and exercises
//   (1) the host launchers with the reference's signatures,
//   (2) the kernel-level templates with the documented <<<>>> launch shapes,
// comparing everything with straightforward host loops.  Prints PASS and
// returns 0, or prints the failing check and returns 1.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

#include <accessor/block_col_major.hpp>
#include <accessor/row_major.hpp>
#include <accessor/scaled_reduced_row_major.hpp>

#include <accblas/dot_kernels.cuh>
#include <accblas/gemv_kernels.cuh>
#include <accblas/trsv_kernels.cuh>

namespace {

int failures = 0;

void check(bool ok, const char* what, double got, double want)
{
    if (!ok) {
        std::printf("FAIL %s: got %.10e want %.10e\n", what, got, want);
        ++failures;
    }
}

template <typename T>
T* to_device(const std::vector<T>& v)
{
    T* p = nullptr;
    CUDA_CALL(cudaMalloc(&p, sizeof(T) * v.size()));
    CUDA_CALL(cudaMemcpy(p, v.data(), sizeof(T) * v.size(),
                         cudaMemcpyHostToDevice));
    return p;
}

template <typename T>
std::vector<T> to_host(const T* p, std::size_t n)
{
    std::vector<T> v(n);
    CUDA_CALL(cudaMemcpy(v.data(), p, sizeof(T) * n, cudaMemcpyDeviceToHost));
    return v;
}

double l1_rel(const std::vector<double>& ref, const std::vector<double>& got)
{
    double d = 0, s = 0;
    for (std::size_t i = 0; i < ref.size(); ++i) {
        d += std::abs(ref[i] - got[i]);
        s += std::abs(ref[i]);
    }
    return d / s;
}

template <typename T>
std::vector<double> widen(const std::vector<T>& v)
{
    std::vector<double> out(v.size());
    for (std::size_t i = 0; i < v.size(); ++i) {
        out[i] = static_cast<double>(v[i]);
    }
    return out;
}

}  // namespace

int main()
{
    using size_type = matrix_info::size_type;
    std::default_random_engine engine(42);
    std::uniform_real_distribution<double> dist(-1.0, 1.0);

    // ------------------------------------------------------------------ GEMV
    const size_type m = 301, n = 517, stride = 520;
    std::vector<double> A(m * stride), x(n), y(m);
    for (auto& v : A) v = dist(engine);
    for (auto& v : x) v = dist(engine);
    for (auto& v : y) v = dist(engine);
    std::vector<float> Af(A.begin(), A.end()), xf(x.begin(), x.end()),
        yf(y.begin(), y.end());
    std::vector<double> want(m), want_f(m);
    for (size_type r = 0; r < m; ++r) {
        double s = 0, sf = 0;
        for (size_type c = 0; c < n; ++c) {
            s += A[r * stride + c] * x[c];
            sf += double(Af[r * stride + c]) * double(xf[c]);
        }
        want[r] = 0.5 * s + 2.0 * y[r];
        want_f[r] = 0.5 * sf + 2.0 * double(yf[r]);
    }
    const matrix_info m_info{{m, n}, stride};
    const matrix_info x_info{{n, 1}};
    const matrix_info res_info{{m, 1}};
    {
        double* dA = to_device(A);
        double* dx = to_device(x);
        double* dy = to_device(y);
        gemv(m_info, 0.5, dA, x_info, dx, res_info, 2.0, dy);
        synchronize();
        const double e = l1_rel(want, to_host(dy, m));
        check(e < 1e-14, "gemv<double>", e, 0);
        float* fA = to_device(Af);
        float* fx = to_device(xf);
        float* fy = to_device(yf);
        acc_gemv<double>(m_info, 0.5, fA, x_info, fx, res_info, 2.0, fy);
        synchronize();
        const double e2 = l1_rel(want_f, widen(to_host(fy, m)));
        check(e2 < 6e-8, "acc_gemv<double,float>", e2, 0);

        // kernel-level launch exactly as in the reference's launcher
        CUDA_CALL(cudaMemcpy(fy, yf.data(), sizeof(float) * m,
                             cudaMemcpyHostToDevice));
        using accessor = gko::acc::reduced_row_major<2, double, float>;
        using range = gko::acc::range<accessor>;
        using c_range = gko::acc::range<typename accessor::const_accessor>;
        std::array<gko::acc::size_type, 1> ms{stride}, xs{1}, rs{1};
        auto m_acc = c_range(m_info.size, fA, ms);
        auto x_acc = c_range(x_info.size, fx, xs);
        auto res_acc = range(res_info.size, fy, rs);
        kernel::acc_gemv<512><<<static_cast<unsigned>(m), 512>>>(
            0.5, m_acc, x_acc, 2.0, res_acc);
        synchronize();
        const double e3 = l1_rel(want_f, widen(to_host(fy, m)));
        check(e3 < 6e-8, "kernel::acc_gemv<512>", e3, 0);
        // range overload of the launcher
        CUDA_CALL(cudaMemcpy(fy, yf.data(), sizeof(float) * m,
                             cudaMemcpyHostToDevice));
        acc_gemv<double, float>(0.5, m_acc, x_acc, 2.0, res_acc);
        synchronize();
        const double e4 = l1_rel(want_f, widen(to_host(fy, m)));
        check(e4 < 6e-8, "acc_gemv(range...)", e4, 0);
        // "all other accessors work accordingly" (README.md:19): the same
        // kernel template over the plain row_major accessor (fp64 in, fp64 out)
        {
            CUDA_CALL(cudaMemcpy(dy, y.data(), sizeof(double) * m,
                                 cudaMemcpyHostToDevice));
            using rm = gko::acc::row_major<double, 2>;
            using rm_range = gko::acc::range<rm>;
            using rm_crange = gko::acc::range<typename rm::const_accessor>;
            auto m_rm = rm_crange(m_info.size, static_cast<const double*>(dA), ms);
            auto x_rm = rm_crange(x_info.size, static_cast<const double*>(dx), xs);
            auto res_rm = rm_range(res_info.size, dy, rs);
            kernel::acc_gemv<512><<<static_cast<unsigned>(m), 512>>>(
                0.5, m_rm, x_rm, 2.0, res_rm);
            synchronize();
            const double e6 = l1_rel(want, to_host(dy, m));
            check(e6 < 1e-14, "kernel::acc_gemv<512> over row_major", e6, 0);
            // the range launcher routes row_major to the C-ABI kernel
            // (arithmetic == storage), bit-identical to gemv<double>
            CUDA_CALL(cudaMemcpy(dy, y.data(), sizeof(double) * m,
                                 cudaMemcpyHostToDevice));
            gemv(m_info, 0.5, dA, x_info, dx, res_info, 2.0, dy);
            synchronize();
            const std::vector<double> fast = to_host(dy, m);
            CUDA_CALL(cudaMemcpy(dy, y.data(), sizeof(double) * m,
                                 cudaMemcpyHostToDevice));
            acc_gemv(0.5, m_rm, x_rm, 2.0, res_rm);
            synchronize();
            const std::vector<double> routed = to_host(dy, m);
            bool same = true;
            for (size_type r = 0; r < m; ++r) {
                same = same && routed[r] == fast[r];
            }
            check(same, "acc_gemv(row_major ranges) == gemv<double> bitwise", 0, 0);
        }
        // scaled_reduced_row_major: int16 storage with one fp64 scale per ROW,
        // through the generic range launcher
        {
            std::vector<std::int16_t> Aq(m * stride);
            std::vector<double> scale(m), want_q(m);
            for (size_type r = 0; r < m; ++r) {
                double big = 0;
                for (size_type c = 0; c < n; ++c) {
                    big = std::max(big, std::abs(A[r * stride + c]));
                }
                scale[r] = big / 32767.0;
                double acc = 0;
                for (size_type c = 0; c < n; ++c) {
                    Aq[r * stride + c] = static_cast<std::int16_t>(
                        std::lrint(A[r * stride + c] / scale[r]));
                    acc += double(Aq[r * stride + c]) * scale[r] * x[c];
                }
                want_q[r] = 0.5 * acc + 2.0 * y[r];
            }
            std::int16_t* dq = to_device(Aq);
            double* ds = to_device(scale);
            CUDA_CALL(cudaMemcpy(dy, y.data(), sizeof(double) * m,
                                 cudaMemcpyHostToDevice));
            using sacc = gko::acc::scaled_reduced_row_major<2, double,
                                                            const std::int16_t, 0b10>;
            using rm = gko::acc::row_major<double, 2>;
            auto m_s = gko::acc::range<sacc>(
                m_info.size, static_cast<const std::int16_t*>(dq), ms,
                static_cast<const double*>(ds));
            auto x_p = gko::acc::range<typename rm::const_accessor>(
                x_info.size, static_cast<const double*>(dx), xs);
            auto r_p = gko::acc::range<rm>(res_info.size, dy, rs);
            acc_gemv(0.5, m_s, x_p, 2.0, r_p);
            synchronize();
            const double e7 = l1_rel(want_q, to_host(dy, m));
            check(e7 < 1e-14, "acc_gemv over scaled_reduced_row_major", e7, 0);
            // writing through the proxy divides by the scale and rounds once
            std::vector<double> probe{0.75 * scale[3] * 1000.0};
            {
                using wacc = gko::acc::scaled_reduced_row_major<2, double,
                                                                std::int16_t, 0b10>;
                std::vector<std::int16_t> hq(Aq);
                std::vector<double> hs(scale);
                auto w = gko::acc::range<wacc>(m_info.size, hq.data(), ms, hs.data());
                w(3, 5) = probe[0];
                check(hq[3 * stride + 5] == 750, "scaled proxy write", hq[3 * stride + 5], 750);
                check(std::abs(double(w(3, 5)) - probe[0]) < 1e-12 * std::abs(probe[0]),
                      "scaled proxy read back", double(w(3, 5)), probe[0]);
            }
            CUDA_CALL(cudaFree(dq));
            CUDA_CALL(cudaFree(ds));
        }
        // block_col_major: the same matrix stored column-major
        {
            std::vector<double> At(n * m);
            for (size_type r = 0; r < m; ++r) {
                for (size_type c = 0; c < n; ++c) {
                    At[c * m + r] = A[r * stride + c];
                }
            }
            double* dt = to_device(At);
            CUDA_CALL(cudaMemcpy(dy, y.data(), sizeof(double) * m,
                                 cudaMemcpyHostToDevice));
            using cm = gko::acc::block_col_major<const double, 2>;
            using rm = gko::acc::row_major<double, 2>;
            std::array<gko::acc::size_type, 1> cs{m};
            auto m_c = gko::acc::range<cm>(m_info.size, static_cast<const double*>(dt), cs);
            auto x_p = gko::acc::range<typename rm::const_accessor>(
                x_info.size, static_cast<const double*>(dx), xs);
            auto r_p = gko::acc::range<rm>(res_info.size, dy, rs);
            acc_gemv(0.5, m_c, x_p, 2.0, r_p);
            synchronize();
            const double e8 = l1_rel(want, to_host(dy, m));
            check(e8 < 1e-14, "acc_gemv over block_col_major", e8, 0);
            CUDA_CALL(cudaFree(dt));
        }
        // plain kernel template
        CUDA_CALL(cudaMemcpy(dy, y.data(), sizeof(double) * m,
                             cudaMemcpyHostToDevice));
        kernel::gemv<512, double><<<static_cast<unsigned>(m), 512>>>(
            m_info, 0.5, dA, x_info, dx, res_info, 2.0, dy);
        synchronize();
        const double e5 = l1_rel(want, to_host(dy, m));
        check(e5 < 1e-14, "kernel::gemv<512,double>", e5, 0);
    }

    // ------------------------------------------------------------------- DOT
    {
        const size_type len = 1'000'003;
        std::vector<double> a(len), b(len);
        for (auto& v : a) v = dist(engine);
        for (auto& v : b) v = dist(engine);
        std::vector<float> af(a.begin(), a.end()), bf(b.begin(), b.end());
        long double s = 0, sf = 0;
        for (size_type i = 0; i < len; ++i) {
            s += static_cast<long double>(a[i]) * b[i];
            sf += static_cast<long double>(af[i]) * bf[i];
        }
        myBlasHandle handle;
        const matrix_info v_info{{len, 1}};
        double* da = to_device(a);
        double* db = to_device(b);
        float* fa = to_device(af);
        float* fb = to_device(bf);
        double* dres = to_device(std::vector<double>{-999.0});
        float* fres = to_device(std::vector<float>{-999.0f});
        dot(&handle, v_info, da, v_info, db, dres);
        synchronize();
        double got = to_host(dres, 1)[0];
        check(std::abs(got - double(s)) < 1e-10, "dot<double>", got, double(s));
        acc_dot<double>(&handle, v_info, fa, v_info, fb, fres);
        synchronize();
        got = to_host(fres, 1)[0];
        check(std::abs(got - double(sf)) <= std::abs(double(sf)) * 1.2e-7,
              "acc_dot<double,float,float>", got, double(sf));
        // kernel-level: accumulate into an initialised scalar
        using accessor = gko::acc::reduced_row_major<2, double, float>;
        using c_range = gko::acc::range<typename accessor::const_accessor>;
        std::array<gko::acc::size_type, 1> one{1};
        auto xa = c_range(v_info.size, fa, one);
        auto ya = c_range(v_info.size, fb, one);
        kernel::init_res<<<1, 1>>>(dres);
        kernel::acc_dot<1024><<<148 * 4, 1024>>>(xa, ya, dres);
        synchronize();
        got = to_host(dres, 1)[0];
        check(std::abs(got - double(sf)) < 1e-9, "kernel::acc_dot<1024>", got,
              double(sf));
    }

    // ------------------------------------------------------------------ TRSV
    for (int variant = 0; variant < 4; ++variant) {
        const bool upper = variant & 1, unit = variant & 2;
        const size_type nt = 333;
        std::vector<double> T(nt * nt), b(nt);
        for (auto& v : T) v = 0.02 * dist(engine);
        for (size_type i = 0; i < nt; ++i) {
            T[i * nt + i] = 1.0 + 0.5 * dist(engine);
        }
        for (auto& v : b) v = dist(engine);
        std::vector<float> Tf(T.begin(), T.end()), bf(b.begin(), b.end());
        std::vector<double> sol(nt);
        for (size_type k = 0; k < nt; ++k) {
            const size_type r = upper ? nt - 1 - k : k;
            long double acc = bf[r];
            for (size_type c = upper ? r + 1 : 0; c < (upper ? nt : r); ++c) {
                acc -= static_cast<long double>(Tf[r * nt + c]) * sol[c];
            }
            sol[r] = unit ? double(acc) : double(acc / Tf[r * nt + r]);
        }
        const matrix_info t_info{{nt, nt}};
        const matrix_info b_info{{nt, 1}};
        const tmtx_t tt = upper ? tmtx_t::upper : tmtx_t::lower;
        const dmtx_t dt = unit ? dmtx_t::unit : dmtx_t::non_unit;
        float* dT = to_device(Tf);
        float* dx = to_device(bf);
        std::uint32_t* helper = to_device(std::vector<std::uint32_t>{0, 0});
        acc_trsv<double>(t_info, tt, dt, dT, b_info, dx, helper);
        synchronize();
        double e = l1_rel(sol, widen(to_host(dx, nt)));
        check(e < 2e-7, "acc_trsv<double,float>", e, variant);

        // kernel-level, launched like the reference's launcher does
        CUDA_CALL(cudaMemcpy(dx, bf.data(), sizeof(float) * nt,
                             cudaMemcpyHostToDevice));
        using accessor = gko::acc::reduced_row_major<2, double, float>;
        using range = gko::acc::range<accessor>;
        using c_range = gko::acc::range<typename accessor::const_accessor>;
        std::array<gko::acc::size_type, 1> ts{nt}, one{1};
        auto t_acc = c_range(t_info.size, dT, ts);
        auto x_acc = range(b_info.size, dx, one);
        const dim3 block(32, 4, 1);
        const dim3 grid(static_cast<unsigned>(ceildiv(nt, size_type{32})), 1, 1);
        kernel::trsv_init<<<1, 1>>>(helper);
        if (upper) {
            if (unit) {
                kernel::acc_upper_trsv<32, 4, dmtx_t::unit>
                    <<<grid, block>>>(t_acc, x_acc, helper);
            } else {
                kernel::acc_upper_trsv<32, 4, dmtx_t::non_unit>
                    <<<grid, block>>>(t_acc, x_acc, helper);
            }
        } else {
            if (unit) {
                kernel::acc_lower_trsv<32, 4, dmtx_t::unit>
                    <<<grid, block>>>(t_acc, x_acc, helper);
            } else {
                kernel::acc_lower_trsv<32, 4, dmtx_t::non_unit>
                    <<<grid, block>>>(t_acc, x_acc, helper);
            }
        }
        synchronize();
        e = l1_rel(sol, widen(to_host(dx, nt)));
        check(e < 2e-7, "kernel::acc_{lower,upper}_trsv<32,4>", e, variant);
    }

    if (failures == 0) {
        std::printf("PASS\n");
        return 0;
    }
    return 1;
}
