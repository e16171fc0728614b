// GEMV benchmark driver: same flags, sweep and CSV as the reference's
// (/root/reference/cuda/gemv_benchmark.cu) on top of the accblas launchers.
//   gemv_benchmark [--error] [--size=N] [--fp16] [--exact]
// Columns: Num rows;GEMV fp64;GEMV fp32;GEMV Acc<fp64, fp64>;GEMV Acc<fp64,
// fp32>;GEMV Acc<fp32, fp32>;CUBLAS GEMV fp64;CUBLAS GEMV fp32  (milliseconds,
// min of 10; with --error: L1 error relative to the first column's result).
#include <cuda_fp16.h>

#include <accblas/gemv_kernels.cuh>

#include "driver_common.cuh"

namespace {

template <typename T>
struct Fixture {
    driver::DeviceBuffer<T> mtx, x, res, res_init;
    Fixture(std::int64_t n) : mtx(n * n), x(n), res(n), res_init(n) {}
    void reset() { res.copy_from(res_init); }
};

}  // namespace

int main(int argc, char** argv)
{
    using ar_type = double;
    using st_type = float;
    using size_type = matrix_info::size_type;
    constexpr ar_type alpha{1.0}, beta{1.0};
    constexpr size_type default_max_size{24500}, min_size{100};

    driver::Options opt;
    if (!driver::parse(argc, argv, default_max_size, min_size, "GEMVs", opt)) {
        return 1;
    }
    const size_type N = opt.max_size;
    const bool measure_error = opt.measure_error;
    auto h = accblas_detail::default_handle();

    // draw order of the reference fixture: matrix, x, res
    Fixture<ar_type> ar(N);
    driver::fill(h, N, N, N, 0, ar.mtx);
    driver::fill(h, N, 1, 1, static_cast<std::uint64_t>(N) * N, ar.x);
    driver::fill(h, N, 1, 1, static_cast<std::uint64_t>(N) * N + N, ar.res_init);
    ar.reset();
    Fixture<st_type> st(N);
    driver::convert(h, N, N, N, ar.mtx, st.mtx);
    driver::convert(h, N, 1, 1, ar.x, st.x);
    driver::convert(h, N, 1, 1, ar.res_init, st.res_init);
    st.reset();
    std::unique_ptr<Fixture<__half>> hf;
    if (opt.fp16) {
        hf = std::make_unique<Fixture<__half>>(N);
        driver::convert(h, N, N, N, ar.mtx, hf->mtx);
        driver::convert(h, N, 1, 1, ar.x, hf->x);
        driver::convert(h, N, 1, 1, ar.res_init, hf->res_init);
        hf->reset();
    }
    auto cublas = cublas_get_handle();

    // reference result (first variant) for the error mode
    std::vector<ar_type> ref, tmp(static_cast<std::size_t>(N));
    ar_type ref_norm{1.0};
    auto error_of = [&](auto& fix, matrix_info info) {
        ar_type err{};
        if (measure_error) {
            const auto raw = fix.res.to_host(info.size[0]);
            std::vector<ar_type> got(raw.size());
            for (std::size_t i = 0; i < raw.size(); ++i) {
                got[i] = driver::widen(raw[i]);
            }
            err = compare(info, ref.data(), got.data(), tmp.data()) / ref_norm;
            fix.reset();
        }
        return err;
    };

    using run_t = std::function<void(matrix_info, matrix_info, matrix_info)>;
    using err_t = std::function<ar_type(matrix_info)>;
    std::vector<std::tuple<std::string, run_t, err_t>> variants = {
        {"GEMV fp64",
         [&](matrix_info m, matrix_info x, matrix_info r) {
             gemv(m, alpha, ar.mtx.data(), x, ar.x.data(), r, beta, ar.res.data());
         },
         [&](matrix_info i) { return error_of(ar, i); }},
        {"GEMV fp32",
         [&](matrix_info m, matrix_info x, matrix_info r) {
             gemv(m, st_type(alpha), st.mtx.data(), x, st.x.data(), r,
                  st_type(beta), st.res.data());
         },
         [&](matrix_info i) { return error_of(st, i); }},
        {"GEMV Acc<fp64, fp64>",
         [&](matrix_info m, matrix_info x, matrix_info r) {
             acc_gemv<ar_type>(m, alpha, ar.mtx.data(), x, ar.x.data(), r, beta,
                               ar.res.data());
         },
         [&](matrix_info i) { return error_of(ar, i); }},
        {"GEMV Acc<fp64, fp32>",
         [&](matrix_info m, matrix_info x, matrix_info r) {
             acc_gemv<ar_type>(m, alpha, st.mtx.data(), x, st.x.data(), r, beta,
                               st.res.data());
         },
         [&](matrix_info i) { return error_of(st, i); }},
        {"GEMV Acc<fp32, fp32>",
         [&](matrix_info m, matrix_info x, matrix_info r) {
             acc_gemv<st_type>(m, st_type(alpha), st.mtx.data(), x, st.x.data(),
                               r, st_type(beta), st.res.data());
         },
         [&](matrix_info i) { return error_of(st, i); }},
        {"CUBLAS GEMV fp64",
         [&](matrix_info m, matrix_info x, matrix_info r) {
             cublas_gemv(cublas.get(), m, alpha, ar.mtx.data(), x, ar.x.data(),
                         r, beta, ar.res.data());
         },
         [&](matrix_info i) { return error_of(ar, i); }},
        {"CUBLAS GEMV fp32",
         [&](matrix_info m, matrix_info x, matrix_info r) {
             cublas_gemv(cublas.get(), m, st_type(alpha), st.mtx.data(), x,
                         st.x.data(), r, st_type(beta), st.res.data());
         },
         [&](matrix_info i) { return error_of(st, i); }},
    };
    if (opt.fp16) {
        variants.push_back(
            {"GEMV Acc<fp64, fp16>",
             [&](matrix_info m, matrix_info x, matrix_info r) {
                 acc_gemv<ar_type>(m, alpha, hf->mtx.data(), x, hf->x.data(), r,
                                   beta, hf->res.data());
             },
             [&](matrix_info i) { return error_of(*hf, i); }});
        variants.push_back(
            {"GEMV Acc<fp32, fp16>",
             [&](matrix_info m, matrix_info x, matrix_info r) {
                 acc_gemv<st_type>(m, st_type(alpha), hf->mtx.data(), x,
                                   hf->x.data(), r, st_type(beta),
                                   hf->res.data());
             },
             [&](matrix_info i) { return error_of(*hf, i); }});
    }

    std::cout << "Num rows";
    for (const auto& v : variants) {
        std::cout << driver::DELIM << (measure_error ? "Error " : "")
                  << std::get<0>(v);
    }
    std::cout << '\n';
    std::cout.precision(16);
    std::cout << std::scientific << std::showpos;

    std::vector<ar_type> results(variants.size());
    for (size_type n = opt.only_max ? N : min_size; n <= N; n += min_size) {
        const matrix_info m_info{{n, n}, N};
        const matrix_info x_info{{n, 1}};
        const matrix_info res_info{{n, 1}};
        if (measure_error) {
            ar.reset();
            std::get<1>(variants[0])(m_info, x_info, res_info);
            synchronize();
            ref = ar.res.to_host(n);
            auto copy = ref;
            ref_norm = reduce<ar_type>(res_info, copy.data(),
                                       [](ar_type a, ar_type b) {
                                           return std::abs(a) + std::abs(b);
                                       });
            ar.reset();
        }
        for (std::size_t i = 0; i < variants.size(); ++i) {
            auto call = [&]() {
                std::get<1>(variants[i])(m_info, x_info, res_info);
            };
            if (measure_error) {
                benchmark_function(call, true);
                results[i] = std::get<2>(variants[i])(x_info);
            } else {
                results[i] = benchmark_function(call, false);
            }
        }
        std::cout << n;
        for (const auto& r : results) {
            std::cout << driver::DELIM << r;
        }
        std::cout << '\n';
    }
    return 0;
}
