// DOT benchmark driver: same flags, sweep and CSV as the reference's
// (/root/reference/cuda/dot_benchmark.cu) on top of the accblas launchers.
//   dot_benchmark [--error] [--size=N] [--fp16] [--exact]
// Default: "Vector Size;<times in ms>;Error <relative errors>".  With --error:
// the MEDIAN relative error over ten re-randomised vector pairs, then a
// separator line and the raw results of every run.
#include <algorithm>
#include <array>

#include <cuda_fp16.h>

#include <accblas/dot_kernels.cuh>

#include "driver_common.cuh"

namespace {

template <typename T>
struct Fixture {
    driver::DeviceBuffer<T> x, y, res;
    explicit Fixture(std::int64_t n) : x(n), y(n), res(1) {}
    double result() const { return driver::widen(res.to_host(1)[0]); }
};

}  // namespace

int main(int argc, char** argv)
{
    using ar_type = double;
    using st_type = float;
    using size_type = matrix_info::size_type;
    constexpr size_type default_max_size{535 * 1000 * 1000};
    constexpr size_type min_size{1'000'000};
    constexpr size_type row_incr{2'000'000};
    constexpr size_type max_randomize_num{10};

    driver::Options opt;
    if (!driver::parse(argc, argv, default_max_size, min_size, "DOTs", opt)) {
        return 1;
    }
    const size_type N = opt.max_size;
    const bool detailed_error = opt.measure_error;

    auto my_handle = std::make_unique<myBlasHandle>();
    auto h = my_handle->get_accblas_handle();
    auto cublas = cublas_get_handle();
    cublas_set_device_ptr_mode(cublas.get());

    Fixture<ar_type> ar(N);
    Fixture<st_type> st(N);
    std::unique_ptr<Fixture<__half>> hf;
    if (opt.fp16) {
        hf = std::make_unique<Fixture<__half>>(N);
    }
    // draw order of the reference fixture: x then y; every re-randomisation
    // continues the same engine (cuda/dot_benchmark.cu:194-201)
    auto generate = [&](size_type round) {
        const std::uint64_t first = 2ull * N * round;
        driver::fill(h, 1, N, N, first, ar.x);
        driver::fill(h, 1, N, N, first + N, ar.y);
        driver::convert(h, 1, N, N, ar.x, st.x);
        driver::convert(h, 1, N, N, ar.y, st.y);
        if (hf) {
            driver::convert(h, 1, N, N, ar.x, hf->x);
            driver::convert(h, 1, N, N, ar.y, hf->y);
        }
    };
    generate(0);

    using run_t = std::function<void(matrix_info, matrix_info)>;
    using res_t = std::function<ar_type()>;
    std::vector<std::tuple<std::string, run_t, res_t>> variants = {
        {"DOT fp64",
         [&](matrix_info xi, matrix_info yi) {
             dot(my_handle.get(), xi, ar.x.data(), yi, ar.y.data(), ar.res.data());
         },
         [&]() { return ar.result(); }},
        {"DOT fp32",
         [&](matrix_info xi, matrix_info yi) {
             dot(my_handle.get(), xi, st.x.data(), yi, st.y.data(), st.res.data());
         },
         [&]() { return st.result(); }},
        {"DOT Acc<fp64, fp64>",
         [&](matrix_info xi, matrix_info yi) {
             acc_dot<ar_type>(my_handle.get(), xi, ar.x.data(), yi, ar.y.data(),
                              ar.res.data());
         },
         [&]() { return ar.result(); }},
        {"DOT Acc<fp64, fp32>",
         [&](matrix_info xi, matrix_info yi) {
             acc_dot<ar_type>(my_handle.get(), xi, st.x.data(), yi, st.y.data(),
                              st.res.data());
         },
         [&]() { return st.result(); }},
        {"DOT Acc<fp32, fp32>",
         [&](matrix_info xi, matrix_info yi) {
             acc_dot<st_type>(my_handle.get(), xi, st.x.data(), yi, st.y.data(),
                              st.res.data());
         },
         [&]() { return st.result(); }},
        {"CUBLAS DOT fp64",
         [&](matrix_info xi, matrix_info yi) {
             cublas_dot(cublas.get(), xi, ar.x.data(), yi, ar.y.data(),
                        ar.res.data());
         },
         [&]() { return ar.result(); }},
        {"CUBLAS DOT fp32",
         [&](matrix_info xi, matrix_info yi) {
             cublas_dot(cublas.get(), xi, st.x.data(), yi, st.y.data(),
                        st.res.data());
         },
         [&]() { return st.result(); }},
    };
    if (opt.fp16) {
        // fp16 storage: the scalar is kept in fp32 (a half cannot hold it)
        variants.push_back(
            {"DOT Acc<fp64, fp16>",
             [&](matrix_info xi, matrix_info yi) {
                 acc_dot<ar_type>(my_handle.get(), xi, hf->x.data(), yi,
                                  hf->y.data(), st.res.data());
             },
             [&]() { return st.result(); }});
        variants.push_back(
            {"DOT Acc<fp32, fp16>",
             [&](matrix_info xi, matrix_info yi) {
                 acc_dot<st_type>(my_handle.get(), xi, hf->x.data(), yi,
                                  hf->y.data(), st.res.data());
             },
             [&]() { return st.result(); }});
    }
    const size_type num = static_cast<size_type>(variants.size());

    std::cout << "Vector Size";
    if (!detailed_error) {
        for (const auto& v : variants) {
            std::cout << driver::DELIM << std::get<0>(v);
        }
    }
    for (const auto& v : variants) {
        std::cout << driver::DELIM << "Error " << std::get<0>(v);
    }
    std::cout << '\n';
    std::cout.precision(16);
    std::cout << std::scientific;

    std::vector<size_type> sizes;
    if (opt.only_max) {
        sizes.push_back(N);
    } else {
        for (size_type n = std::min(N, min_size); n <= N; n += row_incr) {
            sizes.push_back(n);
        }
    }
    const size_type rounds = detailed_error ? max_randomize_num : 1;
    const size_type steps = static_cast<size_type>(sizes.size());
    std::vector<double> times(steps * num);
    // [step][variant][round]
    std::vector<ar_type> raw(steps * num * rounds);
    auto raw_at = [&](size_type rnd, size_type step, size_type bi) -> ar_type& {
        return raw[(step * num + bi) * rounds + rnd];
    };

    for (size_type rnd = 0; rnd < rounds; ++rnd) {
        if (rnd != 0) {
            generate(rnd);
        }
        for (size_type i = 0; i < steps; ++i) {
            const matrix_info x_info{{sizes[i], 1}};
            const matrix_info y_info{{sizes[i], 1}};
            for (size_type bi = 0; bi < num; ++bi) {
                auto call = [&]() { std::get<1>(variants[bi])(x_info, y_info); };
                times[i * num + bi] = benchmark_function(call, detailed_error);
                raw_at(rnd, i, bi) = std::get<2>(variants[bi])();
            }
        }
    }

    auto rel_error = [](ar_type res, ar_type ref) {
        return std::abs(res - ref) / std::abs(ref);
    };
    for (size_type i = 0; i < steps; ++i) {
        std::cout << sizes[i];
        if (!detailed_error) {
            for (size_type bi = 0; bi < num; ++bi) {
                std::cout << driver::DELIM << times[i * num + bi];
            }
            for (size_type bi = 0; bi < num; ++bi) {
                std::cout << driver::DELIM
                          << rel_error(raw_at(0, i, bi), raw_at(0, i, 0));
            }
        } else {
            for (size_type bi = 0; bi < num; ++bi) {
                std::array<ar_type, max_randomize_num> errs{};
                for (size_type rnd = 0; rnd < rounds; ++rnd) {
                    errs[rnd] = rel_error(raw_at(rnd, i, bi), raw_at(rnd, i, 0));
                }
                std::sort(errs.begin(), errs.begin() + rounds);
                const ar_type median =
                    (rounds % 2 == 1)
                        ? errs[rounds / 2]
                        : (errs[rounds / 2 - 1] + errs[rounds / 2]) / 2.0;
                std::cout << driver::DELIM << median;
            }
        }
        std::cout << '\n';
    }
    if (!detailed_error) {
        return 0;
    }
    std::cout << "--------------------------------------------------\n";
    std::cout << "Random iter" << driver::DELIM << "Vector Size";
    for (const auto& v : variants) {
        std::cout << driver::DELIM << "Result " << std::get<0>(v);
    }
    std::cout << '\n';
    for (size_type i = 0; i < steps; ++i) {
        for (size_type rnd = 0; rnd < rounds; ++rnd) {
            std::cout << rnd << driver::DELIM << sizes[i];
            for (size_type bi = 0; bi < num; ++bi) {
                std::cout << driver::DELIM << raw_at(rnd, i, bi);
            }
            std::cout << '\n';
        }
    }
    return 0;
}
