// TRSV benchmark driver: same flags, sweep and CSV as the reference's
// (/root/reference/cuda/trsv_benchmark.cu) on top of the accblas launchers.
//   trsv_benchmark [--error] [--size=N] [--exact] [--lower] [--non-unit]
// The fixture is built like the reference's (cuda/trsv_memory.cuh:110-169):
// uniform(-1,1) matrix, factorised in place with cusolverDnDgetrf, which sees
// the row-major data as its transpose -- so the row-major STRICT UPPER
// triangle holds L^T (unit diagonal, well conditioned; the reference's
// compile-time default upper + unit) and the LOWER triangle holds U^T.
// --lower selects tmtx_t::lower; it is combined with a device-side transpose
// of the factors so the lower triangle is L (well conditioned, unit) unless
// --non-unit asks for U^T as stored.
#include <cusolverDn.h>

#include <accblas/trsv_kernels.cuh>

#include "driver_common.cuh"

namespace {

#define CUSOLVER_CALL(call)                                                 \
    do {                                                                    \
        const cusolverStatus_t status_ = (call);                            \
        if (status_ != CUSOLVER_STATUS_SUCCESS) {                           \
            std::cerr << "cuSOLVER error in file " << __FILE__              \
                      << " L:" << __LINE__ << "; Error: " << status_ << '\n';\
            throw std::runtime_error("cuSOLVER error " +                    \
                                     std::to_string(status_));              \
        }                                                                   \
    } while (false)

template <typename T>
struct Fixture {
    driver::DeviceBuffer<T> mtx, x, x_init;
    explicit Fixture(std::int64_t n) : mtx(n * n), x(n), x_init(n) {}
    void reset() { x.copy_from(x_init); }
};

__global__ void transpose_in_place(double* a, std::int64_t n)
{
    const std::int64_t r = blockIdx.y * std::int64_t{blockDim.y} + threadIdx.y;
    const std::int64_t c = blockIdx.x * std::int64_t{blockDim.x} + threadIdx.x;
    if (r < n && c < r) {
        const double lo = a[r * n + c];
        a[r * n + c] = a[c * n + r];
        a[c * n + r] = lo;
    }
}

void factorize(driver::DeviceBuffer<double>& mtx, std::int64_t n)
{
    cusolverDnHandle_t handle;
    CUSOLVER_CALL(cusolverDnCreate(&handle));
    int lwork = 0;
    CUSOLVER_CALL(cusolverDnDgetrf_bufferSize(handle, static_cast<int>(n),
                                              static_cast<int>(n), mtx.data(),
                                              static_cast<int>(n), &lwork));
    driver::DeviceBuffer<double> work(lwork);
    driver::DeviceBuffer<int> pivot(n), info(1);
    CUSOLVER_CALL(cusolverDnDgetrf(handle, static_cast<int>(n),
                                   static_cast<int>(n), mtx.data(),
                                   static_cast<int>(n), work.data(),
                                   pivot.data(), info.data()));
    synchronize();
    const int status = info.to_host(1)[0];
    if (status != 0) {
        std::cerr << "getrf reported info = " << status << '\n';
    }
    cusolverDnDestroy(handle);
}

}  // namespace

int main(int argc, char** argv)
{
    using ar_type = double;
    using st_type = float;
    using size_type = matrix_info::size_type;
    constexpr size_type default_max_size{24 * 1000}, min_size{100};

    // driver-specific switches first, the rest goes to the common parser
    tmtx_t t_matrix_type = tmtx_t::upper;
    dmtx_t d_matrix_type = dmtx_t::unit;
    std::vector<char*> rest{argv[0]};
    for (int i = 1; i < argc; ++i) {
        const std::string cur(argv[i]);
        if (cur == "--lower") {
            t_matrix_type = tmtx_t::lower;
        } else if (cur == "--non-unit") {
            d_matrix_type = dmtx_t::non_unit;
        } else {
            rest.push_back(argv[i]);
        }
    }
    driver::Options opt;
    if (!driver::parse(static_cast<int>(rest.size()), rest.data(),
                       default_max_size, min_size, "TRSVs", opt)) {
        return 1;
    }
    const size_type N = opt.max_size;
    const bool measure_error = opt.measure_error;
    auto h = accblas_detail::default_handle();

    Fixture<ar_type> ar(N);
    driver::fill(h, N, N, N, 0, ar.mtx);
    driver::fill(h, N, 1, 1, static_cast<std::uint64_t>(N) * N, ar.x_init);
    factorize(ar.mtx, N);
    const bool want_l = d_matrix_type == dmtx_t::unit;
    // after getrf: upper = L^T (unit), lower = U^T.  Transpose when the
    // requested triangle would otherwise hold the wrong factor.
    const bool l_is_upper = true;
    if ((t_matrix_type == tmtx_t::upper) != (want_l == l_is_upper)) {
        const dim3 block(32, 8);
        const dim3 grid(static_cast<unsigned>(ceildiv(N, size_type{32})),
                        static_cast<unsigned>(ceildiv(N, size_type{8})));
        transpose_in_place<<<grid, block>>>(ar.mtx.data(), N);
        CUDA_CALL(cudaGetLastError());
    }
    ar.reset();
    Fixture<st_type> st(N);
    driver::convert(h, N, N, N, ar.mtx, st.mtx);
    driver::convert(h, N, 1, 1, ar.x_init, st.x_init);
    st.reset();
    auto cublas = cublas_get_handle();
    driver::DeviceBuffer<std::uint32_t> trsv_helper(2);

    std::vector<ar_type> ref, tmp(static_cast<std::size_t>(N));
    ar_type ref_norm{1.0};
    auto error_of = [&](auto& fix, matrix_info info) {
        const auto raw = fix.x.to_host(info.size[0]);
        std::vector<ar_type> got(raw.size());
        for (std::size_t i = 0; i < raw.size(); ++i) {
            got[i] = driver::widen(raw[i]);
        }
        const ar_type err =
            compare(info, ref.data(), got.data(), tmp.data()) / ref_norm;
        fix.reset();
        return err;
    };

    using run_t = std::function<void(matrix_info, matrix_info)>;
    using err_t = std::function<ar_type(matrix_info)>;
    const auto tt = t_matrix_type;
    const auto dt = d_matrix_type;
    std::vector<std::tuple<std::string, run_t, err_t>> variants = {
        {"TRSV fp64",
         [&](matrix_info m, matrix_info x) {
             trsv(m, tt, dt, ar.mtx.data(), x, ar.x.data(), trsv_helper.data());
         },
         [&](matrix_info i) { return error_of(ar, i); }},
        {"TRSV fp32",
         [&](matrix_info m, matrix_info x) {
             trsv(m, tt, dt, st.mtx.data(), x, st.x.data(), trsv_helper.data());
         },
         [&](matrix_info i) { return error_of(st, i); }},
        {"TRSV Acc<fp64, fp64>",
         [&](matrix_info m, matrix_info x) {
             acc_trsv<ar_type>(m, tt, dt, ar.mtx.data(), x, ar.x.data(),
                               trsv_helper.data());
         },
         [&](matrix_info i) { return error_of(ar, i); }},
        {"TRSV Acc<fp64, fp32>",
         [&](matrix_info m, matrix_info x) {
             acc_trsv<ar_type>(m, tt, dt, st.mtx.data(), x, st.x.data(),
                               trsv_helper.data());
         },
         [&](matrix_info i) { return error_of(st, i); }},
        {"TRSV Acc<fp32, fp32>",
         [&](matrix_info m, matrix_info x) {
             acc_trsv<st_type>(m, tt, dt, st.mtx.data(), x, st.x.data(),
                               trsv_helper.data());
         },
         [&](matrix_info i) { return error_of(st, i); }},
        {"CUBLAS TRSV fp64",
         [&](matrix_info m, matrix_info x) {
             cublas_trsv(cublas.get(), tt, dt, m, ar.mtx.data(), x, ar.x.data());
         },
         [&](matrix_info i) { return error_of(ar, i); }},
        {"CUBLAS TRSV fp32",
         [&](matrix_info m, matrix_info x) {
             cublas_trsv(cublas.get(), tt, dt, m, st.mtx.data(), x, st.x.data());
         },
         [&](matrix_info i) { return error_of(st, i); }},
    };

    std::cout << "Num rows";
    for (const auto& v : variants) {
        std::cout << driver::DELIM << (measure_error ? "Error " : "")
                  << std::get<0>(v);
    }
    std::cout << '\n';
    std::cout.precision(16);
    std::cout << std::scientific << std::showpos;

    std::vector<ar_type> results(variants.size());
    const size_type start = opt.only_max ? N : std::min(N, min_size);
    for (size_type n = start; n <= N; n += std::min(N, min_size)) {
        const matrix_info m_info{{n, n}, N};
        const matrix_info x_info{{n, 1}};
        if (measure_error) {
            ar.reset();
            std::get<1>(variants[0])(m_info, x_info);
            synchronize();
            ref = ar.x.to_host(n);
            auto copy = ref;
            ref_norm = reduce<ar_type>(x_info, copy.data(),
                                       [](ar_type a, ar_type b) {
                                           return std::abs(a) + std::abs(b);
                                       });
            ar.reset();
        }
        for (std::size_t i = 0; i < variants.size(); ++i) {
            // the solve is in place: x has to be restored before every call,
            // which the reference's timing loop does not do either when it is
            // only timing (the values blow up but the work is the same)
            auto call = [&]() { std::get<1>(variants[i])(m_info, x_info); };
            if (measure_error) {
                benchmark_function(call, true);
                results[i] = std::get<2>(variants[i])(x_info);
            } else {
                results[i] = benchmark_function(call, false);
                ar.reset();
                st.reset();
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
