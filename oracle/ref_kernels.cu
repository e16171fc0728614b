// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Thin extern "C" shim around the REFERENCE's own, unmodified CUDA launchers.
// The three kernel headers are #included from where they lie under
// /root/reference/cuda (the Makefile passes -I$(REFERENCE)/cuda; nothing is
// copied into this repository) and compiled for sm_100a against the clean-room
// accessor headers in ../include/accessor.  The result,
// oracle/_ref/libref_kernels.so, is what the GPU parity tests, the golden
// fixture generator (tests/golden/make_golden.py) and tools/compare_reference.py
// use as "the reference's own CUDA kernels on the same B200".
//
// All launches go to the legacy default stream, as in the reference.
#include <cuda_fp16.h>

#include <cstdint>
#include <iostream>
#include <memory>

#include "dot_kernels.cuh"
#include "gemv_kernels.cuh"
#include "trsv_kernels.cuh"

namespace {

enum { F64 = 0, F32 = 1, F16 = 2 };

matrix_info vec_info(std::int64_t n, std::int64_t inc)
{
    return matrix_info{{n, 1}, inc};
}

myBlasHandle* dot_handle()
{
    static myBlasHandle* handle = new myBlasHandle();
    return handle;
}

cublasHandle_t cublas_handle(bool device_pointer_mode)
{
    static cublasHandle_t handle = [] {
        cublasHandle_t h;
        CUBLAS_CALL(cublasCreate(&h));
        return h;
    }();
    CUBLAS_CALL(cublasSetPointerMode(handle, device_pointer_mode
                                                 ? CUBLAS_POINTER_MODE_DEVICE
                                                 : CUBLAS_POINTER_MODE_HOST));
    return handle;
}

template <typename F>
int guarded(F&& f)
{
    try {
        f();
    } catch (const std::exception& e) {
        std::cerr << "ref_kernels: " << e.what() << '\n';
        return 3;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

}  // namespace

extern "C" {

int ref_sync() { return cudaDeviceSynchronize() == cudaSuccess ? 0 : 3; }

int ref_sm_count()
{
    return dot_handle()->get_device_property().multiProcessorCount;
}

// plain != 0: gemv<T> (cuda/gemv_kernels.cuh:136-147), T given by `st`
// plain == 0: acc_gemv<Ar, St> (cuda/gemv_kernels.cuh:168-193)
int ref_gemv(int ar, int st, int plain, std::int64_t m, std::int64_t n,
             double alpha, const void* A, std::int64_t lda, const void* x,
             std::int64_t incx, double beta, void* y, std::int64_t incy)
{
    const matrix_info mi{{m, n}, lda};
    const matrix_info xi = vec_info(n, incx);
    const matrix_info yi = vec_info(m, incy);
    return guarded([&] {
        auto run_acc = [&](auto a, auto s) {
            using Ar = decltype(a);
            using St = decltype(s);
            acc_gemv<Ar, St>(mi, static_cast<Ar>(alpha),
                             static_cast<const St*>(A), xi,
                             static_cast<const St*>(x), yi,
                             static_cast<Ar>(beta), static_cast<St*>(y));
        };
        if (plain) {
            if (st == F64) {
                gemv<double>(mi, alpha, static_cast<const double*>(A), xi,
                             static_cast<const double*>(x), yi, beta,
                             static_cast<double*>(y));
            } else if (st == F32) {
                gemv<float>(mi, static_cast<float>(alpha),
                            static_cast<const float*>(A), xi,
                            static_cast<const float*>(x), yi,
                            static_cast<float>(beta), static_cast<float*>(y));
            } else {
                throw std::runtime_error("plain gemv: fp64/fp32 only");
            }
        } else if (ar == F64) {
            if (st == F64) run_acc(double{}, double{});
            else if (st == F32) run_acc(double{}, float{});
            else run_acc(double{}, __half{});
        } else if (ar == F32) {
            if (st == F64) run_acc(float{}, double{});
            else if (st == F32) run_acc(float{}, float{});
            else run_acc(float{}, __half{});
        } else {
            throw std::runtime_error("bad arithmetic type");
        }
    });
}

// plain != 0: dot<T> (cuda/dot_kernels.cuh:192-206), result type T
// plain == 0: acc_dot<Ar, St, Res> (cuda/dot_kernels.cuh:224-263);
//             Res is fp64 or fp32
int ref_dot(int ar, int st, int res, int plain, std::int64_t n, const void* x,
            std::int64_t incx, const void* y, std::int64_t incy, void* result)
{
    const matrix_info xi = vec_info(n, incx);
    const matrix_info yi = vec_info(n, incy);
    return guarded([&] {
        auto run_acc = [&](auto a, auto s) {
            using Ar = decltype(a);
            using St = decltype(s);
            if (res == F64) {
                acc_dot<Ar, St, double>(dot_handle(), xi,
                                        static_cast<const St*>(x), yi,
                                        static_cast<const St*>(y),
                                        static_cast<double*>(result));
            } else if (res == F32) {
                acc_dot<Ar, St, float>(dot_handle(), xi,
                                       static_cast<const St*>(x), yi,
                                       static_cast<const St*>(y),
                                       static_cast<float*>(result));
            } else {
                throw std::runtime_error("acc_dot: result fp64/fp32 only");
            }
        };
        if (plain) {
            if (st == F64) {
                dot<double>(dot_handle(), xi, static_cast<const double*>(x), yi,
                            static_cast<const double*>(y),
                            static_cast<double*>(result));
            } else if (st == F32) {
                dot<float>(dot_handle(), xi, static_cast<const float*>(x), yi,
                           static_cast<const float*>(y),
                           static_cast<float*>(result));
            } else {
                throw std::runtime_error("plain dot: fp64/fp32 only");
            }
        } else if (ar == F64) {
            if (st == F64) run_acc(double{}, double{});
            else if (st == F32) run_acc(double{}, float{});
            else run_acc(double{}, __half{});
        } else if (ar == F32) {
            if (st == F64) run_acc(float{}, double{});
            else if (st == F32) run_acc(float{}, float{});
            else run_acc(float{}, __half{});
        } else {
            throw std::runtime_error("bad arithmetic type");
        }
    });
}

// plain != 0: trsv<T> (cuda/trsv_kernels.cuh:455-488)
// plain == 0: acc_trsv<Ar, St> (cuda/trsv_kernels.cuh:918-961)
// helper: device pointer to two uint32 (cuda/trsv_benchmark.cu:98)
int ref_trsv(int ar, int st, int plain, int upper, int unit, std::int64_t n,
             const void* A, std::int64_t lda, void* x, std::int64_t incx,
             std::uint32_t* helper)
{
    const matrix_info mi{{n, n}, lda};
    const matrix_info xi = vec_info(n, incx);
    const tmtx_t tt = upper ? tmtx_t::upper : tmtx_t::lower;
    const dmtx_t dt = unit ? dmtx_t::unit : dmtx_t::non_unit;
    return guarded([&] {
        auto run_acc = [&](auto a, auto s) {
            using Ar = decltype(a);
            using St = decltype(s);
            acc_trsv<Ar, St>(mi, tt, dt, static_cast<const St*>(A), xi,
                             static_cast<St*>(x), helper);
        };
        if (plain) {
            if (st == F64) {
                trsv<double>(mi, tt, dt, static_cast<const double*>(A), xi,
                             static_cast<double*>(x), helper);
            } else if (st == F32) {
                trsv<float>(mi, tt, dt, static_cast<const float*>(A), xi,
                            static_cast<float*>(x), helper);
            } else {
                throw std::runtime_error("plain trsv: fp64/fp32 only");
            }
        } else if (ar == F64) {
            if (st == F64) run_acc(double{}, double{});
            else if (st == F32) run_acc(double{}, float{});
            else run_acc(double{}, __half{});
        } else if (ar == F32) {
            if (st == F64) run_acc(float{}, double{});
            else if (st == F32) run_acc(float{}, float{});
            else run_acc(float{}, __half{});
        } else {
            throw std::runtime_error("bad arithmetic type");
        }
    });
}

// The reference's cuBLAS wrappers (cuda/gemv_kernels.cuh:198-243,
// cuda/dot_kernels.cuh:268-299, cuda/trsv_kernels.cuh:964-1008).
int ref_cublas_gemv(int t, std::int64_t m, std::int64_t n, double alpha,
                    const void* A, std::int64_t lda, const void* x,
                    std::int64_t incx, double beta, void* y, std::int64_t incy)
{
    const matrix_info mi{{m, n}, lda};
    return guarded([&] {
        if (t == F64) {
            cublas_gemv<double>(cublas_handle(false), mi, alpha,
                                static_cast<const double*>(A), vec_info(n, incx),
                                static_cast<const double*>(x), vec_info(m, incy),
                                beta, static_cast<double*>(y));
        } else {
            cublas_gemv<float>(cublas_handle(false), mi,
                               static_cast<float>(alpha),
                               static_cast<const float*>(A), vec_info(n, incx),
                               static_cast<const float*>(x), vec_info(m, incy),
                               static_cast<float>(beta), static_cast<float*>(y));
        }
    });
}

// result is a DEVICE pointer (cuda/dot_benchmark.cu:79)
int ref_cublas_dot(int t, std::int64_t n, const void* x, std::int64_t incx,
                   const void* y, std::int64_t incy, void* result)
{
    return guarded([&] {
        if (t == F64) {
            cublas_dot<double>(cublas_handle(true), vec_info(n, incx),
                               static_cast<const double*>(x), vec_info(n, incy),
                               static_cast<const double*>(y),
                               static_cast<double*>(result));
        } else {
            cublas_dot<float>(cublas_handle(true), vec_info(n, incx),
                              static_cast<const float*>(x), vec_info(n, incy),
                              static_cast<const float*>(y),
                              static_cast<float*>(result));
        }
    });
}

int ref_cublas_trsv(int t, int upper, int unit, std::int64_t n, const void* A,
                    std::int64_t lda, void* x, std::int64_t incx)
{
    const matrix_info mi{{n, n}, lda};
    const tmtx_t tt = upper ? tmtx_t::upper : tmtx_t::lower;
    const dmtx_t dt = unit ? dmtx_t::unit : dmtx_t::non_unit;
    return guarded([&] {
        if (t == F64) {
            cublas_trsv<double>(cublas_handle(false), tt, dt, mi,
                                static_cast<const double*>(A), vec_info(n, incx),
                                static_cast<double*>(x));
        } else {
            cublas_trsv<float>(cublas_handle(false), tt, dt, mi,
                               static_cast<const float*>(A), vec_info(n, incx),
                               static_cast<float*>(x));
        }
    });
}

}  // extern "C"
