// Host utilities of the drop-in C++ layer: the descriptor, error macros,
// timing protocol and error metric the reference's drivers and launchers use
// (/root/reference/cuda/utils.cuh).  API-compatible, written for this project;
// everything is `inline` so the header can be included from several
// translation units.
#pragma once

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include <cublas_v2.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../accblas.h"

// Row-major matrix descriptor: 2-D extent and the distance between rows in
// elements (cuda/utils.cuh:18-56).
struct matrix_info {
    using size_type = std::int64_t;
    const std::array<size_type, 2> size;
    const size_type stride;

    constexpr matrix_info(const std::array<size_type, 2> size_,
                          const size_type stride_)
        : size(size_), stride{stride_}
    {}
    // contiguous rows
    constexpr matrix_info(const std::array<size_type, 2> size_)
        : matrix_info{size_, size_[1]}
    {}

    // elements spanned including the padding between rows
    size_type get_1d_size() const { return size[0] * stride; }
    // elements that belong to the matrix
    size_type get_num_elems() const { return size[0] * size[1]; }
};

template <typename ValueType>
constexpr ValueType ceildiv(ValueType a, ValueType b)
{
    return (a <= 0) ? a / b : (a - 1) / b + 1;
}

#define CUDA_CALL(call)                                                     \
    do {                                                                    \
        const cudaError_t accblas_err_ = (call);                            \
        if (accblas_err_ != cudaSuccess) {                                  \
            std::cerr << "Cuda error in file " << __FILE__                  \
                      << " L:" << __LINE__ << "; Error: "                   \
                      << cudaGetErrorString(accblas_err_) << '\n';          \
            throw std::runtime_error(cudaGetErrorString(accblas_err_));     \
        }                                                                   \
    } while (false)

#define CUBLAS_CALL(call)                                                   \
    do {                                                                    \
        const cublasStatus_t accblas_err_ = (call);                         \
        if (accblas_err_ != CUBLAS_STATUS_SUCCESS) {                        \
            std::cerr << "CuBLAS error in file " << __FILE__                \
                      << " L:" << __LINE__ << "; Error: " << accblas_err_   \
                      << '\n';                                              \
            throw std::runtime_error(std::string("Error: ") +               \
                                     std::to_string(accblas_err_));         \
        }                                                                   \
    } while (false)

// accblas C-ABI status -> exception, same behaviour as the macros above
#define ACCBLAS_CALL(call)                                                  \
    do {                                                                    \
        const int accblas_status_ = (call);                                 \
        if (accblas_status_ != ACCBLAS_OK) {                                \
            std::cerr << "accblas error in file " << __FILE__               \
                      << " L:" << __LINE__ << "; Error: "                   \
                      << accblas_last_error() << '\n';                      \
            throw std::runtime_error(accblas_last_error());                 \
        }                                                                   \
    } while (false)

inline void synchronize() { CUDA_CALL(cudaDeviceSynchronize()); }

// storage / arithmetic type -> accblas_dtype tag
namespace accblas_detail {
template <typename T>
struct dtype_of;
template <>
struct dtype_of<double> {
    static constexpr accblas_dtype value = ACCBLAS_F64;
};
template <>
struct dtype_of<float> {
    static constexpr accblas_dtype value = ACCBLAS_F32;
};
template <>
struct dtype_of<__half> {
    static constexpr accblas_dtype value = ACCBLAS_F16;
};

// One lazily created handle per device for the launchers whose reference
// signature carries no handle (gemv, trsv).
inline accblas_handle_t default_handle()
{
    static accblas_handle_t handles[64] = {};
    int device = 0;
    CUDA_CALL(cudaGetDevice(&device));
    if (device < 0 || device >= 64) {
        throw std::runtime_error("accblas: device index out of range");
    }
    if (handles[device] == nullptr) {
        ACCBLAS_CALL(accblas_create(&handles[device], device));
    }
    return handles[device];
}
}  // namespace accblas_detail

// ---------------------------------------------------------------------------
// Host plumbing the drivers use by name: cublas_get_handle(),
// cublas_set_device_ptr_mode(), benchmark_function().  Own implementation; only
// the names, the argument meaning and the timing protocol are the reference's.
// ---------------------------------------------------------------------------
namespace accblas_detail {

// cuBLAS handle with scalar arguments read from the host by default
struct cublas_handle_deleter {
    void operator()(cublasContext* handle) const
    {
        if (handle != nullptr) {
            cublasDestroy(handle);
        }
    }
};

// Two events on the default stream that live as long as the stopwatch.
class stopwatch {
public:
    stopwatch()
    {
        CUDA_CALL(cudaEventCreate(&events_[0]));
        CUDA_CALL(cudaEventCreate(&events_[1]));
    }
    ~stopwatch()
    {
        cudaEventDestroy(events_[0]);
        cudaEventDestroy(events_[1]);
    }
    stopwatch(const stopwatch&) = delete;
    stopwatch& operator=(const stopwatch&) = delete;

    // milliseconds the default stream spent on `work`
    template <typename Callable>
    double time(Callable& work)
    {
        CUDA_CALL(cudaEventRecord(events_[0], nullptr));
        work();
        CUDA_CALL(cudaEventRecord(events_[1], nullptr));
        CUDA_CALL(cudaEventSynchronize(events_[1]));
        float elapsed_ms = 0.0f;
        CUDA_CALL(cudaEventElapsedTime(&elapsed_ms, events_[0], events_[1]));
        return static_cast<double>(elapsed_ms);
    }

private:
    cudaEvent_t events_[2];
};

}  // namespace accblas_detail

inline std::unique_ptr<cublasContext, accblas_detail::cublas_handle_deleter>
cublas_get_handle()
{
    cublasHandle_t raw = nullptr;
    CUBLAS_CALL(cublasCreate(&raw));
    std::unique_ptr<cublasContext, accblas_detail::cublas_handle_deleter> owner(
        raw);
    CUBLAS_CALL(cublasSetPointerMode(raw, CUBLAS_POINTER_MODE_HOST));
    return owner;
}

// scalar results / arguments live in device memory (DOT writes its result
// there, cuda/dot_benchmark.cu:79)
inline void cublas_set_device_ptr_mode(cublasHandle_t handle)
{
    CUBLAS_CALL(cublasSetPointerMode(handle, CUBLAS_POINTER_MODE_DEVICE));
}

// The reference's timing protocol (cuda/utils.cuh:236-262): one warm-up call,
// then ten single calls, each bracketed by CUDA events on the default stream;
// the MINIMUM in milliseconds is what the drivers print.  skip == true (error
// mode): run once so that the result exists, report 0.
template <typename Callable>
double benchmark_function(Callable func, bool skip = false)
{
    func();
    synchronize();
    if (skip) {
        return 0.0;
    }
    accblas_detail::stopwatch watch;
    double fastest = std::numeric_limits<double>::infinity();
    for (int repetition = 0; repetition < 10; ++repetition) {
        fastest = std::min(fastest, watch.time(func));
    }
    return fastest;
}

// Pairwise (halving) reduction of a strided single-column vector, in place
// (cuda/utils.cuh:281-300): the summation order of the reference's norms.
template <typename OutputType, typename InputType, typename ReduceOp>
OutputType reduce(const matrix_info info, InputType* tmp, ReduceOp op)
{
    assert(info.size[1] == 1);
    const std::int64_t n = info.size[0];
    std::int64_t live = n;
    std::int64_t half = ceildiv(n, std::int64_t{2});
    while (half > 1) {
        for (std::int64_t i = 0; i + half < live; ++i) {
            tmp[i * info.stride] =
                op(tmp[i * info.stride], tmp[(i + half) * info.stride]);
        }
        live = half;
        half = ceildiv(half, std::int64_t{2});
    }
    return static_cast<OutputType>(n == 1 ? op(tmp[0], {})
                                          : op(tmp[0], tmp[info.stride]));
}

// sum_i |mtx1_i - mtx2_i| with the tree above (cuda/utils.cuh:315-332)
template <typename ReferenceType, typename OtherType, typename ValueType>
ValueType compare(const matrix_info info, const ReferenceType* mtx1,
                  const OtherType* mtx2, ValueType* tmp)
{
    assert(info.size[1] == 1);
    for (matrix_info::size_type row = 0; row < info.size[0]; ++row) {
        const auto idx = row * info.stride;
        const ValueType a = static_cast<ValueType>(mtx1[idx]);
        const ValueType b = static_cast<ValueType>(mtx2[idx]);
        tmp[idx] = std::abs(a - b);
    }
    return reduce<ValueType>(info, tmp,
                             [](ValueType a, ValueType b) { return a + b; });
}
