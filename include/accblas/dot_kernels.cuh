// Drop-in replacement for the reference's dot_kernels.cuh
// (/root/reference/cuda/dot_kernels.cuh): same names, template parameters and
// argument meaning.
//
//   myBlasHandle    device properties + scratch, as in the reference; here it
//                   wraps an accblas handle (workspace for the deterministic
//                   two-pass reduction).
//   host launchers  dot<T>, acc_dot<Ar, St, Res>  -> accblas_dot: ONE launch
//                   (no init_res / cast_result launches, no atomics; the result
//                   is bit-reproducible for a fixed n and device).
//   kernel::dot<block_size, T>, kernel::acc_dot<block_size>(x, y, res)
//                   accessor-generic __global__ kernels with the documented
//                   signatures; like the reference's they ACCUMULATE into *res
//                   (initialise it with kernel::init_res) and combine block
//                   partials with atomic_add.
//   cublas_dot      vendor baseline.
#pragma once

#include <cinttypes>

#include <cublas_v2.h>

#include <accessor/range.hpp>
#include <accessor/reduced_row_major.hpp>

#include "atomics.cuh"
#include "kernel_utils.cuh"
#include "utils.cuh"

constexpr int grids_per_sm{32};
constexpr int dot_block_size{1024};

class myBlasHandle {
public:
    myBlasHandle()
    {
        int device = 0;
        CUDA_CALL(cudaGetDevice(&device));
        CUDA_CALL(cudaGetDeviceProperties(&device_prop_, device));
        ACCBLAS_CALL(accblas_create(&handle_, device));
        CUDA_CALL(cudaMalloc(&device_storage_, device_storage_size_bytes_));
    }
    myBlasHandle(const myBlasHandle&) = delete;
    myBlasHandle& operator=(const myBlasHandle&) = delete;
    ~myBlasHandle()
    {
        cudaFree(device_storage_);
        accblas_destroy(handle_);
    }

    const cudaDeviceProp& get_device_property() const { return device_prop_; }

    // device pointer to one scratch value of type T
    template <typename T>
    T* get_device_value_ptr()
    {
        static_assert(sizeof(T) < device_storage_size_bytes_,
                      "The expected type is too large for the device storage!");
        return reinterpret_cast<T*>(device_storage_);
    }

    accblas_handle_t get_accblas_handle() const { return handle_; }

private:
    static constexpr std::size_t device_storage_size_bytes_{16};
    cudaDeviceProp device_prop_;
    accblas_handle_t handle_{nullptr};
    void* device_storage_{nullptr};
};


namespace kernel {

template <typename ValueType>
__global__ __launch_bounds__(1) void init_res(ValueType* __restrict__ res)
{
    *res = ValueType{0};
}

// *res += sum_i x[i*x_stride] * y[i*y_stride]   (plain pointers)
template <std::int64_t block_size, typename ValueType>
__global__ __launch_bounds__(block_size) void dot(
    const std::int32_t n, const ValueType* __restrict__ x,
    const std::int32_t x_stride, const ValueType* __restrict__ y,
    const std::int32_t y_stride, ValueType* __restrict__ res)
{
    ValueType acc{};
    const std::int64_t step = std::int64_t{block_size} * gridDim.x;
    for (std::int64_t i = std::int64_t{blockIdx.x} * block_size + threadIdx.x;
         i < n; i += step) {
        acc += x[i * x_stride] * y[i * y_stride];
    }
    const ValueType total = detail::block_total<block_size>(acc);
    if (threadIdx.x == 0) {
        atomic_add(res, total);
    }
}

// The same through ranges (2-D, second index 0, as in the reference).
template <std::int64_t block_size, typename XRange, typename YRange,
          typename ResType>
__global__ __launch_bounds__(block_size) void acc_dot(XRange x, YRange y,
                                                      ResType* __restrict__ res)
{
    using ar_type = decltype(x(0, 0) + y(0, 0));
    ar_type acc{};
    const std::int64_t n = x.length(0);
    const std::int64_t step = std::int64_t{block_size} * gridDim.x;
    for (std::int64_t i = std::int64_t{blockIdx.x} * block_size + threadIdx.x;
         i < n; i += step) {
        acc += x(i, 0) * y(i, 0);
    }
    const ar_type total = detail::block_total<block_size>(acc);
    if (threadIdx.x == 0) {
        atomic_add(res, static_cast<ResType>(total));
    }
}

template <typename InType, typename OutType>
__global__ __launch_bounds__(1) void cast_result(const InType* __restrict__ in,
                                                 OutType* __restrict__ out)
{
    *out = static_cast<OutType>(*in);
}

}  // namespace kernel


// *res = x . y, all in ValueType (cuda/dot_kernels.cuh:192-206); `res` is a
// device pointer.
template <typename ValueType>
void dot(myBlasHandle* handle, const matrix_info x_info, const ValueType* x,
         const matrix_info y_info, const ValueType* y, ValueType* res)
{
    constexpr accblas_dtype t = accblas_detail::dtype_of<ValueType>::value;
    ACCBLAS_CALL(accblas_dot(handle->get_accblas_handle(), t, t, t,
                             x_info.size[0], x, x_info.stride, y, y_info.stride,
                             res, nullptr));
}

// *res = x . y accumulated in ArType on StType storage, stored as ResType
// (cuda/dot_kernels.cuh:224-263).
template <typename ArType, typename StType, typename ResType>
void acc_dot(myBlasHandle* handle, const matrix_info x_info, const StType* x,
             const matrix_info y_info, const StType* y, ResType* res)
{
    ACCBLAS_CALL(accblas_dot(handle->get_accblas_handle(),
                             accblas_detail::dtype_of<ArType>::value,
                             accblas_detail::dtype_of<StType>::value,
                             accblas_detail::dtype_of<ResType>::value,
                             x_info.size[0], x, x_info.stride, y, y_info.stride,
                             res, nullptr));
}


inline void cublas_dot(cublasHandle_t handle, int n, const double* x, int incx,
                       const double* y, int incy, double* res)
{
    CUBLAS_CALL(cublasDdot(handle, n, x, incx, y, incy, res));
}

inline void cublas_dot(cublasHandle_t handle, int n, const float* x, int incx,
                       const float* y, int incy, float* res)
{
    CUBLAS_CALL(cublasSdot(handle, n, x, incx, y, incy, res));
}

template <typename ValueType>
void cublas_dot(cublasHandle_t handle, const matrix_info x_info,
                const ValueType* x, const matrix_info y_info,
                const ValueType* y, ValueType* res)
{
    cublas_dot(handle, static_cast<int>(x_info.size[0]), x,
               static_cast<int>(x_info.stride), y,
               static_cast<int>(y_info.stride), res);
}
