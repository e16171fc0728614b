// Vendor baselines measured next to the accblas kernels (NOT on the product
// path, built into a separate libaccblas_baselines.so): cuBLAS GEMV / DOT /
// TRSV called the way the reference's drivers call them for row-major data
// (/root/reference/cuda/gemv_kernels.cuh:232-243 -- OP_T with lda = row stride;
// cuda/dot_kernels.cuh:291-299 with the device pointer mode of
// cuda/dot_benchmark.cu:79; cuda/trsv_kernels.cuh:989-1008 -- fill mode
// swapped + OP_T).  The 64-bit entry points are used so n = 2^32 works.
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

namespace {

cublasHandle_t g_handle = nullptr;

int ensure_handle(cudaStream_t stream, cublasPointerMode_t mode)
{
    if (g_handle == nullptr) {
        if (cublasCreate(&g_handle) != CUBLAS_STATUS_SUCCESS) {
            g_handle = nullptr;
            return 3;
        }
    }
    if (cublasSetStream(g_handle, stream) != CUBLAS_STATUS_SUCCESS ||
        cublasSetPointerMode(g_handle, mode) != CUBLAS_STATUS_SUCCESS) {
        return 3;
    }
    return 0;
}

int check(cublasStatus_t st, const char* what)
{
    if (st != CUBLAS_STATUS_SUCCESS) {
        fprintf(stderr, "accblas_baselines: %s failed with cuBLAS status %d\n",
                what, static_cast<int>(st));
        return 3;
    }
    return 0;
}

}  // namespace

extern "C" {

// dtype: 0 = fp64, 1 = fp32 (same codes as accblas_dtype)
int accblas_baseline_cublas_gemv(int dtype, int64_t m, int64_t n, double alpha,
                                 const void* A, int64_t lda, const void* x,
                                 int64_t incx, double beta, void* y,
                                 int64_t incy, void* stream)
{
    if (ensure_handle(static_cast<cudaStream_t>(stream),
                      CUBLAS_POINTER_MODE_HOST)) {
        return 3;
    }
    // row-major m x n == column-major n x m, transposed
    if (dtype == 0) {
        return check(cublasDgemv_64(g_handle, CUBLAS_OP_T, n, m, &alpha,
                                    static_cast<const double*>(A), lda,
                                    static_cast<const double*>(x), incx, &beta,
                                    static_cast<double*>(y), incy),
                     "cublasDgemv");
    }
    const float a = static_cast<float>(alpha), b = static_cast<float>(beta);
    return check(cublasSgemv_64(g_handle, CUBLAS_OP_T, n, m, &a,
                                static_cast<const float*>(A), lda,
                                static_cast<const float*>(x), incx, &b,
                                static_cast<float*>(y), incy),
                 "cublasSgemv");
}

// result: DEVICE pointer
int accblas_baseline_cublas_dot(int dtype, int64_t n, const void* x,
                                int64_t incx, const void* y, int64_t incy,
                                void* result, void* stream)
{
    if (ensure_handle(static_cast<cudaStream_t>(stream),
                      CUBLAS_POINTER_MODE_DEVICE)) {
        return 3;
    }
    if (dtype == 0) {
        return check(cublasDdot_64(g_handle, n, static_cast<const double*>(x),
                                   incx, static_cast<const double*>(y), incy,
                                   static_cast<double*>(result)),
                     "cublasDdot");
    }
    return check(cublasSdot_64(g_handle, n, static_cast<const float*>(x), incx,
                               static_cast<const float*>(y), incy,
                               static_cast<float*>(result)),
                 "cublasSdot");
}

// upper/unit refer to the ROW-MAJOR matrix, as in accblas_trsv
int accblas_baseline_cublas_trsv(int dtype, int upper, int unit, int64_t n,
                                 const void* A, int64_t lda, void* x,
                                 int64_t incx, void* stream)
{
    if (ensure_handle(static_cast<cudaStream_t>(stream),
                      CUBLAS_POINTER_MODE_HOST)) {
        return 3;
    }
    const cublasFillMode_t uplo =
        upper ? CUBLAS_FILL_MODE_LOWER : CUBLAS_FILL_MODE_UPPER;
    const cublasDiagType_t diag =
        unit ? CUBLAS_DIAG_UNIT : CUBLAS_DIAG_NON_UNIT;
    if (dtype == 0) {
        return check(cublasDtrsv_64(g_handle, uplo, CUBLAS_OP_T, diag, n,
                                    static_cast<const double*>(A), lda,
                                    static_cast<double*>(x), incx),
                     "cublasDtrsv");
    }
    return check(cublasStrsv_64(g_handle, uplo, CUBLAS_OP_T, diag, n,
                                static_cast<const float*>(A), lda,
                                static_cast<float*>(x), incx),
                 "cublasStrsv");
}

}  // extern "C"
