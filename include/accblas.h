/*
 * accblas.h -- C ABI of the B200-native (sm_100a) accessor-BLAS hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes and
 * returns an int status (ACCBLAS_OK == 0).  Nothing here allocates behind the
 * caller's back except the per-handle workspace, and nothing falls back to a
 * CPU implementation: without a CUDA device every compute call fails with
 * ACCBLAS_ERR_CUDA.
 *
 * Operands follow the reference's conventions (row-major matrix with a row
 * stride in ELEMENTS, vectors with an element stride, one storage type for
 * all operands of a call, arithmetic in `ar`):
 *
 *   reference interface (file:line under /root/reference)     -> entry point
 *   acc_gemv<Ar,St>   cuda/gemv_kernels.cuh:168-193            -> accblas_gemv
 *   gemv<T>           cuda/gemv_kernels.cuh:136-147            -> accblas_gemv (ar == st)
 *   acc_dot<Ar,St,Res> cuda/dot_kernels.cuh:224-263            -> accblas_dot
 *   dot<T>            cuda/dot_kernels.cuh:192-206             -> accblas_dot (ar == st == res)
 *   acc_trsv<Ar,St>   cuda/trsv_kernels.cuh:918-961            -> accblas_trsv
 *   trsv<T>           cuda/trsv_kernels.cuh:455-488            -> accblas_trsv (ar == st)
 *   convert_mtx       cuda/matrix_helper.cuh:93-103            -> accblas_convert
 *   gen_mtx / write_random cuda/matrix_helper.cuh:28-75        -> accblas_fill_uniform
 *   myBlasHandle      cuda/dot_kernels.cuh:29-65               -> accblas_handle_t
 *
 * The header-only C++ layer in include/accblas/*.cuh keeps the reference's
 * templated launcher signatures and forwards to these functions.
 */
#ifndef ACCBLAS_H_
#define ACCBLAS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACCBLAS_VERSION 100

/* storage / arithmetic / result element types */
typedef enum accblas_dtype {
    ACCBLAS_F64 = 0,
    ACCBLAS_F32 = 1,
    ACCBLAS_F16 = 2 /* storage only */
} accblas_dtype;

typedef enum accblas_status {
    ACCBLAS_OK = 0,
    ACCBLAS_ERR_INVALID = 1,     /* bad argument (null pointer, negative size, stride < cols ...) */
    ACCBLAS_ERR_UNSUPPORTED = 2, /* dtype combination not instantiated (e.g. fp16 arithmetic) */
    ACCBLAS_ERR_CUDA = 3,        /* a CUDA runtime call or kernel launch failed */
    ACCBLAS_ERR_ALLOC = 4,       /* workspace allocation failed */
    ACCBLAS_ERR_DATA = 5,        /* data-dependent failure (non-normal draw in fill_uniform) */
    ACCBLAS_ERR_PEER = 6         /* multi-GPU exchange: a peer did not arrive within the time limit */
} accblas_status;

/* which triangle of the row-major matrix / whether the diagonal is implied
 * (mirrors tmtx_t / dmtx_t, cuda/trsv_kernels.cuh:22,29) */
typedef enum accblas_uplo { ACCBLAS_UPPER = 0, ACCBLAS_LOWER = 1 } accblas_uplo;
typedef enum accblas_diag { ACCBLAS_NON_UNIT = 0, ACCBLAS_UNIT = 1 } accblas_diag;

typedef struct accblas_handle_s* accblas_handle_t;

/* cudaStream_t without dragging cuda_runtime.h into C callers */
typedef void* accblas_stream_t;

int accblas_version(void);
const char* accblas_status_string(int status);
/* message of the last failing call on this thread ("" if none) */
const char* accblas_last_error(void);
size_t accblas_sizeof(accblas_dtype t);

/* Handle = device id + SM count + device workspace (DOT partials and the
 * completion counter, TRSV progress vector).  Calls sharing a handle must be
 * stream-ordered with respect to each other (same rule as myBlasHandle).
 * device < 0 selects the current device. */
int accblas_create(accblas_handle_t* handle, int device);
int accblas_destroy(accblas_handle_t handle);
int accblas_get_sm_count(accblas_handle_t handle, int* sm_count);

/* y = alpha * A * x + beta * y.
 * A: m x n, row-major, lda >= n elements of `st`; x: n elements with stride
 * incx; y: m elements with stride incy (read only if beta != 0, written
 * rounded to `st`).  alpha/beta are converted to `ar` before use.
 * All pointers are DEVICE pointers. */
int accblas_gemv(accblas_handle_t handle, accblas_dtype ar, accblas_dtype st,
                 int64_t m, int64_t n, double alpha, const void* A,
                 int64_t lda, const void* x, int64_t incx, double beta,
                 void* y, int64_t incy, accblas_stream_t stream);

/* *result = sum_i x[i*incx] * y[i*incy], accumulated in `ar` with a
 * deterministic two-pass tree (fixed for a given n, dtype pair and SM count),
 * written as `res` to the DEVICE pointer `result`.  Large contiguous operands:
 * the last ~12 % of the data is handed to the SMs dynamically in small chunks
 * (load balance), each chunk with its own partial sum folded in chunk order --
 * the bits do not depend on which SM took which chunk.  The partials are
 * folded in fp64 and rounded to `ar` once. */
int accblas_dot(accblas_handle_t handle, accblas_dtype ar, accblas_dtype st,
                accblas_dtype res, int64_t n, const void* x, int64_t incx,
                const void* y, int64_t incy, void* result,
                accblas_stream_t stream);

/* Multi-GPU DOT (BASELINE.json config 5; the reference is single-device,
 * cuda/dot_kernels.cuh:33): every rank calls accblas_dot_allreduce on its own
 * index range; the kernel's last CTA writes the rank's partial (in `ar`) into
 * every peer's mailbox over NVLink, waits for the peers' partials in its own
 * mailbox and sums them in rank order, so *result (as `res`) is the dot
 * product of the whole vectors on every rank, with identical bits on all of
 * them -- one launch per rank, no separate collective.  At most 8 ranks.
 *
 * Set-up, once per group: either (one process per GPU) every rank calls
 * accblas_peer_export, the 64-byte CUDA IPC handles are gathered in rank
 * order and passed to accblas_peer_connect_ipc; or (one process, several
 * handles) accblas_peer_mailbox + accblas_peer_connect_ptrs.  The mailbox is
 * allocated and zeroed by export / mailbox and never cleared afterwards, so
 * NO barrier is needed between connect and the first call: an entry a faster
 * peer has already published stays where it is.  A handle belongs to one
 * group at a time; accblas_peer_disconnect (after the group's last call has
 * completed on every rank) frees the mailbox, and a new group starts from a
 * fresh export.  All ranks must issue the same sequence of
 * accblas_dot_allreduce calls.
 *
 * Failure: a peer that has not published its partial after the time limit
 * (accblas_peer_set_timeout, default 30 s, 0 = wait for ever) makes THIS
 * rank's result of that call NaN and raises a sticky failure word in mapped
 * host memory.  From then on accblas_dot_allreduce on this handle returns
 * ACCBLAS_ERR_PEER without launching (the ranks no longer agree on the call
 * number); accblas_peer_status reports the failing call number at any time
 * without synchronising.  Recovery: accblas_peer_disconnect on every rank and
 * a new group, or fall back to accblas_dot + an external all-reduce. */
int accblas_peer_export(accblas_handle_t handle, void* ipc_handle_64_bytes);
int accblas_peer_connect_ipc(accblas_handle_t handle, int world, int rank,
                             const void* ipc_handles /* world x 64 bytes */);
int accblas_peer_mailbox(accblas_handle_t handle, void** device_ptr);
int accblas_peer_connect_ptrs(accblas_handle_t handle, int world, int rank,
                              void* const* mailboxes, const int* devices);
int accblas_peer_disconnect(accblas_handle_t handle);
int accblas_peer_set_timeout(accblas_handle_t handle, double seconds);
/* ACCBLAS_OK, or ACCBLAS_ERR_PEER with *failed_call = number of the
 * accblas_dot_allreduce call (counted from 1 since connect) that timed out */
int accblas_peer_status(accblas_handle_t handle,
                        unsigned long long* failed_call);
int accblas_dot_allreduce(accblas_handle_t handle, accblas_dtype ar,
                          accblas_dtype st, accblas_dtype res, int64_t n,
                          const void* x, int64_t incx, const void* y,
                          int64_t incy, void* result, accblas_stream_t stream);

/* In-place triangular solve T * x_out = x_in with T the `uplo` triangle of
 * the row-major n x n matrix A (unit or stored diagonal).  x is read and
 * written as `st`; every solved entry is rounded to `st` before later rows
 * consume it, as in the reference. */
int accblas_trsv(accblas_handle_t handle, accblas_dtype ar, accblas_dtype st,
                 int uplo, int diag, int64_t n, const void* A, int64_t lda,
                 void* x, int64_t incx, accblas_stream_t stream);

/* out[r*ld_out + c] = static_cast<dst>(in[r*ld_in + c]) on the device
 * (one round-to-nearest-even when narrowing, exact when widening). */
int accblas_convert(accblas_handle_t handle, accblas_dtype dst,
                    accblas_dtype src, int64_t rows, int64_t cols,
                    const void* in, int64_t ld_in, void* out, int64_t ld_out,
                    accblas_stream_t stream);

/* out[r*ld + c] = static_cast<dst>(u_k), k = first_draw + r*cols + c, where
 * u_k is the k-th value std::uniform_real_distribution<double>(-1,1) yields
 * from std::default_random_engine(seed) (libstdc++: minstd_rand0, two engine
 * calls per draw) -- the stream the reference's fixtures are built from
 * (cuda/gemv_benchmark.cu:81-83).  Returns ACCBLAS_ERR_DATA (after writing
 * everything) if a drawn value was not a normal number, which is where the
 * host generator would have re-drawn; this call synchronises the stream. */
int accblas_fill_uniform(accblas_handle_t handle, accblas_dtype dst,
                         int64_t rows, int64_t cols, void* out, int64_t ld,
                         uint32_t seed, uint64_t first_draw,
                         accblas_stream_t stream);

/* Reference error metric on the device: sum_i |ref_i - res_i| and sum_i |ref_i|
 * (fp64, deterministic tree) written to out2[0], out2[1] (DEVICE pointer,
 * two doubles).  cuda/utils.cuh:315-332 + cuda/gemv_benchmark.cu:219-232. */
int accblas_l1_error(accblas_handle_t handle, accblas_dtype ref_t,
                     accblas_dtype res_t, int64_t n, const void* ref,
                     int64_t inc_ref, const void* res, int64_t inc_res,
                     double* out2, accblas_stream_t stream);

/* Host-buffer entry points: same semantics, HOST pointers; the call stages
 * the operands through device memory owned by the handle (H2D, kernel, D2H)
 * on `stream` and synchronises it before returning.  These are what a host
 * application that keeps its data in host memory calls; they are also the
 * path bench.py times as `e2e`. */
int accblas_gemv_host(accblas_handle_t handle, accblas_dtype ar,
                      accblas_dtype st, int64_t m, int64_t n, double alpha,
                      const void* A, int64_t lda, const void* x, int64_t incx,
                      double beta, void* y, int64_t incy,
                      accblas_stream_t stream);
int accblas_dot_host(accblas_handle_t handle, accblas_dtype ar,
                     accblas_dtype st, accblas_dtype res, int64_t n,
                     const void* x, int64_t incx, const void* y, int64_t incy,
                     void* result, accblas_stream_t stream);
int accblas_trsv_host(accblas_handle_t handle, accblas_dtype ar,
                      accblas_dtype st, int uplo, int diag, int64_t n,
                      const void* A, int64_t lda, void* x, int64_t incx,
                      accblas_stream_t stream);

/* Development knob, not part of the drop-in surface: sets a launch-shape
 * parameter (the fields of accblas::Tuning, csrc/tuning.h, by name)
 * process-wide; unknown keys and out-of-range values are rejected.  Not
 * synchronised: set it while no other thread is inside the library.  Results
 * are bit-reproducible for a fixed configuration only. */
int accblas_tune(const char* key, int value);

#ifdef __cplusplus
}
#endif

#endif /* ACCBLAS_H_ */
