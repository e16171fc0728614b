// extern "C" boundary of libaccblas_b200.so (declared in include/accblas.h).
// Argument validation, handle/workspace management, host-buffer staging.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "common.cuh"
#include "tuning.h"

namespace accblas {

// implemented in the kernel translation units
int gemv_impl(Handle*, int ar, int st, std::int64_t m, std::int64_t n,
              double alpha, const void* A, std::int64_t lda, const void* x,
              std::int64_t incx, double beta, void* y, std::int64_t incy,
              cudaStream_t);
int dot_impl(Handle*, int ar, int st, int res, std::int64_t n, const void* x,
             std::int64_t incx, const void* y, std::int64_t incy, void* result,
             cudaStream_t, const PeerExchange* px = nullptr);
int trsv_impl(Handle*, int ar, int st, int uplo, int diag, std::int64_t n,
              const void* A, std::int64_t lda, void* x, std::int64_t incx,
              cudaStream_t, long long* trace = nullptr);
int convert_impl(Handle*, int dst, int src, std::int64_t rows,
                 std::int64_t cols, const void* in, std::int64_t ld_in,
                 void* out, std::int64_t ld_out, cudaStream_t);
int fill_uniform_impl(Handle*, int dst, std::int64_t rows, std::int64_t cols,
                      void* out, std::int64_t ld, std::uint32_t seed,
                      std::uint64_t first_draw, cudaStream_t);
int l1_error_impl(Handle*, int ref_t, int res_t, std::int64_t n,
                  const void* ref, std::int64_t inc_ref, const void* res,
                  std::int64_t inc_res, double* out2, cudaStream_t);

int set_gemv_trace(unsigned long long* ptr);

namespace {
thread_local char g_error[512] = "";
Tuning g_tuning;
}  // namespace

Tuning& tuning() { return g_tuning; }

void set_error(const char* fmt, ...)
{
    va_list args;
    va_start(args, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, args);
    va_end(args);
}

int cuda_fail(cudaError_t err, const char* what)
{
    set_error("CUDA error %d (%s) in %s", static_cast<int>(err),
              cudaGetErrorString(err), what);
    return ACCBLAS_ERR_CUDA;
}

int ensure_workspace(Handle* h, size_t bytes, cudaStream_t stream)
{
    const size_t need = kControlBytes + bytes;
    if (h->ws != nullptr && h->ws_bytes >= need) {
        return ACCBLAS_OK;
    }
    size_t cap = h->ws_bytes ? h->ws_bytes : (size_t{4} << 20);
    while (cap < need) {
        cap *= 2;
    }
    if (h->ws != nullptr) {
        // earlier kernels on this handle may still read the old block
        ACCBLAS_CUDA(cudaStreamSynchronize(stream));
        ACCBLAS_CUDA(cudaFree(h->ws));
        h->ws = nullptr;
        h->ws_bytes = 0;
    }
    void* p = nullptr;
    cudaError_t err = cudaMalloc(&p, cap);
    if (err != cudaSuccess) {
        set_error("workspace allocation of %zu bytes failed: %s", cap,
                  cudaGetErrorString(err));
        cudaGetLastError();
        return ACCBLAS_ERR_ALLOC;
    }
    h->ws = p;
    h->ws_bytes = cap;
    ACCBLAS_CUDA(cudaMemsetAsync(p, 0, kControlBytes, stream));
    return ACCBLAS_OK;
}

namespace {

int ensure_stage(Handle* h, int slot, size_t bytes)
{
    if (h->stage[slot] != nullptr && h->stage_bytes[slot] >= bytes) {
        return ACCBLAS_OK;
    }
    if (h->stage[slot] != nullptr) {
        ACCBLAS_CUDA(cudaFree(h->stage[slot]));
        h->stage[slot] = nullptr;
        h->stage_bytes[slot] = 0;
    }
    void* p = nullptr;
    cudaError_t err = cudaMalloc(&p, bytes ? bytes : 16);
    if (err != cudaSuccess) {
        set_error("staging allocation of %zu bytes failed: %s", bytes,
                  cudaGetErrorString(err));
        cudaGetLastError();
        return ACCBLAS_ERR_ALLOC;
    }
    h->stage[slot] = p;
    h->stage_bytes[slot] = bytes ? bytes : 16;
    return ACCBLAS_OK;
}

struct DeviceGuard {
    int previous = -1;
    bool switched = false;
    int enter(int device)
    {
        ACCBLAS_CUDA(cudaGetDevice(&previous));
        if (previous != device) {
            ACCBLAS_CUDA(cudaSetDevice(device));
            switched = true;
        }
        return ACCBLAS_OK;
    }
    ~DeviceGuard()
    {
        if (switched) {
            cudaSetDevice(previous);
        }
    }
};

bool check_handle(accblas_handle_t handle)
{
    if (handle == nullptr) {
        set_error("null handle");
        return false;
    }
    return true;
}

// elements spanned by a strided vector / matrix
size_t vec_span(std::int64_t n, std::int64_t inc)
{
    return n > 0 ? static_cast<size_t>((n - 1) * inc + 1) : 0;
}
size_t mat_span(std::int64_t rows, std::int64_t cols, std::int64_t ld)
{
    return rows > 0 ? static_cast<size_t>((rows - 1) * ld + cols) : 0;
}

}  // namespace
}  // namespace accblas

using accblas::Handle;

#define ACCBLAS_ENTER(handle)                                   \
    if (!accblas::check_handle(handle)) {                       \
        return ACCBLAS_ERR_INVALID;                             \
    }                                                           \
    Handle* h = reinterpret_cast<Handle*>(handle);              \
    accblas::DeviceGuard guard__;                               \
    {                                                           \
        int rc__ = guard__.enter(h->device);                    \
        if (rc__ != ACCBLAS_OK) {                               \
            return rc__;                                        \
        }                                                       \
    }                                                           \
    cudaStream_t s = static_cast<cudaStream_t>(stream)

extern "C" {

int accblas_version(void) { return ACCBLAS_VERSION; }

const char* accblas_status_string(int status)
{
    switch (status) {
    case ACCBLAS_OK:
        return "ok";
    case ACCBLAS_ERR_INVALID:
        return "invalid argument";
    case ACCBLAS_ERR_UNSUPPORTED:
        return "unsupported dtype combination";
    case ACCBLAS_ERR_CUDA:
        return "CUDA failure";
    case ACCBLAS_ERR_ALLOC:
        return "device allocation failure";
    case ACCBLAS_ERR_DATA:
        return "data-dependent failure";
    case ACCBLAS_ERR_PEER:
        return "multi-GPU exchange failure";
    default:
        return "unknown status";
    }
}

const char* accblas_last_error(void) { return accblas::g_error; }

size_t accblas_sizeof(accblas_dtype t)
{
    switch (t) {
    case ACCBLAS_F64:
        return 8;
    case ACCBLAS_F32:
        return 4;
    case ACCBLAS_F16:
        return 2;
    default:
        return 0;
    }
}

int accblas_create(accblas_handle_t* handle, int device)
{
    if (handle == nullptr) {
        accblas::set_error("null handle pointer");
        return ACCBLAS_ERR_INVALID;
    }
    *handle = nullptr;
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        cudaGetLastError();
        accblas::set_error(
            "no CUDA device available (%s); this library has no CPU path",
            err != cudaSuccess ? cudaGetErrorString(err) : "device count 0");
        return ACCBLAS_ERR_CUDA;
    }
    if (device < 0) {
        ACCBLAS_CUDA(cudaGetDevice(&device));
    }
    if (device >= count) {
        accblas::set_error("device %d out of range (count %d)", device, count);
        return ACCBLAS_ERR_INVALID;
    }
    Handle* h = new (std::nothrow) Handle();
    if (h == nullptr) {
        accblas::set_error("out of host memory");
        return ACCBLAS_ERR_ALLOC;
    }
    h->device = device;
    accblas::DeviceGuard guard;
    int rc = guard.enter(device);
    if (rc != ACCBLAS_OK) {
        delete h;
        return rc;
    }
    err = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount,
                                 device);
    if (err != cudaSuccess) {
        delete h;
        return accblas::cuda_fail(err, "cudaDeviceGetAttribute");
    }
    rc = accblas::ensure_workspace(h, 0, nullptr);
    if (rc == ACCBLAS_OK) {
        err = cudaStreamSynchronize(nullptr);
        if (err != cudaSuccess) {
            rc = accblas::cuda_fail(err, "cudaStreamSynchronize");
        }
    }
    if (rc != ACCBLAS_OK) {
        if (h->ws) {
            cudaFree(h->ws);
        }
        delete h;
        return rc;
    }
    *handle = reinterpret_cast<accblas_handle_t>(h);
    return ACCBLAS_OK;
}

int accblas_destroy(accblas_handle_t handle)
{
    if (handle == nullptr) {
        return ACCBLAS_OK;
    }
    Handle* h = reinterpret_cast<Handle*>(handle);
    accblas::DeviceGuard guard;
    guard.enter(h->device);
    cudaDeviceSynchronize();
    if (h->ws) {
        cudaFree(h->ws);
    }
    for (int i = 0; i < 4; ++i) {
        if (h->stage[i]) {
            cudaFree(h->stage[i]);
        }
    }
    for (int i = 0; i < accblas::kMaxPeers; ++i) {
        if (h->peer_mailbox[i] && h->peer_is_ipc[i]) {
            cudaIpcCloseMemHandle(h->peer_mailbox[i]);
        }
    }
    if (h->mailbox) {
        cudaFree(h->mailbox);
    }
    if (h->peer_status_host) {
        cudaFreeHost(h->peer_status_host);
    }
    delete h;
    return ACCBLAS_OK;
}

int accblas_get_sm_count(accblas_handle_t handle, int* sm_count)
{
    if (!accblas::check_handle(handle) || sm_count == nullptr) {
        return ACCBLAS_ERR_INVALID;
    }
    *sm_count = reinterpret_cast<Handle*>(handle)->sm_count;
    return ACCBLAS_OK;
}

int accblas_gemv(accblas_handle_t handle, accblas_dtype ar, accblas_dtype st,
                 int64_t m, int64_t n, double alpha, const void* A,
                 int64_t lda, const void* x, int64_t incx, double beta,
                 void* y, int64_t incy, accblas_stream_t stream)
{
    ACCBLAS_ENTER(handle);
    if (m < 0 || n < 0 || lda < n || incx < 1 || incy < 1) {
        accblas::set_error(
            "gemv: bad shape (m=%lld n=%lld lda=%lld incx=%lld incy=%lld)",
            (long long)m, (long long)n, (long long)lda, (long long)incx,
            (long long)incy);
        return ACCBLAS_ERR_INVALID;
    }
    if (m > 0 && (y == nullptr || (n > 0 && (A == nullptr || x == nullptr)))) {
        accblas::set_error("gemv: null operand");
        return ACCBLAS_ERR_INVALID;
    }
    return accblas::gemv_impl(h, ar, st, m, n, alpha, A, lda, x, incx, beta, y,
                              incy, s);
}

int accblas_dot(accblas_handle_t handle, accblas_dtype ar, accblas_dtype st,
                accblas_dtype res, int64_t n, const void* x, int64_t incx,
                const void* y, int64_t incy, void* result,
                accblas_stream_t stream)
{
    ACCBLAS_ENTER(handle);
    if (n < 0 || incx < 1 || incy < 1 || !accblas::valid_dtype(res)) {
        accblas::set_error("dot: bad shape (n=%lld incx=%lld incy=%lld res=%d)",
                           (long long)n, (long long)incx, (long long)incy,
                           (int)res);
        return ACCBLAS_ERR_INVALID;
    }
    if (result == nullptr || (n > 0 && (x == nullptr || y == nullptr))) {
        accblas::set_error("dot: null operand");
        return ACCBLAS_ERR_INVALID;
    }
    return accblas::dot_impl(h, ar, st, res, n, x, incx, y, incy, result, s);
}

// ---- multi-GPU DOT: partials combined inside the kernel over peer memory ----
// Life cycle of the exchange state of a handle:
//   export / mailbox   allocates the mailbox ZEROED (once); from here on peers
//                      may map it and publish into it at any time
//   connect_*          maps the peers; does NOT touch this rank's mailbox, so
//                      an entry a faster peer has already published survives
//                      (no barrier is needed between connect and the first
//                      call); epochs count from 1
//   disconnect         unmaps the peers and FREES the mailbox: a new group
//                      starts from a fresh export, never from reused slots
static int ensure_mailbox(Handle* h)
{
    if (h->mailbox == nullptr) {
        // plain cudaMalloc: the allocation must be exportable with CUDA IPC
        ACCBLAS_CUDA(cudaMalloc(&h->mailbox, accblas::kMailboxBytes));
        ACCBLAS_CUDA(cudaMemset(h->mailbox, 0, accblas::kMailboxBytes));
        ACCBLAS_CUDA(cudaDeviceSynchronize());
    }
    if (h->peer_status_host == nullptr) {
        void* host = nullptr;
        ACCBLAS_CUDA(cudaHostAlloc(&host, sizeof(unsigned long long),
                                   cudaHostAllocMapped));
        h->peer_status_host = static_cast<unsigned long long*>(host);
        *h->peer_status_host = 0ull;
        void* dev = nullptr;
        ACCBLAS_CUDA(cudaHostGetDevicePointer(&dev, host, 0));
        h->peer_status_dev = static_cast<unsigned long long*>(dev);
    }
    return ACCBLAS_OK;
}

static void peer_unmap(Handle* h)
{
    for (int i = 0; i < accblas::kMaxPeers; ++i) {
        if (h->peer_mailbox[i] && h->peer_is_ipc[i]) {
            cudaIpcCloseMemHandle(h->peer_mailbox[i]);
        }
        h->peer_mailbox[i] = nullptr;
        h->peer_is_ipc[i] = false;
    }
    h->peer_world = 0;
    h->peer_rank = 0;
    h->peer_epoch = 0;
}

int accblas_peer_mailbox(accblas_handle_t handle, void** device_ptr)
{
    accblas_stream_t stream = nullptr;
    ACCBLAS_ENTER(handle);
    (void)s;
    if (device_ptr == nullptr) {
        return ACCBLAS_ERR_INVALID;
    }
    int rc = ensure_mailbox(h);
    if (rc != ACCBLAS_OK) {
        return rc;
    }
    *device_ptr = h->mailbox;
    return ACCBLAS_OK;
}

int accblas_peer_export(accblas_handle_t handle, void* ipc_handle_64_bytes)
{
    accblas_stream_t stream = nullptr;
    ACCBLAS_ENTER(handle);
    (void)s;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (ipc_handle_64_bytes == nullptr) {
        return ACCBLAS_ERR_INVALID;
    }
    int rc = ensure_mailbox(h);
    if (rc != ACCBLAS_OK) {
        return rc;
    }
    cudaIpcMemHandle_t ipc;
    ACCBLAS_CUDA(cudaIpcGetMemHandle(&ipc, h->mailbox));
    memcpy(ipc_handle_64_bytes, &ipc, sizeof(ipc));
    return ACCBLAS_OK;
}

static int peer_begin(Handle* h, int world, int rank)
{
    if (world < 1 || world > accblas::kMaxPeers || rank < 0 || rank >= world) {
        accblas::set_error("peer connect: world=%d rank=%d (at most %d peers)",
                           world, rank, accblas::kMaxPeers);
        return ACCBLAS_ERR_INVALID;
    }
    if (h->peer_world != 0) {
        accblas::set_error(
            "peer connect: handle already belongs to a group of %d; call "
            "accblas_peer_disconnect and export again",
            h->peer_world);
        return ACCBLAS_ERR_INVALID;
    }
    int rc = ensure_mailbox(h);
    if (rc != ACCBLAS_OK) {
        return rc;
    }
    h->peer_world = world;
    h->peer_rank = rank;
    h->peer_epoch = 0;
    h->peer_mailbox[rank] = h->mailbox;
    return ACCBLAS_OK;
}

int accblas_peer_connect_ipc(accblas_handle_t handle, int world, int rank,
                             const void* ipc_handles)
{
    accblas_stream_t stream = nullptr;
    ACCBLAS_ENTER(handle);
    (void)s;
    if (ipc_handles == nullptr) {
        return ACCBLAS_ERR_INVALID;
    }
    int rc = peer_begin(h, world, rank);
    if (rc != ACCBLAS_OK) {
        return rc;
    }
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            continue;
        }
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, static_cast<const char*>(ipc_handles) + 64 * r, 64);
        void* mapped = nullptr;
        cudaError_t err = cudaIpcOpenMemHandle(&mapped, ipc,
                                               cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess) {
            peer_unmap(h);
            return accblas::cuda_fail(err, "cudaIpcOpenMemHandle");
        }
        h->peer_mailbox[r] = mapped;
        h->peer_is_ipc[r] = true;
    }
    return ACCBLAS_OK;
}

int accblas_peer_connect_ptrs(accblas_handle_t handle, int world, int rank,
                              void* const* mailboxes, const int* devices)
{
    accblas_stream_t stream = nullptr;
    ACCBLAS_ENTER(handle);
    (void)s;
    if (mailboxes == nullptr || devices == nullptr) {
        return ACCBLAS_ERR_INVALID;
    }
    int rc = peer_begin(h, world, rank);
    if (rc != ACCBLAS_OK) {
        return rc;
    }
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            continue;
        }
        if (devices[r] != h->device) {
            cudaError_t err = cudaDeviceEnablePeerAccess(devices[r], 0);
            if (err == cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
            } else if (err != cudaSuccess) {
                peer_unmap(h);
                return accblas::cuda_fail(err, "cudaDeviceEnablePeerAccess");
            }
        }
        h->peer_mailbox[r] = mailboxes[r];
    }
    return ACCBLAS_OK;
}

int accblas_peer_disconnect(accblas_handle_t handle)
{
    accblas_stream_t stream = nullptr;
    ACCBLAS_ENTER(handle);
    (void)s;
    ACCBLAS_CUDA(cudaDeviceSynchronize());
    peer_unmap(h);
    if (h->mailbox) {
        ACCBLAS_CUDA(cudaFree(h->mailbox));
        h->mailbox = nullptr;
    }
    if (h->peer_status_host) {
        *h->peer_status_host = 0ull;
    }
    return ACCBLAS_OK;
}

int accblas_peer_set_timeout(accblas_handle_t handle, double seconds)
{
    if (!accblas::check_handle(handle) || !(seconds >= 0.0) ||
        seconds > 86400.0) {
        accblas::set_error("peer timeout: 0 (wait for ever) ... 86400 s");
        return ACCBLAS_ERR_INVALID;
    }
    reinterpret_cast<Handle*>(handle)->peer_timeout_ns =
        static_cast<unsigned long long>(seconds * 1e9);
    return ACCBLAS_OK;
}

int accblas_peer_status(accblas_handle_t handle,
                        unsigned long long* failed_call)
{
    if (!accblas::check_handle(handle)) {
        return ACCBLAS_ERR_INVALID;
    }
    Handle* h = reinterpret_cast<Handle*>(handle);
    const unsigned long long failed =
        h->peer_status_host
            ? *reinterpret_cast<volatile unsigned long long*>(
                  h->peer_status_host)
            : 0ull;
    if (failed_call != nullptr) {
        *failed_call = failed;
    }
    if (failed != 0ull) {
        accblas::set_error(
            "dot_allreduce call #%llu: a peer did not publish its partial "
            "within %.1f s; this rank's result of that call is NaN and the "
            "group is out of step (disconnect and form a new group)",
            failed, h->peer_timeout_ns * 1e-9);
        return ACCBLAS_ERR_PEER;
    }
    return ACCBLAS_OK;
}

int accblas_dot_allreduce(accblas_handle_t handle, accblas_dtype ar,
                          accblas_dtype st, accblas_dtype res, int64_t n,
                          const void* x, int64_t incx, const void* y,
                          int64_t incy, void* result, accblas_stream_t stream)
{
    ACCBLAS_ENTER(handle);
    if (n < 0 || incx < 1 || incy < 1 || !accblas::valid_dtype(res)) {
        accblas::set_error("dot_allreduce: bad shape");
        return ACCBLAS_ERR_INVALID;
    }
    if (result == nullptr || (n > 0 && (x == nullptr || y == nullptr))) {
        accblas::set_error("dot_allreduce: null operand");
        return ACCBLAS_ERR_INVALID;
    }
    if (h->peer_world < 1) {
        accblas::set_error("dot_allreduce: call accblas_peer_connect_* first");
        return ACCBLAS_ERR_INVALID;
    }
    // an earlier call on this handle timed out: the ranks no longer agree on
    // the epoch, nothing is launched (sticky until accblas_peer_disconnect)
    int rc = accblas_peer_status(handle, nullptr);
    if (rc != ACCBLAS_OK) {
        return rc;
    }
    accblas::PeerExchange px;
    px.world = h->peer_world;
    px.rank = h->peer_rank;
    px.epoch = ++h->peer_epoch;
    px.timeout_ns = h->peer_timeout_ns;
    px.status = h->peer_status_dev;
    for (int r = 0; r < px.world; ++r) {
        px.mailbox[r] = h->peer_mailbox[r];
    }
    return accblas::dot_impl(h, ar, st, res, n, x, incx, y, incy, result, s,
                             &px);
}

int accblas_trsv(accblas_handle_t handle, accblas_dtype ar, accblas_dtype st,
                 int uplo, int diag, int64_t n, const void* A, int64_t lda,
                 void* x, int64_t incx, accblas_stream_t stream)
{
    ACCBLAS_ENTER(handle);
    if (n < 0 || lda < n || incx < 1 ||
        (uplo != ACCBLAS_UPPER && uplo != ACCBLAS_LOWER) ||
        (diag != ACCBLAS_UNIT && diag != ACCBLAS_NON_UNIT)) {
        accblas::set_error(
            "trsv: bad argument (n=%lld lda=%lld incx=%lld uplo=%d diag=%d)",
            (long long)n, (long long)lda, (long long)incx, uplo, diag);
        return ACCBLAS_ERR_INVALID;
    }
    if (n > 0 && (A == nullptr || x == nullptr)) {
        accblas::set_error("trsv: null operand");
        return ACCBLAS_ERR_INVALID;
    }
    return accblas::trsv_impl(h, ar, st, uplo, diag, n, A, lda, x, incx, s);
}

int accblas_convert(accblas_handle_t handle, accblas_dtype dst,
                    accblas_dtype src, int64_t rows, int64_t cols,
                    const void* in, int64_t ld_in, void* out, int64_t ld_out,
                    accblas_stream_t stream)
{
    ACCBLAS_ENTER(handle);
    if (rows < 0 || cols < 0 || ld_in < cols || ld_out < cols) {
        accblas::set_error("convert: bad shape");
        return ACCBLAS_ERR_INVALID;
    }
    if (rows > 0 && cols > 0 && (in == nullptr || out == nullptr)) {
        accblas::set_error("convert: null operand");
        return ACCBLAS_ERR_INVALID;
    }
    return accblas::convert_impl(h, dst, src, rows, cols, in, ld_in, out,
                                 ld_out, s);
}

int accblas_fill_uniform(accblas_handle_t handle, accblas_dtype dst,
                         int64_t rows, int64_t cols, void* out, int64_t ld,
                         uint32_t seed, uint64_t first_draw,
                         accblas_stream_t stream)
{
    ACCBLAS_ENTER(handle);
    if (rows < 0 || cols < 0 || ld < cols) {
        accblas::set_error("fill_uniform: bad shape");
        return ACCBLAS_ERR_INVALID;
    }
    if (rows > 0 && cols > 0 && out == nullptr) {
        accblas::set_error("fill_uniform: null operand");
        return ACCBLAS_ERR_INVALID;
    }
    return accblas::fill_uniform_impl(h, dst, rows, cols, out, ld, seed,
                                      first_draw, s);
}

int accblas_l1_error(accblas_handle_t handle, accblas_dtype ref_t,
                     accblas_dtype res_t, int64_t n, const void* ref,
                     int64_t inc_ref, const void* res, int64_t inc_res,
                     double* out2, accblas_stream_t stream)
{
    ACCBLAS_ENTER(handle);
    if (n < 0 || inc_ref < 1 || inc_res < 1 || out2 == nullptr ||
        (n > 0 && (ref == nullptr || res == nullptr))) {
        accblas::set_error("l1_error: bad argument");
        return ACCBLAS_ERR_INVALID;
    }
    return accblas::l1_error_impl(h, ref_t, res_t, n, ref, inc_ref, res,
                                  inc_res, out2, s);
}

// ---------------------------------------------------------------------------
// host-buffer entry points
// ---------------------------------------------------------------------------
int accblas_gemv_host(accblas_handle_t handle, accblas_dtype ar,
                      accblas_dtype st, int64_t m, int64_t n, double alpha,
                      const void* A, int64_t lda, const void* x, int64_t incx,
                      double beta, void* y, int64_t incy,
                      accblas_stream_t stream)
{
    ACCBLAS_ENTER(handle);
    const size_t es = accblas_sizeof(st);
    if (es == 0 || m < 0 || n < 0 || lda < n || incx < 1 || incy < 1) {
        accblas::set_error("gemv_host: bad argument");
        return ACCBLAS_ERR_INVALID;
    }
    if (m == 0) {
        return ACCBLAS_OK;
    }
    const size_t a_bytes = accblas::mat_span(m, n, lda) * es;
    const size_t x_bytes = accblas::vec_span(n, incx) * es;
    const size_t y_bytes = accblas::vec_span(m, incy) * es;
    int rc;
    if ((rc = accblas::ensure_stage(h, 0, a_bytes)) != ACCBLAS_OK ||
        (rc = accblas::ensure_stage(h, 1, x_bytes)) != ACCBLAS_OK ||
        (rc = accblas::ensure_stage(h, 2, y_bytes)) != ACCBLAS_OK) {
        return rc;
    }
    if (a_bytes) {
        ACCBLAS_CUDA(cudaMemcpyAsync(h->stage[0], A, a_bytes,
                                     cudaMemcpyHostToDevice, s));
    }
    if (x_bytes) {
        ACCBLAS_CUDA(cudaMemcpyAsync(h->stage[1], x, x_bytes,
                                     cudaMemcpyHostToDevice, s));
    }
    if (beta != 0.0 || incy != 1) {
        ACCBLAS_CUDA(cudaMemcpyAsync(h->stage[2], y, y_bytes,
                                     cudaMemcpyHostToDevice, s));
    }
    rc = accblas::gemv_impl(h, ar, st, m, n, alpha, h->stage[0], lda,
                            h->stage[1], incx, beta, h->stage[2], incy, s);
    if (rc != ACCBLAS_OK) {
        return rc;
    }
    ACCBLAS_CUDA(cudaMemcpyAsync(y, h->stage[2], y_bytes,
                                 cudaMemcpyDeviceToHost, s));
    ACCBLAS_CUDA(cudaStreamSynchronize(s));
    return ACCBLAS_OK;
}

int accblas_dot_host(accblas_handle_t handle, accblas_dtype ar,
                     accblas_dtype st, accblas_dtype res, int64_t n,
                     const void* x, int64_t incx, const void* y, int64_t incy,
                     void* result, accblas_stream_t stream)
{
    ACCBLAS_ENTER(handle);
    const size_t es = accblas_sizeof(st);
    const size_t rs = accblas_sizeof(res);
    if (es == 0 || rs == 0 || n < 0 || incx < 1 || incy < 1 ||
        result == nullptr) {
        accblas::set_error("dot_host: bad argument");
        return ACCBLAS_ERR_INVALID;
    }
    const size_t x_bytes = accblas::vec_span(n, incx) * es;
    const size_t y_bytes = accblas::vec_span(n, incy) * es;
    int rc;
    if ((rc = accblas::ensure_stage(h, 0, x_bytes)) != ACCBLAS_OK ||
        (rc = accblas::ensure_stage(h, 1, y_bytes)) != ACCBLAS_OK ||
        (rc = accblas::ensure_stage(h, 2, 16)) != ACCBLAS_OK) {
        return rc;
    }
    if (x_bytes) {
        ACCBLAS_CUDA(cudaMemcpyAsync(h->stage[0], x, x_bytes,
                                     cudaMemcpyHostToDevice, s));
        ACCBLAS_CUDA(cudaMemcpyAsync(h->stage[1], y, y_bytes,
                                     cudaMemcpyHostToDevice, s));
    }
    rc = accblas::dot_impl(h, ar, st, res, n, h->stage[0], incx, h->stage[1],
                           incy, h->stage[2], s);
    if (rc != ACCBLAS_OK) {
        return rc;
    }
    ACCBLAS_CUDA(cudaMemcpyAsync(result, h->stage[2], rs,
                                 cudaMemcpyDeviceToHost, s));
    ACCBLAS_CUDA(cudaStreamSynchronize(s));
    return ACCBLAS_OK;
}

int accblas_trsv_host(accblas_handle_t handle, accblas_dtype ar,
                      accblas_dtype st, int uplo, int diag, int64_t n,
                      const void* A, int64_t lda, void* x, int64_t incx,
                      accblas_stream_t stream)
{
    ACCBLAS_ENTER(handle);
    const size_t es = accblas_sizeof(st);
    if (es == 0 || n < 0 || lda < n || incx < 1) {
        accblas::set_error("trsv_host: bad argument");
        return ACCBLAS_ERR_INVALID;
    }
    if (n == 0) {
        return ACCBLAS_OK;
    }
    const size_t a_bytes = accblas::mat_span(n, n, lda) * es;
    const size_t x_bytes = accblas::vec_span(n, incx) * es;
    int rc;
    if ((rc = accblas::ensure_stage(h, 0, a_bytes)) != ACCBLAS_OK ||
        (rc = accblas::ensure_stage(h, 1, x_bytes)) != ACCBLAS_OK) {
        return rc;
    }
    ACCBLAS_CUDA(cudaMemcpyAsync(h->stage[0], A, a_bytes,
                                 cudaMemcpyHostToDevice, s));
    ACCBLAS_CUDA(cudaMemcpyAsync(h->stage[1], x, x_bytes,
                                 cudaMemcpyHostToDevice, s));
    rc = accblas::trsv_impl(h, ar, st, uplo, diag, n, h->stage[0], lda,
                            h->stage[1], incx, s);
    if (rc != ACCBLAS_OK) {
        return rc;
    }
    ACCBLAS_CUDA(cudaMemcpyAsync(x, h->stage[1], x_bytes,
                                 cudaMemcpyDeviceToHost, s));
    ACCBLAS_CUDA(cudaStreamSynchronize(s));
    return ACCBLAS_OK;
}

#if defined(ACCBLAS_DEV_HOOKS)
// Development aids, compiled only into libaccblas_b200_dev.so (build.py
// build_dev; never into the product library).
// accblas_trsv that also records phase timestamps per block row into `trace`
// (device pointer, ceil(n/128) * 64 long longs: 64 slots per block row).
int accblas_dev_trsv_trace(accblas_handle_t handle, int ar, int st, int uplo,
                           int diag, int64_t n, const void* A, int64_t lda,
                           void* x, int64_t incx, long long* trace,
                           accblas_stream_t stream)
{
    ACCBLAS_ENTER(handle);
    if (n < 0 || lda < n || incx < 1 || trace == nullptr ||
        (n > 0 && (A == nullptr || x == nullptr))) {
        accblas::set_error("trsv trace: bad argument");
        return ACCBLAS_ERR_INVALID;
    }
    return accblas::trsv_impl(h, ar, st, uplo, diag, n, A, lda, x, incx, s,
                              trace);
}

// per-CTA start/end timestamps of the next GEMV launches (3 x grid unsigned
// long longs, device pointer; nullptr switches it off)
int accblas_dev_gemv_trace(unsigned long long* trace)
{
    return accblas::set_gemv_trace(trace);
}
#endif  // ACCBLAS_DEV_HOOKS

// Development knob (not part of the drop-in surface): set a launch-shape
// parameter by name.  Unknown keys and out-of-range values are rejected.
int accblas_tune(const char* key, int value)
{
    using accblas::Tuning;
    struct Knob {
        const char* name;
        int Tuning::*field;
        int lo, hi;
    };
    static const Knob knobs[] = {
        {"dot_unroll", &Tuning::dot_unroll, 0, 4},
        {"dot_block", &Tuning::dot_block, 0, 1024},
        {"dot_ctas_per_sm", &Tuning::dot_ctas_per_sm, 0, 32},
        {"dot_pdl", &Tuning::dot_pdl, 0, 1},
        {"dot_intmix", &Tuning::dot_intmix, 0, 1},
        {"dot_pool_pct", &Tuning::dot_pool_pct, 0, 100},
        {"dot_chunk_tiles", &Tuning::dot_chunk_tiles, 1, 4096},
        {"gemv_unroll", &Tuning::gemv_unroll, 0, 4},
        {"gemv_variant", &Tuning::gemv_variant, 0, 5},
        {"gemv_ctas_per_sm", &Tuning::gemv_ctas_per_sm, 0, 32},
        {"gemv_pipe", &Tuning::gemv_pipe, -1, 4},
        {"gemv_intwords", &Tuning::gemv_intwords, 0, 4},
        {"gemv_pdl", &Tuning::gemv_pdl, 0, 1},
        {"gemv_force_pieces", &Tuning::gemv_force_pieces, -1, 8},
        {"gemv_taper", &Tuning::gemv_taper, 0, 1},
        {"gemv_stages", &Tuning::gemv_stages, 0, 4},
        {"trsv_variant", &Tuning::trsv_variant, -1, 1},
        {"trsv_whole_block_spin", &Tuning::trsv_whole_block_spin, 0, 1},
        {"trsv_l2_ahead", &Tuning::trsv_l2_ahead, -1, 1 << 20},
        {"fill_generic", &Tuning::fill_generic, 0, 1},
    };
    if (key == nullptr) {
        return ACCBLAS_ERR_INVALID;
    }
    for (const Knob& k : knobs) {
        if (strcmp(key, k.name) == 0) {
            if (value < k.lo || value > k.hi) {
                accblas::set_error("tuning key '%s': %d outside [%d, %d]", key,
                                   value, k.lo, k.hi);
                return ACCBLAS_ERR_INVALID;
            }
            accblas::tuning().*(k.field) = value;
            return ACCBLAS_OK;
        }
    }
    accblas::set_error("unknown tuning key '%s'", key);
    return ACCBLAS_ERR_INVALID;
}

}  // extern "C"
