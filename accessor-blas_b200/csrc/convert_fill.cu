// Device-side fixture kernels:
//   * storage conversion  out = static_cast<Dst>(in)   (host loop in the
//     reference: /root/reference/cuda/matrix_helper.cuh:93-103 and the
//     convert() members of cuda/{gemv,dot,trsv}_memory.cuh)
//   * uniform(-1,1) generation reproducing, draw for draw, what
//     std::uniform_real_distribution<double>(-1,1) yields from
//     std::default_random_engine(seed) in libstdc++ (host loop in the
//     reference: cuda/matrix_helper.cuh:28-75, seeded at
//     cuda/gemv_benchmark.cu:81-83).  The engine is minstd_rand0
//     (s' = 16807 s mod 2^31-1), so draw k can be computed independently by
//     modular exponentiation -- no sequential dependency, any number of GPUs
//     can fill their own slab of the same global stream.
//   * the reference's L1 error metric (cuda/utils.cuh:315-332).
#include "common.cuh"
#include "tuning.h"

namespace accblas {
namespace {

constexpr int kElemsPerThread = 4;

template <typename Dst, typename Src>
__device__ __forceinline__ Dst cast_one(Src v)
{
    return static_cast<Dst>(v);
}
template <>
__device__ __forceinline__ __half cast_one<__half, double>(double v)
{
    return __double2half(v);  // one rounding, NOT double -> float -> half
}
template <>
__device__ __forceinline__ __half cast_one<__half, float>(float v)
{
    return __float2half_rn(v);
}
template <>
__device__ __forceinline__ __half cast_one<__half, __half>(__half v)
{
    return v;
}
template <>
__device__ __forceinline__ float cast_one<float, __half>(__half v)
{
    return __half2float(v);
}
template <>
__device__ __forceinline__ double cast_one<double, __half>(__half v)
{
    return static_cast<double>(__half2float(v));
}

template <typename T, int N>
struct alignas(sizeof(T) * N) Pack {
    T v[N];
};

// grid.x tiles the columns (kElemsPerThread per thread), grid.y strides rows.
template <typename Dst, typename Src, bool VECTOR>
__global__ __launch_bounds__(256) void convert_kernel(
    std::int64_t rows, std::int64_t cols, const Src* __restrict__ in,
    std::int64_t ld_in, Dst* __restrict__ out, std::int64_t ld_out)
{
    const std::int64_t c0 =
        (std::int64_t{blockIdx.x} * blockDim.x + threadIdx.x) * kElemsPerThread;
    if (c0 >= cols) {
        return;
    }
    for (std::int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const Src* src = in + r * ld_in + c0;
        Dst* dst = out + r * ld_out + c0;
        if (VECTOR && c0 + kElemsPerThread <= cols) {
            const auto p =
                *reinterpret_cast<const Pack<Src, kElemsPerThread>*>(src);
            Pack<Dst, kElemsPerThread> q;
#pragma unroll
            for (int i = 0; i < kElemsPerThread; ++i) {
                q.v[i] = cast_one<Dst, Src>(p.v[i]);
            }
            *reinterpret_cast<Pack<Dst, kElemsPerThread>*>(dst) = q;
        } else {
            for (int i = 0; i < kElemsPerThread && c0 + i < cols; ++i) {
                dst[i] = cast_one<Dst, Src>(src[i]);
            }
        }
    }
}

template <typename Dst, typename Src>
int launch_convert(std::int64_t rows, std::int64_t cols, const void* in,
                   std::int64_t ld_in, void* out, std::int64_t ld_out,
                   cudaStream_t stream)
{
    if (rows == 0 || cols == 0) {
        return ACCBLAS_OK;
    }
    // a contiguous matrix is one long row
    if (ld_in == cols && ld_out == cols) {
        cols = rows * cols;
        rows = 1;
        ld_in = ld_out = cols;
    }
    const bool vec =
        reinterpret_cast<std::uintptr_t>(in) %
                (sizeof(Src) * kElemsPerThread) == 0 &&
        reinterpret_cast<std::uintptr_t>(out) %
                (sizeof(Dst) * kElemsPerThread) == 0 &&
        (rows == 1 || (ld_in % kElemsPerThread == 0 &&
                       ld_out % kElemsPerThread == 0));
    const std::int64_t per_block = 256 * kElemsPerThread;
    const std::int64_t gx = (cols + per_block - 1) / per_block;
    const std::int64_t gy = rows < 32768 ? rows : 32768;
    if (gx > 0x7fffffffLL) {
        set_error("convert: row too long");
        return ACCBLAS_ERR_INVALID;
    }
    const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(gy));
    if (vec) {
        convert_kernel<Dst, Src, true><<<grid, 256, 0, stream>>>(
            rows, cols, static_cast<const Src*>(in), ld_in,
            static_cast<Dst*>(out), ld_out);
    } else {
        convert_kernel<Dst, Src, false><<<grid, 256, 0, stream>>>(
            rows, cols, static_cast<const Src*>(in), ld_in,
            static_cast<Dst*>(out), ld_out);
    }
    ACCBLAS_CUDA(cudaGetLastError());
    return ACCBLAS_OK;
}

// ---------------------------------------------------------------------------
// minstd_rand0 closed form
// ---------------------------------------------------------------------------
constexpr std::uint64_t kLcgA = 16807;
constexpr std::uint64_t kLcgM = 2147483647;  // 2^31 - 1

__host__ __device__ __forceinline__ std::uint64_t mod_m31(std::uint64_t v)
{
    // v < 2^62: two folds bring it below 2^31 + 1
    v = (v & kLcgM) + (v >> 31);
    v = (v & kLcgM) + (v >> 31);
    return v >= kLcgM ? v - kLcgM : v;
}

__host__ __device__ __forceinline__ std::uint64_t pow_a(std::uint64_t e)
{
    std::uint64_t base = kLcgA, r = 1;
    while (e) {
        if (e & 1) {
            r = mod_m31(r * base);
        }
        base = mod_m31(base * base);
        e >>= 1;
    }
    return r;
}

// value of draw number `k` given the engine state right before it
__device__ __forceinline__ double draw_uniform(std::uint64_t& state)
{
    // generate_canonical<double, 53>: two engine calls, range R = 2^31 - 2
    const double R = 2147483646.0;
    const double RR = 4611686009837453316.0;  // R*R rounded to double
    state = mod_m31(state * kLcgA);
    const double lo = static_cast<double>(state - 1);
    state = mod_m31(state * kLcgA);
    const double hi = static_cast<double>(state - 1);
    // every operation rounded on its own: the host code has no FMA
    const double sum = __dadd_rn(lo, __dmul_rn(hi, R));
    double u = __ddiv_rn(sum, RR);
    if (u >= 1.0) {
        u = __longlong_as_double(0x3FEFFFFFFFFFFFFFLL);  // nextafter(1, 0)
    }
    return __dadd_rn(__dmul_rn(u, 2.0), -1.0);
}

// ---- streaming generator ---------------------------------------------------
// Both factors of every product are below 2^31, so the 62-bit product is ONE
// 32x32->64 multiply (IMAD.WIDE.U32) and the two folds are 32-bit adds.
__host__ __device__ __forceinline__ std::uint32_t mulmod31(std::uint32_t a,
                                                           std::uint32_t b)
{
    const std::uint64_t v = std::uint64_t{a} * b;
    const std::uint32_t m = static_cast<std::uint32_t>(kLcgM);
    std::uint32_t t = (static_cast<std::uint32_t>(v) & m) +
                      static_cast<std::uint32_t>(v >> 31);  // < 2^32
    t = (t & m) + (t >> 31);                                // <= m + 1
    return t >= m ? t - m : t;
}

// 16807^e mod (2^31 - 1); the exponent is first reduced modulo the group
// order 2^31 - 2 (the modulus is prime), which bounds the loop at 31 turns
__host__ __device__ __forceinline__ std::uint32_t pow_a31(std::uint64_t e)
{
    e %= (kLcgM - 1);
    std::uint32_t base = static_cast<std::uint32_t>(kLcgA), r = 1;
    while (e) {
        if (e & 1) {
            r = mulmod31(r, base);
        }
        base = mulmod31(base, base);
        e >>= 1;
    }
    return r;
}

// Same value as draw_uniform, with the quotient sum / RR formed as
//   q = RN(sum * y), r = sum - q * RR (exact in an FMA), u = RN(q + r * y),
// y = RN(1 / RR): the classic correctly-rounded division by a constant
// (Markstein; y is the correctly rounded reciprocal, q is within one ulp of
// the quotient, so the corrected u IS the rounded quotient) -- three FP64
// instructions instead of the ~25 of the general division routine.
// tests/test_gpu_parity.py::test_fill_fast_equals_generic compares 2^27 draws
// of this path with the __ddiv_rn one bit for bit.
__device__ __forceinline__ double draw_uniform_fast(std::uint32_t& state)
{
    const double R = 2147483646.0;
    const double RR = 4611686009837453316.0;
    constexpr double kInvRR = 1.0 / 4611686009837453316.0;  // RN at compile time
    state = mulmod31(state, static_cast<std::uint32_t>(kLcgA));
    const double lo = static_cast<double>(static_cast<int>(state - 1));
    state = mulmod31(state, static_cast<std::uint32_t>(kLcgA));
    const double hi = static_cast<double>(static_cast<int>(state - 1));
    const double sum = __dadd_rn(lo, __dmul_rn(hi, R));
    const double q = __dmul_rn(sum, kInvRR);
    const double rem = __fma_rn(-q, RR, sum);
    double u = __fma_rn(rem, kInvRR, q);
    if (u >= 1.0) {
        u = __longlong_as_double(0x3FEFFFFFFFFFFFFFLL);
    }
    // 2u is exact, so this FMA rounds exactly like (u * 2.0) + (-1.0)
    return __fma_rn(u, 2.0, -1.0);
}

// Contiguous output: thread t owns the 4-draw groups t, t + T, t + 2T, ...
// (T = threads in the grid).  ONE modular exponentiation per thread positions
// the engine at its first group; after the 8 engine calls of a group the
// state is moved to the thread's next group with a single multiplication by
// the launch constant 16807^(8(T-1)).  Stores are 16-byte vectors, fully
// coalesced.  (The per-row kernel below redoes the exponentiation for every 4
// draws: 258 GB/s on B200; it remains for strided / unaligned outputs.)
template <typename Dst>
__global__ __launch_bounds__(256) void fill_linear_kernel(
    std::int64_t count, Dst* __restrict__ out, std::uint32_t seed_state,
    std::uint64_t first_draw, std::uint32_t jump,
    unsigned* __restrict__ bad_counter)
{
    const std::int64_t threads = std::int64_t{gridDim.x} * 256;
    std::int64_t p =
        (std::int64_t{blockIdx.x} * 256 + threadIdx.x) * kElemsPerThread;
    if (p >= count) {
        return;
    }
    std::uint32_t state = mulmod31(
        pow_a31(2 * (first_draw + static_cast<std::uint64_t>(p))), seed_state);
    unsigned bad = 0;
    for (; p < count; p += threads * kElemsPerThread) {
        Pack<Dst, kElemsPerThread> q;
        double v[kElemsPerThread];
#pragma unroll
        for (int i = 0; i < kElemsPerThread; ++i) {
            v[i] = draw_uniform_fast(state);
            q.v[i] = cast_one<Dst, double>(v[i]);
        }
        const int valid = (count - p >= kElemsPerThread)
                              ? kElemsPerThread
                              : static_cast<int>(count - p);
#pragma unroll
        for (int i = 0; i < kElemsPerThread; ++i) {
            const double av = fabs(v[i]);
            bad += (i >= valid || (av >= 2.2250738585072014e-308 &&
                                   av <= 1.7976931348623157e308))
                       ? 0u
                       : 1u;
        }
        if (valid == kElemsPerThread) {
            *reinterpret_cast<Pack<Dst, kElemsPerThread>*>(out + p) = q;
        } else {
#pragma unroll
            for (int i = 0; i < kElemsPerThread; ++i) {
                if (i < valid) {
                    out[p + i] = q.v[i];
                }
            }
        }
        state = mulmod31(state, jump);
    }
    if (bad) {
        atomicAdd(bad_counter, bad);
    }
}

template <typename Dst>
__global__ __launch_bounds__(256) void fill_uniform_kernel(
    std::int64_t rows, std::int64_t cols, Dst* __restrict__ out,
    std::int64_t ld, std::uint64_t seed_state, std::uint64_t first_draw,
    unsigned* __restrict__ bad_counter)
{
    const std::int64_t c0 =
        (std::int64_t{blockIdx.x} * blockDim.x + threadIdx.x) * kElemsPerThread;
    if (c0 >= cols) {
        return;
    }
    unsigned bad = 0;
    for (std::int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const std::uint64_t k =
            first_draw + static_cast<std::uint64_t>(r) * cols + c0;
        // state after 2k engine calls
        std::uint64_t state = mod_m31(pow_a(2 * k) * seed_state);
        Dst* dst = out + r * ld + c0;
#pragma unroll
        for (int i = 0; i < kElemsPerThread; ++i) {
            const double v = draw_uniform(state);
            if (c0 + i < cols) {
                const double av = fabs(v);
                // std::isnormal: neither zero, subnormal, infinite nor NaN
                bad += (av >= 2.2250738585072014e-308 &&
                        av <= 1.7976931348623157e308)
                           ? 0u
                           : 1u;
                dst[i] = cast_one<Dst, double>(v);
            }
        }
    }
    if (bad) {
        atomicAdd(bad_counter, bad);
    }
}

template <typename Dst>
int launch_fill(Handle* h, std::int64_t rows, std::int64_t cols, void* out,
                std::int64_t ld, std::uint32_t seed, std::uint64_t first_draw,
                cudaStream_t stream)
{
    if (rows == 0 || cols == 0) {
        return ACCBLAS_OK;
    }
    int rc = ensure_workspace(h, 0, stream);
    if (rc != ACCBLAS_OK) {
        return rc;
    }
    // linear_congruential_engine::seed: state = seed mod m, 0 -> 1
    std::uint64_t s0 = seed % kLcgM;
    if (s0 == 0) {
        s0 = 1;
    }
    unsigned* flag = control_words(h) + kCtlFillFlag;
    ACCBLAS_CUDA(cudaMemsetAsync(flag, 0, sizeof(unsigned), stream));
    const bool contiguous = (ld == cols) || rows == 1;
    const bool aligned =
        reinterpret_cast<std::uintptr_t>(out) %
            (sizeof(Dst) * kElemsPerThread) == 0;
    if (contiguous && aligned && tuning().fill_generic == 0) {
        const std::int64_t count = rows * cols;
        const std::int64_t groups =
            (count + kElemsPerThread - 1) / kElemsPerThread;
        std::int64_t grid = (groups + 255) / 256;
        const std::int64_t cap = std::int64_t{h->sm_count} * 16;
        grid = grid > cap ? cap : grid;
        const std::uint64_t threads = static_cast<std::uint64_t>(grid) * 256;
        // engine calls between the end of one group and the start of the
        // thread's next one: 2 * 4 * (T - 1)
        const std::uint32_t jump = pow_a31(8 * (threads - 1));
        fill_linear_kernel<Dst><<<static_cast<unsigned>(grid), 256, 0, stream>>>(
            count, static_cast<Dst*>(out), static_cast<std::uint32_t>(s0),
            first_draw, jump, flag);
    } else {
        const std::int64_t per_block = 256 * kElemsPerThread;
        const std::int64_t gx = (cols + per_block - 1) / per_block;
        const std::int64_t gy = rows < 32768 ? rows : 32768;
        const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(gy));
        fill_uniform_kernel<Dst><<<grid, 256, 0, stream>>>(
            rows, cols, static_cast<Dst*>(out), ld, s0, first_draw, flag);
    }
    ACCBLAS_CUDA(cudaGetLastError());
    unsigned bad = 0;
    ACCBLAS_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(unsigned),
                                 cudaMemcpyDeviceToHost, stream));
    ACCBLAS_CUDA(cudaStreamSynchronize(stream));
    if (bad != 0) {
        set_error("fill_uniform: %u draws were not normal numbers "
                  "(the host generator would have re-drawn)", bad);
        return ACCBLAS_ERR_DATA;
    }
    return ACCBLAS_OK;
}

// ---------------------------------------------------------------------------
// L1 error metric
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ double as_double(T v)
{
    return static_cast<double>(v);
}
template <>
__device__ __forceinline__ double as_double<__half>(__half v)
{
    return static_cast<double>(__half2float(v));
}

template <typename Ref, typename Res, int BLOCK>
__global__ __launch_bounds__(BLOCK) void l1_error_kernel(
    std::int64_t n, const Ref* __restrict__ ref, std::int64_t inc_ref,
    const Res* __restrict__ res, std::int64_t inc_res,
    double* __restrict__ partials, unsigned* __restrict__ counter,
    double* __restrict__ out2)
{
    __shared__ double scratch[kWarp];
    __shared__ bool is_last;
    double diff = 0.0, norm = 0.0;
    for (std::int64_t i = std::int64_t{blockIdx.x} * BLOCK + threadIdx.x; i < n;
         i += std::int64_t{gridDim.x} * BLOCK) {
        const double a = as_double(ref[i * inc_ref]);
        const double b = as_double(res[i * inc_res]);
        diff += fabs(a - b);
        norm += fabs(a);
    }
    const double d = block_sum(diff, scratch);
    const double s = block_sum(norm, scratch);
    if (threadIdx.x == 0) {
        volatile double* p = partials;
        p[2 * blockIdx.x] = d;
        p[2 * blockIdx.x + 1] = s;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) {
        return;
    }
    __threadfence();
    double dd = 0.0, ss = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += BLOCK) {
        dd += __ldcg(partials + 2 * i);
        ss += __ldcg(partials + 2 * i + 1);
    }
    dd = block_sum(dd, scratch);
    ss = block_sum(ss, scratch);
    if (threadIdx.x == 0) {
        out2[0] = dd;
        out2[1] = ss;
        *counter = 0u;
    }
}

template <typename Ref, typename Res>
int launch_l1(Handle* h, std::int64_t n, const void* ref, std::int64_t inc_ref,
              const void* res, std::int64_t inc_res, double* out2,
              cudaStream_t stream)
{
    constexpr int BLOCK = 256;
    std::int64_t grid = (n + BLOCK - 1) / BLOCK;
    const std::int64_t cap = std::int64_t{h->sm_count} * 4;
    grid = grid < 1 ? 1 : (grid > cap ? cap : grid);
    const std::int64_t max_grid = kScratchBytes / (2 * sizeof(double));
    grid = grid > max_grid ? max_grid : grid;
    int rc = ensure_workspace(h, kScratchBytes, stream);
    if (rc != ACCBLAS_OK) {
        return rc;
    }
    l1_error_kernel<Ref, Res, BLOCK>
        <<<static_cast<unsigned>(grid), BLOCK, 0, stream>>>(
            n, static_cast<const Ref*>(ref), inc_ref,
            static_cast<const Res*>(res), inc_res,
            static_cast<double*>(payload(h)),
            control_words(h) + kCtlErrCounter, out2);
    ACCBLAS_CUDA(cudaGetLastError());
    return ACCBLAS_OK;
}

template <typename F>
int dispatch_two(int a, int b, F&& f)
{
    if (!valid_dtype(a) || !valid_dtype(b)) {
        set_error("invalid dtype (%d, %d)", a, b);
        return ACCBLAS_ERR_INVALID;
    }
    auto inner = [&](auto a_tag) {
        switch (b) {
        case ACCBLAS_F64:
            return f(a_tag, double{});
        case ACCBLAS_F32:
            return f(a_tag, float{});
        default:
            return f(a_tag, __half{});
        }
    };
    switch (a) {
    case ACCBLAS_F64:
        return inner(double{});
    case ACCBLAS_F32:
        return inner(float{});
    default:
        return inner(__half{});
    }
}

}  // namespace

int convert_impl(Handle*, int dst, int src, std::int64_t rows,
                 std::int64_t cols, const void* in, std::int64_t ld_in,
                 void* out, std::int64_t ld_out, cudaStream_t stream)
{
    return dispatch_two(dst, src, [&](auto d, auto s) {
        return launch_convert<decltype(d), decltype(s)>(rows, cols, in, ld_in,
                                                        out, ld_out, stream);
    });
}

int fill_uniform_impl(Handle* h, int dst, std::int64_t rows, std::int64_t cols,
                      void* out, std::int64_t ld, std::uint32_t seed,
                      std::uint64_t first_draw, cudaStream_t stream)
{
    switch (dst) {
    case ACCBLAS_F64:
        return launch_fill<double>(h, rows, cols, out, ld, seed, first_draw,
                                   stream);
    case ACCBLAS_F32:
        return launch_fill<float>(h, rows, cols, out, ld, seed, first_draw,
                                  stream);
    case ACCBLAS_F16:
        return launch_fill<__half>(h, rows, cols, out, ld, seed, first_draw,
                                   stream);
    default:
        set_error("invalid dtype %d", dst);
        return ACCBLAS_ERR_INVALID;
    }
}

int l1_error_impl(Handle* h, int ref_t, int res_t, std::int64_t n,
                  const void* ref, std::int64_t inc_ref, const void* res,
                  std::int64_t inc_res, double* out2, cudaStream_t stream)
{
    return dispatch_two(ref_t, res_t, [&](auto a, auto b) {
        return launch_l1<decltype(a), decltype(b)>(h, n, ref, inc_ref, res,
                                                   inc_res, out2, stream);
    });
}

}  // namespace accblas
