// Development microbenchmark (run under gpurun): one-way latency of a
// "publish a value, the other SM polls it" hand-off, which is what bounds the
// blocked TRSV chain.  Variants:
//   global memory: volatile, relaxed.gpu, release/acquire.gpu, atomics
//   distributed shared memory inside a thread-block cluster
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pingpong pingpong.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

enum Mode { VOLATILE = 0, RELAXED_GPU = 1, RELEASE_ACQUIRE = 2, ATOMIC = 3, LDCV = 4 };

template <int MODE>
__device__ __forceinline__ void publish(unsigned long long* p, unsigned long long v)
{
    if (MODE == VOLATILE) {
        *reinterpret_cast<volatile unsigned long long*>(p) = v;
    } else if (MODE == RELAXED_GPU || MODE == LDCV) {
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    } else if (MODE == RELEASE_ACQUIRE) {
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    } else {
        atomicExch(p, v);
    }
}
template <int MODE>
__device__ __forceinline__ unsigned long long peek(unsigned long long* p)
{
    unsigned long long v;
    if (MODE == VOLATILE) {
        v = *reinterpret_cast<volatile unsigned long long*>(p);
    } else if (MODE == RELAXED_GPU) {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    } else if (MODE == RELEASE_ACQUIRE) {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    } else if (MODE == LDCV) {
        asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    } else {
        v = atomicAdd(p, 0ull);
    }
    return v;
}

// CTA 0 and CTA `other` bounce a counter `iters` times; one thread each.
template <int MODE>
__global__ void pingpong_global(unsigned long long* a, unsigned long long* b, int iters, int other,
                                long long* cycles, unsigned* smids)
{
    if (threadIdx.x != 0) return;
    if (blockIdx.x != 0 && blockIdx.x != other) return;
    unsigned smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    smids[blockIdx.x == 0 ? 0 : 1] = smid;
    const bool first = blockIdx.x == 0;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (first) {
            publish<MODE>(a, i);
            while (peek<MODE>(b) != (unsigned long long)i) {}
        } else {
            while (peek<MODE>(a) != (unsigned long long)i) {}
            publish<MODE>(b, i);
        }
    }
    long long t1 = clock64();
    if (first) *cycles = t1 - t0;
}

// Hand-off with an audience: CTA 0 publishes 32 consecutive values (8 lanes of
// 4 warps, like the TRSV sub-block publish), CTA 1 polls them with 32 lanes
// and answers; CTAs 2 .. 1+spectators poll the same 32 values in a tight loop
// (sleep_ns == 0) or with __nanosleep(sleep_ns) between polls.
__global__ void pingpong_herd(unsigned long long* a, unsigned long long* b, int iters, int spectators,
                              unsigned sleep_ns, long long* cycles)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (blockIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 1; i <= iters; ++i) {
            if (lane % 4 == 0) {
                publish<VOLATILE>(a + warp * 8 + lane / 4, i);
            }
            if (warp == 0) {
                while (peek<VOLATILE>(b) != (unsigned long long)i) {}
            }
            __syncthreads();
        }
        long long t1 = clock64();
        if (threadIdx.x == 0) *cycles = t1 - t0;
    } else if (blockIdx.x == 1) {
        if (warp == 0) {
            for (int i = 1; i <= iters; ++i) {
                while (peek<VOLATILE>(a + lane) < (unsigned long long)i) {}
                __syncwarp();
                if (lane == 0) publish<VOLATILE>(b, i);
            }
        }
    } else if (blockIdx.x < 2 + spectators) {
        if (warp == 0) {
            for (int i = 1; i <= iters; ++i) {
                while (peek<VOLATILE>(a + lane) < (unsigned long long)i) {
                    if (sleep_ns) __nanosleep(sleep_ns);
                }
                __syncwarp();
            }
        }
    }
}

void run_herd(int spectators, unsigned sleep_ns)
{
    unsigned long long* flags;
    long long* cyc;
    cudaMalloc(&flags, 8192);
    cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(flags, 0, 8192);
        pingpong_herd<<<148, 128>>>(flags, flags + 512, iters, spectators, sleep_ns, cyc);
        cudaDeviceSynchronize();
    }
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("herd: %3d spectators, sleep %4u ns: round trip %6.0f cycles  (%s)\n", spectators, sleep_ns,
           double(c) / iters, cudaGetErrorString(cudaGetLastError()));
    cudaFree(flags);
    cudaFree(cyc);
}

// Does a store become visible promptly when the publishing warp does NOT issue
// another memory instruction afterwards?  CTA 0: store, then (mode 0) poll at
// once / (mode 1) spin on the clock for `delay` cycles first / (mode 2)
// __syncthreads, then the clock spin.  CTA 1 stamps nothing; it answers as
// soon as it sees the value.  Reported: round trip minus the delay actually
// spent (if the store left at once, the answer is already there: ~0..small).
__global__ void store_linger(unsigned long long* a, unsigned long long* b, int iters, int mode, int delay,
                             long long* cycles, int background, const uint4* bg, uint4* sink)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (blockIdx.x == 0) {
        long long total = 0;
        for (int i = 1; i <= iters; ++i) {
            long long t0 = clock64();
            if (lane % 4 == 0) {
                publish<VOLATILE>(a + warp * 8 + lane / 4, i);
            }
            if (mode == 2) __syncthreads();
            long long spent = 0;
            if (mode >= 1) {
                long long t = clock64();
                while (clock64() - t < delay) {}
                spent = clock64() - t;
            }
            if (warp == 0) {
                while (peek<VOLATILE>(b) != (unsigned long long)i) {}
            }
            __syncthreads();
            total += clock64() - t0 - spent;
        }
        if (threadIdx.x == 0) *cycles = total;
    } else if (blockIdx.x == 1) {
        if (warp == 0) {
            for (int i = 1; i <= iters; ++i) {
                while (peek<VOLATILE>(a + 24 + lane / 4) < (unsigned long long)i) {}
                __syncwarp();
                if (lane == 0) publish<VOLATILE>(b, i);
            }
        }
    } else if (background) {
        // memory traffic from every other SM while the two CTAs talk
        uint4 acc = make_uint4(0, 0, 0, 0);
        const size_t n16 = (size_t(1) << 30) / 16;
        size_t idx = (size_t(blockIdx.x) * blockDim.x + threadIdx.x);
        volatile unsigned long long* done = b + 8;
        for (int rep = 0; rep < 100000; ++rep) {
            for (int u = 0; u < 8; ++u) {
                uint4 v = bg[(idx + size_t(u) * 148 * 128 + size_t(rep) * 8 * 148 * 128) % n16];
                acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
            }
            if (*done) break;
        }
        if (acc.x == 0x12345) sink[idx] = acc;
    }
    if (blockIdx.x == 0) {
        __syncthreads();
        if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(b + 8) = 1;
    }
}

void run_linger(int mode, int delay, int background)
{
    unsigned long long* flags;
    long long* cyc;
    uint4* bg;
    cudaMalloc(&flags, 8192);
    cudaMalloc(&cyc, 8);
    cudaMalloc(&bg, size_t(1) << 30);
    const int iters = 500;
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(flags, 0, 8192);
        store_linger<<<148, 128>>>(flags, flags + 512, iters, mode, delay, cyc, background, bg, bg);
        cudaDeviceSynchronize();
    }
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("linger: mode %d delay %5d background %d: round trip minus delay %6.0f cycles  (%s)\n", mode, delay,
           background, double(c) / iters, cudaGetErrorString(cudaGetLastError()));
    cudaFree(flags);
    cudaFree(cyc);
    cudaFree(bg);
}

// same inside a cluster of `CS` CTAs through distributed shared memory:
// rank 0 <-> rank `other`
__global__ void pingpong_dsmem(int iters, int other, long long* cycles, unsigned* smids)
{
    __shared__ unsigned long long flag;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    if (threadIdx.x == 0) flag = 0;
    cluster.sync();
    if (threadIdx.x == 0 && (rank == 0 || rank == (unsigned)other)) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        smids[rank == 0 ? 0 : 1] = smid;
        const unsigned peer = rank == 0 ? other : 0;
        unsigned long long* remote = cluster.map_shared_rank(&flag, peer);
        volatile unsigned long long* mine = &flag;
        long long t0 = clock64();
        for (int i = 1; i <= iters; ++i) {
            if (rank == 0) {
                *reinterpret_cast<volatile unsigned long long*>(remote) = i;
                while (*mine != (unsigned long long)i) {}
            } else {
                while (*mine != (unsigned long long)i) {}
                *reinterpret_cast<volatile unsigned long long*>(remote) = i;
            }
        }
        long long t1 = clock64();
        if (rank == 0) *cycles = t1 - t0;
    }
    cluster.sync();
}

template <int MODE>
void run_global(const char* name, int other)
{
    unsigned long long* flags;
    long long* cyc;
    unsigned* smids;
    cudaMalloc(&flags, 4096);
    cudaMalloc(&cyc, 8);
    cudaMalloc(&smids, 8);
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(flags, 0, 4096);
        pingpong_global<MODE><<<148, 32>>>(flags, flags + 256, iters, other, cyc, smids);
        cudaDeviceSynchronize();
    }
    long long c;
    unsigned s[2];
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(s, smids, 8, cudaMemcpyDeviceToHost);
    printf("global %-16s CTA 0 (sm %3u) <-> CTA %3d (sm %3u): one-way %6.0f cycles  (%s)\n", name, s[0], other,
           s[1], double(c) / iters / 2, cudaGetErrorString(cudaGetLastError()));
    cudaFree(flags);
    cudaFree(cyc);
    cudaFree(smids);
}

void run_dsmem(int cs, int other)
{
    long long* cyc;
    unsigned* smids;
    cudaMalloc(&cyc, 8);
    cudaMalloc(&smids, 8);
    cudaMemset(cyc, 0, 8);
    const int iters = 2000;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs);
    cfg.blockDim = dim3(32);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cs > 8) {
        cudaFuncSetAttribute(pingpong_dsmem, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    }
    cudaError_t e = cudaSuccess;
    for (int rep = 0; rep < 2; ++rep) {
        e = cudaLaunchKernelEx(&cfg, pingpong_dsmem, iters, other, cyc, smids);
        cudaDeviceSynchronize();
    }
    long long c;
    unsigned s[2];
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(s, smids, 8, cudaMemcpyDeviceToHost);
    printf("dsmem cluster %2d: rank 0 (sm %3u) <-> rank %2d (sm %3u): one-way %6.0f cycles  (%s / %s)\n", cs, s[0],
           other, s[1], double(c) / iters / 2, cudaGetErrorString(e), cudaGetErrorString(cudaGetLastError()));
    cudaFree(cyc);
    cudaFree(smids);
}

int main()
{
    for (int other : {1, 74}) {
        run_global<VOLATILE>("volatile", other);
        run_global<RELAXED_GPU>("relaxed.gpu", other);
        run_global<RELEASE_ACQUIRE>("release/acquire", other);
        run_global<LDCV>("st.relaxed/ld.cv", other);
        run_global<ATOMIC>("atomic", other);
    }
    for (int bgd : {0, 1}) {
        run_linger(0, 0, bgd);
        run_linger(1, 3000, bgd);
        run_linger(2, 3000, bgd);
    }
    for (int spect : {0, 126}) {
        run_herd(spect, 0);
    }
    for (unsigned ns : {50u, 100u, 200u, 500u}) {
        run_herd(126, ns);
    }
    for (int cs : {2, 4, 8, 16}) {
        for (int other : {1, cs - 1}) {
            run_dsmem(cs, other);
        }
    }
    return 0;
}
