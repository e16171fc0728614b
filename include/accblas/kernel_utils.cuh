// Device-side reduction helpers with the reference's names and call shapes
// (/root/reference/cuda/kernel_utils.cuh): a sub-warp butterfly over a
// cooperative-groups tile and a block reduction whose result lands in
// shared[0].  Written for sm_100a only: no pre-CUDA-11 path, the block
// reduction folds one value per warp with shuffles instead of halving through
// shared memory with a barrier per level.
#pragma once

#include <cinttypes>

#include <cooperative_groups.h>

namespace kernel {

namespace cg = cooperative_groups;
constexpr int WARP_SIZE{32};

// Butterfly reduction inside a (sub-)warp tile; every thread of the tile
// returns the same value.  Masks 1, 2, 4, ... as in the reference, so the
// association order per lane is the same.
template <unsigned int subgroup_size, typename ValueType, typename Callable,
          typename... TileParams>
__device__ __forceinline__ ValueType reduce(
    const cg::thread_block_tile<subgroup_size, TileParams...>& tile,
    ValueType local_data, Callable&& reduce_op)
{
#pragma unroll
    for (unsigned int mask = 1; mask < subgroup_size; mask <<= 1) {
        local_data = reduce_op(local_data, tile.shfl_xor(local_data, mask));
    }
    return local_data;
}

// Reduces group.size() values held in `shared` (one per thread, already
// written by the caller); the result is left in shared[0].  group.size() must
// be a multiple of WARP_SIZE and at most WARP_SIZE * WARP_SIZE.
template <typename Group, typename ValueType, typename Callable>
__device__ void reduce(Group&& group, ValueType* __restrict__ shared,
                       Callable&& reduce_op)
{
    const auto tid = group.thread_rank();
    const auto warp = cg::tiled_partition<WARP_SIZE>(group);
    const unsigned num_warps = group.size() / WARP_SIZE;
    group.sync();
    ValueType v = reduce(warp, shared[tid], reduce_op);
    group.sync();
    if (warp.thread_rank() == 0) {
        shared[tid / WARP_SIZE] = v;
    }
    group.sync();
    if (tid < WARP_SIZE) {
        // pad with copies of the first partial's neutral partner: fold only
        // the live slots, in slot order
        ValueType total = shared[0];
        for (unsigned w = 1; w < num_warps; ++w) {
            total = reduce_op(total, shared[w]);
        }
        if (tid == 0) {
            shared[0] = total;
        }
    }
    group.sync();
}

namespace detail {

// One value per thread -> block total: warp butterfly + one shared slot per warp; result valid in thread 0
template <int block_size, typename T>
__device__ __forceinline__ T block_total(T v)
{
    __shared__ T slots[WARP_SIZE];
#pragma unroll
    for (int mask = WARP_SIZE / 2; mask > 0; mask >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, mask);
    }
    const int lane = threadIdx.x % WARP_SIZE;
    const int warp = threadIdx.x / WARP_SIZE;
    if (lane == 0) {
        slots[warp] = v;
    }
    __syncthreads();
    T total{};
    if (warp == 0) {
        constexpr int warps = (block_size + WARP_SIZE - 1) / WARP_SIZE;
        total = lane < warps ? slots[lane] : T{};
#pragma unroll
        for (int mask = WARP_SIZE / 2; mask > 0; mask >>= 1) {
            total += __shfl_xor_sync(0xffffffffu, total, mask);
        }
    }
    return total;
}

}  // namespace detail

}  // namespace kernel
