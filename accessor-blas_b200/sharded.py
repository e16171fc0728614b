"""Multi-GPU sharding of the two paths that shard.

One process per GPU, `torch.distributed` (NCCL over NVLink) for the plumbing.

  GEMV  rows are split into contiguous slabs, one per rank; x is broadcast
        once; every rank writes its own slice of the result.  There is no
        reduction across ranks, so the result is bit-identical to the 1-GPU
        run of the same kernel configuration.
  DOT   the index range is split into contiguous, 16-byte aligned chunks; each
        rank produces one partial in the arithmetic type with the
        deterministic two-pass kernel; ONE single-element all-reduce combines
        them; the cast to the result type happens after the reduction.  With
        fused=True the exchange happens inside the DOT kernel over peer memory
        (accblas_dot_allreduce): no collective call at all.
  TRSV  does not shard (one dependency chain): replicas only.

The reference has no multi-device code at all (device 0 is hard-coded,
/root/reference/cuda/dot_kernels.cuh:33); this module is new work defined by
BASELINE.json's config 5.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def row_partition(m: int, world: int, rank: int, align: int = 4) -> Tuple[int, int]:
    """(first_row, num_rows) of `rank`'s slab: contiguous, sizes differ by at
    most one `align`-row group (plus the ragged tail), every slab but the last a multiple of `align` rows (the GEMV
    kernel works on groups of 4 rows)."""
    if world < 1 or not 0 <= rank < world or m < 0:
        raise ValueError("bad partition request")
    groups = (m + align - 1) // align
    base, rem = divmod(groups, world)
    first_group = rank * base + min(rank, rem)
    n_groups = base + (1 if rank < rem else 0)
    first = min(first_group * align, m)
    last = min((first_group + n_groups) * align, m)
    return first, last - first


def range_partition(n: int, world: int, rank: int, align: int = 8) -> Tuple[int, int]:
    """(first_index, count) of `rank`'s chunk of a length-n vector; chunk
    starts are multiples of `align` elements so 128-bit loads stay aligned for
    every storage type (8 halves = 16 bytes)."""
    return row_partition(n, world, rank, align)


def broadcast_vector(x: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    """x is replicated once before the (repeated) GEMVs.  fp16 tensors are
    moved as raw bytes so every backend accepts them."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        payload = x.view(torch.uint8) if x.dtype == torch.float16 else x
        dist.broadcast(payload, src=src, group=group)
    return x


def allreduce_partial(partial: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the per-rank partials (one element, arithmetic type)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def gather_rows(y_local: torch.Tensor, m: int, group=None) -> torch.Tensor:
    """Optional: assemble the full result on every rank."""
    world = dist.get_world_size(group)
    sizes = [row_partition(m, world, r)[1] for r in range(world)]
    # all_gather wants equal shapes: pad every slab to the largest one
    widest = max(sizes)
    padded = torch.zeros(widest, dtype=y_local.dtype, device=y_local.device)
    padded[:y_local.numel()] = y_local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])


class ShardedGemv:
    """Row-sharded y = alpha*A*x + beta*y.  Each rank holds rows
    [first, first+rows) of A (row stride lda) and the matching slice of y."""

    def __init__(self, handle, ar, m: int, n: int, lda: int, group=None):
        self.handle, self.ar, self.m, self.n, self.lda = handle, ar, m, n, lda
        self.group = group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.first, self.rows = row_partition(m, world, rank)

    def __call__(self, alpha: float, A_local: torch.Tensor, x: torch.Tensor,
                 beta: float, y_local: torch.Tensor, stream=None) -> torch.Tensor:
        self.handle.gemv(self.ar, self.rows, self.n, alpha, A_local, self.lda,
                         x, 1, beta, y_local, 1, stream)
        return y_local


def connect_peers(handle, group=None) -> bool:
    """One-time set-up of the in-kernel all-reduce: every rank exports the CUDA
    IPC handle of its mailbox, the handles are gathered in rank order and each
    rank maps its peers' mailboxes.  Returns False (nothing connected) when
    there is a single rank, more than 8, or no CUDA process group."""
    if not (dist.is_available() and dist.is_initialized()):
        return False
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world < 2 or world > 8:
        return False
    key = (id(group), world, rank)
    if getattr(handle, "_peer_group", None) == key:
        return True          # this handle already belongs to this group
    if getattr(handle, "_peer_group", None) is not None:
        # a handle belongs to one group at a time: leave the old one first (every
        # rank does, after the group's last call has completed everywhere)
        torch.cuda.synchronize()
        dist.barrier(group=group)
        handle.peer_disconnect()
    mine = handle.peer_export()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    handle.peer_connect_ipc(world, rank, b"".join(gathered))
    handle._peer_group = key
    # no barrier needed: the mailbox was zeroed when it was exported and connect
    # does not touch it, so an entry a faster peer publishes right away stays
    return True


class ShardedDot:
    """Range-sharded dot product.

    fused=False: per-rank partial + ONE single-element NCCL all-reduce.
    fused=True : the kernel's last CTA exchanges the partials with the peers
                 over NVLink (P2P stores into every rank's mailbox) and sums
                 them in rank order -- one launch per rank, no collective
                 call, identical bits on all ranks (`connect_peers` first).
    """

    def __init__(self, handle, ar, n: int, group=None, fused: bool = False):
        self.handle, self.ar, self.n, self.group = handle, ar, n, group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.first, self.count = range_partition(n, world, rank)
        self._partial: Optional[torch.Tensor] = None
        self.fused = bool(fused) and world > 1
        if self.fused and not connect_peers(handle, group):
            self.fused = False

    def __call__(self, x_local: torch.Tensor, y_local: torch.Tensor,
                 res_dtype: torch.dtype, stream=None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Returns a one-element tensor of `res_dtype` holding the dot product
        of the whole vectors.  The result is written to `out` if given, else
        to a FRESH tensor -- never to a buffer a later call overwrites (a CG
        step keeps alpha and beta alive across two calls)."""
        if out is None:
            out = torch.empty(1, dtype=res_dtype, device=x_local.device)
        elif out.dtype != res_dtype or out.numel() != 1:
            raise ValueError("out must be a one-element tensor of res_dtype")
        if self.fused:
            self.handle.dot_allreduce(self.ar, self.count, x_local, 1, y_local, 1,
                                      out, stream)
            return out
        if self._partial is None:
            self._partial = torch.zeros(1, dtype=self.ar, device=x_local.device)
        self.handle.dot(self.ar, self.count, x_local, 1, y_local, 1,
                        self._partial, stream)
        allreduce_partial(self._partial, self.group)
        out.copy_(self._partial)      # the cast to the result type, after the reduction
        return out
