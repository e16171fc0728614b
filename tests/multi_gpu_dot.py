"""Two or more ranks, one process per GPU (run under torchrun on a multi-GPU
box; not collected by pytest): the in-kernel all-reduce of ShardedDot(fused)
against the NCCL path.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multi_gpu_dot.py
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import accessor_blas_b200 as ab  # noqa: E402
from accessor_blas_b200.sharded import ShardedDot  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    h = ab.Handle(local)
    dev = torch.device(f"cuda:{local}")
    n = 2 ** 26 + 12345
    for ar, st in ((torch.float64, torch.float32), (torch.float32, torch.float16),
                   (torch.float64, torch.float64)):
        nccl = ShardedDot(h, ar, n, fused=False)
        fused = ShardedDot(h, ar, n, fused=True)
        assert fused.fused, "peer connection failed"
        x = torch.empty(nccl.count, dtype=st, device=dev)
        y = torch.empty(nccl.count, dtype=st, device=dev)
        h.fill_uniform(1, nccl.count, x, nccl.count, 42, nccl.first)
        h.fill_uniform(1, nccl.count, y, nccl.count, 42, n + nccl.first)
        ref = nccl(x, y, ar).clone()
        outs = [fused(x, y, ar).clone() for _ in range(4)]
        torch.cuda.synchronize()
        # rank-order sum of the partials, formed independently on the host
        part = torch.zeros(1, dtype=ar, device=dev)
        h.dot(ar, nccl.count, x, 1, y, 1, part)
        parts = [torch.zeros(1, dtype=ar, device=dev) for _ in range(world)]
        dist.all_gather(parts, part)
        want = torch.zeros(1, dtype=ar, device=dev)
        for p in parts:
            want = want + p
        for o in outs:
            assert torch.equal(o, want), (rank, ar, st, o.item(), want.item())
        rel = abs(ref.item() - want.item()) / max(abs(want.item()), 1e-300)
        assert rel < (1e-12 if ar == torch.float64 else 1e-5), (ref.item(), want.item())
        # every rank holds the same bits
        allv = [torch.zeros(1, dtype=ar, device=dev) for _ in range(world)]
        dist.all_gather(allv, outs[-1])
        assert all(torch.equal(v, allv[0]) for v in allv)
        # timing: 20 back-to-back calls each
        for name, op in (("nccl", nccl), ("fused", fused)):
            for _ in range(3):
                op(x, y, ar)
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                op(x, y, ar)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"DOT n={n} Acc<{ar},{st}> x{world} GPUs, {name}: {ms.item() * 1e3:.1f} us per call",
                      flush=True)
    dist.barrier()
    if rank == 0:
        print("multi-GPU dot ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
