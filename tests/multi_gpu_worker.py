"""One process per GPU (launched by tests/test_gpu_multi.py through
torch.distributed.run, or by hand on a multi-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_worker.py [bench]

Checks of the sharded paths of BASELINE.json configs[4] (the reference itself
is single-device, /root/reference/cuda/dot_kernels.cuh:33):
  GEMV  every rank's row slab is bit-identical to the same rows of the
        full-matrix GEMV on one GPU;
  DOT   accblas_dot_allreduce (exchange inside the kernel over peer memory)
        returns, on EVERY rank and call after call, exactly the rank-order sum
        of the per-rank accblas_dot partials; it agrees with the 1-element
        NCCL all-reduce path to rounding and with the double-double oracle on
        the whole vectors;
  the sticky time-out of the exchange (one rank stays away on purpose).
With `bench`: 20 back-to-back calls of both DOT paths, max over ranks.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import accessor_blas_b200 as ab  # noqa: E402
from accessor_blas_b200 import capi, sharded  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    h = ab.Handle(local)
    names = {torch.float64: "fp64", torch.float32: "fp32", torch.float16: "fp16"}

    # ---- GEMV: row slabs against the full matrix on one GPU -------------------
    m, n = 4 * 1031 * world + 6, 4096 + 8
    for ar, st in ((torch.float64, torch.float32), (torch.float32, torch.float16),
                   (torch.float64, torch.float64)):
        A = torch.empty(m * n, dtype=st, device=dev)
        x = torch.empty(n, dtype=st, device=dev)
        y0 = torch.empty(m, dtype=st, device=dev)
        h.fill_uniform(m, n, A, n, 42, 0)
        if rank == 0:
            h.fill_uniform(n, 1, x, 1, 42, m * n)
        sharded.broadcast_vector(x)
        h.fill_uniform(m, 1, y0, 1, 42, m * n + n)
        full = y0.clone()
        h.gemv(ar, m, n, 1.0, A, n, x, 1, 1.0, full, 1)
        sg = sharded.ShardedGemv(h, ar, m, n, n)
        first, rows = sg.first, sg.rows
        # the slab is generated in place from the global draw index, as bench.py does
        A_loc = torch.empty(rows * n, dtype=st, device=dev)
        h.fill_uniform(rows, n, A_loc, n, 42, first * n)
        assert torch.equal(A_loc, A[first * n:(first + rows) * n])
        y_loc = y0[first:first + rows].clone()
        sg(1.0, A_loc, x, 1.0, y_loc)
        assert torch.equal(y_loc, full[first:first + rows]), (rank, names[ar], names[st])
        gathered = sharded.gather_rows(y_loc, m)
        assert torch.equal(gathered, full), (rank, "gather", names[ar], names[st])
        del A, A_loc

    # ---- DOT: in-kernel exchange == rank-order sum of the partials ------------
    from oracle_binding import Oracle
    orc = Oracle()
    nd = 2 ** 22 + 12345
    for ar, st in ((torch.float64, torch.float32), (torch.float32, torch.float16),
                   (torch.float64, torch.float64), (torch.float32, torch.float32)):
        nccl = sharded.ShardedDot(h, ar, nd, fused=False)
        fused = sharded.ShardedDot(h, ar, nd, fused=True)
        assert fused.fused, "peer connection failed"
        xl = torch.empty(nccl.count, dtype=st, device=dev)
        yl = torch.empty(nccl.count, dtype=st, device=dev)
        h.fill_uniform(1, nccl.count, xl, nccl.count, 42, nccl.first)
        h.fill_uniform(1, nccl.count, yl, nccl.count, 42, nd + nccl.first)
        ref = nccl(xl, yl, ar)
        outs = [fused(xl, yl, ar) for _ in range(5)]
        assert len({o.data_ptr() for o in outs}) == len(outs)   # results are not aliased
        torch.cuda.synchronize()
        part = torch.zeros(1, dtype=ar, device=dev)
        h.dot(ar, nccl.count, xl, 1, yl, 1, part)
        parts = [torch.zeros(1, dtype=ar, device=dev) for _ in range(world)]
        dist.all_gather(parts, part)
        want = torch.zeros(1, dtype=ar, device=dev)
        for p in parts:      # rank order, in the arithmetic type
            want = want + p
        for o in outs:
            assert torch.equal(o, want), (rank, names[ar], names[st], o.item(), want.item())
        allv = [torch.zeros(1, dtype=ar, device=dev) for _ in range(world)]
        dist.all_gather(allv, outs[-1])
        assert all(torch.equal(v, allv[0]) for v in allv)          # same bits everywhere
        rel = abs(ref.item() - want.item()) / max(abs(want.item()), 1e-300)
        assert rel < (1e-12 if ar == torch.float64 else 1e-5), (ref.item(), want.item())
        # the whole vectors against the double-double oracle (rank 0)
        if rank == 0:
            xh = torch.empty(nd, dtype=st, device=dev)
            yh = torch.empty(nd, dtype=st, device=dev)
            h.fill_uniform(1, nd, xh, nd, 42, 0)
            h.fill_uniform(1, nd, yh, nd, 42, nd)
            xn, yn = xh.cpu().numpy(), yh.cpu().numpy()
            exact = orc.exact_dot(xn, yn)
            scale = float(np.abs(xn.astype(np.float64) * yn.astype(np.float64)).sum())
            tol = 5e-14 if ar == torch.float64 else 2e-5
            assert abs(want.item() - exact) <= tol * scale, (want.item(), exact)
        h.peer_disconnect()
        dist.barrier()

    # ---- a peer that stays away: NaN + sticky ACCBLAS_ERR_PEER, no hang -------
    if world >= 2:
        h2 = ab.Handle(local)
        assert sharded.connect_peers(h2)
        h2.peer_set_timeout(0.2)
        res = torch.zeros(1, dtype=torch.float64, device=dev)
        xl = torch.ones(1024, dtype=torch.float32, device=dev)
        if rank != world - 1:     # the last rank does not call
            h2.dot_allreduce(torch.float64, 1024, xl, 1, xl, 1, res)
            torch.cuda.synchronize()
            assert torch.isnan(res).item()
            assert h2.peer_status() == 1
            try:
                h2.dot_allreduce(torch.float64, 1024, xl, 1, xl, 1, res)
                raise AssertionError("expected ACCBLAS_ERR_PEER")
            except ab.AccblasError as e:
                assert e.status == capi.ERR_PEER
        dist.barrier()
        h2.peer_disconnect()
        dist.barrier()

    if len(sys.argv) > 1 and sys.argv[1] == "bench":
        nb = 2 ** 26
        for ar, st in ((torch.float64, torch.float32), (torch.float32, torch.float16)):
            nccl = sharded.ShardedDot(h, ar, nb, fused=False)
            fused = sharded.ShardedDot(h, ar, nb, fused=True)
            xl = torch.empty(nccl.count, dtype=st, device=dev)
            yl = torch.empty(nccl.count, dtype=st, device=dev)
            h.fill_uniform(1, nccl.count, xl, nccl.count, 42, nccl.first)
            h.fill_uniform(1, nccl.count, yl, nccl.count, 42, nb + nccl.first)
            out = torch.zeros(1, dtype=ar, device=dev)
            for name, op in (("nccl", nccl), ("fused", fused)):
                for _ in range(3):
                    op(xl, yl, ar, out=out)
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    op(xl, yl, ar, out=out)
                e1.record()
                torch.cuda.synchronize()
                ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                if rank == 0:
                    print(f"DOT n={nb} Acc<{names[ar]},{names[st]}> x{world} GPUs, {name}: "
                          f"{ms.item() * 1e3:.1f} us per call", flush=True)
            h.peer_disconnect()
            dist.barrier()
    dist.barrier()
    if rank == 0:
        print("multi-GPU ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
