"""Host-side logic of the N>1 path on CPU: world_size-2 (and 3) `gloo` groups.
The per-rank compute is stood in for by the CPU oracle (tests only); what is
under test is the partitioning and the collectives of sharded.py."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def test_partitions_cover_everything():
    from accessor_blas_b200 import sharded
    for total in (0, 1, 7, 8, 9, 1000, 16384, 131072, 2 ** 32, 12345677):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            for rank in range(world):
                first, count = sharded.row_partition(total, world, rank)
                assert first == nxt and count >= 0
                if rank < world - 1 and count:
                    assert (first + count) % 4 == 0 or first + count == total
                nxt = first + count
            assert nxt == total
            starts = [sharded.range_partition(total, world, r)[0] for r in range(world)]
            assert all(s % 8 == 0 or s == total for s in starts)
            sizes = [sharded.row_partition(total, world, r)[1] for r in range(world)]
            assert max(sizes) - min(sizes) < 8


def _worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from accessor_blas_b200 import sharded
    from oracle_binding import Oracle
    orc = Oracle()
    orc.L.oracle_set_num_threads(1)

    # --- DOT: range-sharded partial + one all-reduce --------------------------
    n = 100_003
    x = orc.uniform(n, seed=42).astype(np.float32)
    y = orc.uniform(n, seed=42, first_draw=n).astype(np.float32)
    first, count = sharded.range_partition(n, world, rank)
    partial = orc.cpu_dot(np.float64, x[first:first + count].copy(),
                          y[first:first + count].copy(), np.float64)
    t = torch.tensor([partial], dtype=torch.float64)
    sharded.allreduce_partial(t)
    exact = orc.exact_dot(x, y)
    assert abs(t.item() - exact) < 1e-11, (t.item(), exact)

    # --- GEMV: x broadcast once, row slabs, optional gather --------------------
    m, k = 103, 257
    A = orc.uniform(m * k, seed=7).astype(np.float32)
    xv = torch.from_numpy(orc.uniform(k, seed=8).astype(np.float32)) if rank == 0 \
        else torch.zeros(k, dtype=torch.float32)
    sharded.broadcast_vector(xv)
    r0, rows = sharded.row_partition(m, world, rank)
    y_local = np.zeros(rows, dtype=np.float32)
    orc.cpu_gemv(np.float64, A[r0 * k:(r0 + rows) * k].copy(), rows, k, k,
                 xv.numpy().copy(), 1.0, 0.0, y_local)
    full = sharded.gather_rows(torch.from_numpy(y_local), m)
    want = np.zeros(m, dtype=np.float32)
    orc.cpu_gemv(np.float64, A, m, k, k, orc.uniform(k, seed=8).astype(np.float32),
                 1.0, 0.0, want)
    assert np.array_equal(full.numpy(), want)  # no cross-rank reduction: bit-identical

    # --- in-kernel exchange: the set-up handshake (host logic only; the GPU
    #     side is tests/test_gpu_parity.py + tests/multi_gpu_dot.py) ----------
    class FakeHandle:
        def __init__(self, r):
            self.r, self.connected, self.calls = r, None, 0

        def peer_export(self):
            return bytes([self.r]) * 64

        def peer_connect_ipc(self, w, r, handles):
            self.connected = (w, r, handles)

        def dot_allreduce(self, ar, n_, x_, incx, y_, incy, result, stream=None):
            self.calls += 1
            result.fill_(float(n_))

    fh = FakeHandle(rank)
    sd = sharded.ShardedDot(fh, torch.float64, n, fused=True)
    assert sd.fused
    w, r, handles = fh.connected
    assert (w, r) == (world, rank)
    assert handles == b"".join(bytes([q]) * 64 for q in range(world))  # rank order
    xl = torch.zeros(sd.count, dtype=torch.float32)
    out = sd(xl, xl, torch.float32)
    assert fh.calls == 1 and out.dtype == torch.float32 and out.item() == float(sd.count)
    assert not sharded.ShardedDot(FakeHandle(rank), torch.float64, n, fused=False).fused

    # fp16 vectors travel as raw words
    hv = torch.arange(16, dtype=torch.float16) if rank == 0 else torch.zeros(16, dtype=torch.float16)
    sharded.broadcast_vector(hv)
    assert hv.tolist() == list(range(16))
    Path(out_dir, f"ok{rank}").write_text("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_paths_over_gloo(tmp_path, world):
    import accessor_blas_b200  # noqa: F401  (registers the package for children)
    port = 29500 + os.getpid() % 2000 + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
