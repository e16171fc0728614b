"""Per-phase timeline of the TRSV kernel (development tool; run under gpurun).

Slots per block row k (SM cycles unless noted):
  0 start | 1 diag tile in smem | 2 sub-blocks inverted | 3 last dependency:
  A loads issued | 4 x block seen by the poller | 5 CTA released | 6 row sums
  reduced | 7..10 sub-steps done | 12 globaltimer(ns) at the end |
  13 globaltimer(ns) when the last x block was seen
"""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import accessor_blas_b200 as ab  # noqa: E402
from accessor_blas_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
st = {"f64": torch.float64, "f32": torch.float32, "f16": torch.float16}[
    sys.argv[2] if len(sys.argv) > 2 else "f32"]
dev = torch.device("cuda:0")
h = ab.Handle(0)
lib = capi.load()
fn = lib.accblas_dev_trsv_trace
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
               ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
               ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]

g = torch.empty(n * n, dtype=torch.float64, device=dev)
h.fill_uniform(n, n, g, n, 42, 0)
LU, _ = torch.linalg.lu_factor(g.view(n, n))
A = LU.contiguous().view(-1).to(st)
b = torch.empty(n, dtype=torch.float64, device=dev)
h.fill_uniform(n, 1, b, 1, 42, n * n)
b = b.to(st)
nb = (n + 127) // 128
trace = torch.zeros(nb * 64, dtype=torch.int64, device=dev)
for it in range(3):
    x = b.clone()
    trace.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(h._h, 0, ab.dtype_code(st), ab.LOWER, ab.UNIT, n, A.data_ptr(), n, x.data_ptr(), 1,
            trace.data_ptr(), torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0
    print(f"run {it}: {e0.elapsed_time(e1) * 1e3:.1f} us total")
t = trace.cpu().numpy().reshape(nb, 64)
names = ["load diag", "invert", "(wait) ->issue last", "poll last x", "release", "tile+reduce"]
for k in sorted(set([0, 1, 2, nb // 4, nb // 2, nb - 2, nb - 1])):
    if k < 0 or k >= nb:
        continue
    d = np.diff(t[k, :7])
    print(f"block {k:4d}: " + "  ".join(f"{nm}={int(v)}" for nm, v in zip(names, d)))
for k in (nb // 2, nb - 1):
    b5 = t[k, 5]
    print(f"block {k} tile phase (cycles after the poller's barrier arrival): x in registers "
          f"+{int(t[k, 32] - b5)}, FMAs done +{int(t[k, 33] - b5)}, row sum written "
          f"+{int(t[k, 34] - b5)}, barrier passed +{int(t[k, 6] - b5)}")
    base = t[k, 6]
    parts = []
    for step in range(4):
        b, r, u = (int(t[k, 16 + 4 * step + i] - base) for i in range(3))
        parts.append(f"step{step}: matvec begin +{b}, sol ready +{r}" +
                     (f", update done +{u}" if step < 3 else ""))
    print(f"block {k} (cycles after the row-sum barrier): " + "; ".join(parts) +
          f"; loop end +{int(t[k, 10] - base)}")
ends = t[:, 12].astype(np.float64)
seen = t[:, 13].astype(np.float64)
step = np.diff(ends)
print(f"end-to-end per block step (globaltimer ns): median {np.median(step):.0f}, "
      f"mean {step.mean():.0f}, p10 {np.percentile(step, 10):.0f}, p90 {np.percentile(step, 90):.0f}")
lat = seen[1:] - ends[:-1]
print(f"publish(k-1 end) -> seen by k (ns): median {np.median(lat):.0f}, "
      f"p10 {np.percentile(lat, 10):.0f}, p90 {np.percentile(lat, 90):.0f}")
crit = (t[1:, 10] - t[1:, 4]).astype(np.float64)
print(f"seen -> own end (cycles): median {np.median(crit):.0f}")
print(f"first block done at {(ends[0] - ends.min()):.0f} ns after the earliest end; "
      f"chain length {(ends.max() - ends[0]) / 1e3:.1f} us")
