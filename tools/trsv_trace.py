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
if len(sys.argv) > 4:
    ab.tune("trsv_l2_ahead", int(sys.argv[4]))
st = {"f64": torch.float64, "f32": torch.float32, "f16": torch.float16}[
    sys.argv[2] if len(sys.argv) > 2 else "f32"]
ar_code = {"f64": 0, "f32": 1}[sys.argv[3] if len(sys.argv) > 3 else "f64"]
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
    rc = fn(h._h, ar_code, ab.dtype_code(st), ab.LOWER, ab.UNIT, n, A.data_ptr(), n, x.data_ptr(), 1,
            trace.data_ptr(), torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0
    print(f"run {it}: {e0.elapsed_time(e1) * 1e3:.1f} us total")
t = trace.cpu().numpy().reshape(nb, 64)
names = ["load diag", "invert+products", "(wait) ->last panel", "last x seen", "last FMAs",
         "reduce+group barrier"]
for k in sorted(set([0, 1, 2, nb // 4, nb // 2, nb - 2, nb - 1])):
    if k < 0 or k >= nb:
        continue
    d = np.diff(t[k, :7])
    print(f"block {k:4d}: " + "  ".join(f"{nm}={int(v)}" for nm, v in zip(names, d)))
for k in (nb // 2, nb - 1):
    base = t[k, 4]
    parts = []
    for step in range(4):
        b, r = (int(t[k, 16 + 4 * step + i] - base) for i in range(2))
        parts.append(f"x{step}: begin +{b}, in smem +{r}")
    print(f"block {k} (cycles after its last x arrived): tile FMAs done +{int(t[k, 5] - base)}; "
          f"rhs complete +{int(t[k, 6] - base)}; y FMAs +{int(t[k, 35] - base)}; y reduced +{int(t[k, 36] - base)}; " +
          "; ".join(parts) + f"; CTA done +{int(t[k, 10] - base)}")
ends = t[:, 12].astype(np.float64)
seen = t[:, 13].astype(np.float64)
step = np.diff(ends)
print(f"end-to-end per block step (globaltimer ns): median {np.median(step):.0f}, "
      f"mean {step.mean():.0f}, p10 {np.percentile(step, 10):.0f}, p90 {np.percentile(step, 90):.0f}")
lat = seen[1:] - ends[:-1]
print(f"publish(k-1 end) -> seen by k (ns): median {np.median(lat):.0f}, "
      f"p10 {np.percentile(lat, 10):.0f}, p90 {np.percentile(lat, 90):.0f}")
crit = (t[1:, 29] - t[1:, 4]).astype(np.float64)
print(f"last x seen -> own last sub-block in smem (cycles): median {np.median(crit):.0f}")
print(f"first block done at {(ends[0] - ends.min()):.0f} ns after the earliest end; "
      f"chain length {(ends.max() - ends[0]) / 1e3:.1f} us")

pub3 = t[:, 43].astype(np.float64)
inside = pub3[1:] - seen[1:]
hand = seen[1:] - pub3[:-1]
print(f"globaltimer: last x seen -> own x3 stored (ns): mean {inside.mean():.0f} median {np.median(inside):.0f}")
print(f"globaltimer: x3 stored by k-1 -> last x seen by k (ns): mean {hand.mean():.0f} median {np.median(hand):.0f} "
      f"p10 {np.percentile(hand, 10):.0f} p90 {np.percentile(hand, 90):.0f}")
print("hand-off by position (ns):", " ".join(f"{int(v)}" for v in hand[::8]))
print(f"timer granularity: {np.gcd.reduce(np.diff(np.unique(t[:, 13])).astype(np.int64))} ns")

ph = t[nb - 1, 48:52].astype(np.float64) / max(nb - 1, 1)
print(f"CTA {nb - 1}: fast-path blocks {int(t[nb - 1, 52])} of {nb - 1}")
print(f"CTA {nb - 1}: cycles per block iteration: x wait+check {ph[0]:.0f}, loads issued + panel widened {ph[1]:.0f}, "
      f"re-poll + first barrier {ph[2]:.0f}, FMAs {ph[3]:.0f}")
