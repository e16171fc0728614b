"""Per-CTA timeline of the GEMV kernel (development tool; run under gpurun)."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import accessor_blas_b200 as ab  # noqa: E402
from accessor_blas_b200 import capi  # noqa: E402

st = {"f64": torch.float64, "f32": torch.float32, "f16": torch.float16}[
    sys.argv[1] if len(sys.argv) > 1 else "f32"]
ar = {"f64": torch.float64, "f32": torch.float32}[sys.argv[2] if len(sys.argv) > 2 else "f64"]
m = n = 16384
dev = torch.device("cuda:0")
h = ab.Handle(0)
lib = capi.load()
lib.accblas_dev_gemv_trace.argtypes = [ctypes.c_void_p]
A = torch.empty(m * n, dtype=st, device=dev)
x = torch.empty(n, dtype=st, device=dev)
y = torch.zeros(m, dtype=st, device=dev)
h.fill_uniform(m, n, A, n, 42, 0)
h.fill_uniform(n, 1, x, 1, 42, m * n)
cap = m
trace = torch.zeros(3 * cap, dtype=torch.int64, device=dev)
for _ in range(3):
    h.gemv(ar, m, n, 1.0, A, n, x, 1, 0.0, y, 1)
torch.cuda.synchronize()
assert lib.accblas_dev_gemv_trace(trace.data_ptr()) == 0
h.gemv(ar, m, n, 1.0, A, n, x, 1, 0.0, y, 1)
torch.cuda.synchronize()
assert lib.accblas_dev_gemv_trace(None) == 0
t = trace.cpu().numpy().reshape(cap, 3)
t = t[t[:, 0] > 0]
grid = t.shape[0]
t0 = t[:, 0].min()
start = (t[:, 0] - t0) / 1e3
end = (t[:, 1] - t0) / 1e3
dur = end - start
total = end.max()
print(f"kernel span {total:.1f} us, {grid} CTAs; CTA duration us: median {np.median(dur):.2f} "
      f"p10 {np.percentile(dur, 10):.2f} p90 {np.percentile(dur, 90):.2f} max {dur.max():.2f}")
first_wave = np.sort(start)[:444]
print(f"first 444 CTAs start within {first_wave.max():.2f} us")
edges = np.linspace(0, total, 41)
active = [(np.logical_and(start <= e, end > e)).sum() for e in edges]
print("resident CTAs over time:", " ".join(str(a) for a in active))
bytes_per_cta = m * n * A.element_size() / grid
done = np.sort(end)
for frac in (0.5, 0.9, 0.95, 0.98, 0.99, 1.0):
    i = int(frac * grid) - 1
    print(f"  {frac * 100:5.1f}% of CTAs finished at {done[i]:7.2f} us")
last_start = start.max()
print(f"last CTA started at {last_start:.2f} us; tail after that {total - last_start:.2f} us")
steady = grid * bytes_per_cta / 1e3 / total
print(f"effective {steady:.0f} GB/s; bytes/CTA {bytes_per_cta}")
mid = np.logical_and(end > 0.2 * total, end <= 0.8 * total).sum()
print(f"mid-span completion rate: {mid * bytes_per_cta / (0.6 * total) / 1e3:.0f} GB/s")
per_sm = np.bincount(t[:, 2].astype(int), minlength=148)
print(f"CTAs per SM: min {per_sm.min()} max {per_sm.max()}")
