"""Per-phase timeline of the cluster TRSV kernel (development tool; run under
gpurun with the development library):

    python accessor-blas_b200/build.py --dev
    ACCBLAS_LIB=accessor-blas_b200/libaccblas_b200_dev.so python tools/trsv_trace.py [n] [st] [ar]

Slots per block row k (SM clock unless noted): 0 start | 1 diagonal tile in
shared memory | 2 sub-blocks inverted, products formed | 3 panels left of the
last one streamed | 4 last x sub-block consumed | 6 right-hand side complete |
10 solved | 12 globaltimer (ns) at the end | 13 globaltimer when the last x
sub-block had been consumed | 40..43 globaltimer when sub-block 0..3 was
published.
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
ar_code = {"f64": 0, "f32": 1}[sys.argv[3] if len(sys.argv) > 3 else "f64"]
dev = torch.device("cuda:0")
h = ab.Handle(0)
import os
if "TRSV_VARIANT" in os.environ:   # 1: the single-CTA kernel (its phase sums are printed below)
    ab.tune("trsv_variant", int(os.environ["TRSV_VARIANT"]))
lib = capi.load()
fn = lib.accblas_dev_trsv_trace   # AttributeError: not the development library
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
    assert rc == 0, capi.load().accblas_last_error()
    print(f"run {it}: {e0.elapsed_time(e1) * 1e3:.1f} us total")
t = trace.cpu().numpy().reshape(nb, 64)
if os.environ.get("TRSV_VARIANT") == "1":
    ph = t[nb - 1, 48:53]
    per = ph[:4] / max(nb - 1, 1)
    print(f"single-CTA kernel, last block row, sums over its {nb - 1} block iterations (cycles; per iteration): "
          f"x wait+check {ph[0]} ({per[0]:.0f})  loads issued+widen {ph[1]} ({per[1]:.0f})  "
          f"re-poll+first barrier {ph[2]} ({per[2]:.0f})  FMAs {ph[3]} ({per[3]:.0f})  fast-path iterations {ph[4]}")
for k in sorted(set([0, 1, 2, 7, 8, 9, nb // 4, nb // 2, nb - 2, nb - 1])):
    if 0 <= k < nb:
        print(f"block {k:4d}: tile {t[k, 1] - t[k, 0]:6d}  invert+products {t[k, 2] - t[k, 1]:6d}  "
              f"stream {t[k, 3] - t[k, 2]:8d}  wait+last panel {t[k, 4] - t[k, 3]:8d}  "
              f"rhs {t[k, 6] - t[k, 4]:5d}  solve {t[k, 10] - t[k, 6]:5d}   (cycles)")
ends = t[:, 12].astype(np.float64)
seen = t[:, 13].astype(np.float64)
pub = t[:, 40:44].astype(np.float64)
step = np.diff(ends)
print(f"end-to-end per block step (ns): median {np.median(step):.0f} mean {step.mean():.0f} "
      f"p10 {np.percentile(step, 10):.0f} p90 {np.percentile(step, 90):.0f}; chain {(ends.max() - ends[0]) / 1e3:.1f} us")
hand = seen[1:] - pub[:-1, 3]
inside = pub[1:, 3] - seen[1:]
k = np.arange(1, nb)
for name, sel in (("inside a cluster", k % 8 != 0), ("across clusters", k % 8 == 0)):
    if sel.any():
        print(f"hand-off {name} (x3 published by k-1 -> consumed by k, ns): median {np.median(hand[sel]):.0f} "
              f"p10 {np.percentile(hand[sel], 10):.0f} p90 {np.percentile(hand[sel], 90):.0f}")
print(f"inside a CTA (last x consumed -> own x3 published, ns): median {np.median(inside):.0f} "
      f"p10 {np.percentile(inside, 10):.0f} p90 {np.percentile(inside, 90):.0f}")
got = t[:, 44:48].astype(np.float64)
lat = got[1:] - pub[:-1]
for name, sel in (("inside a cluster", k % 8 != 0), ("across clusters", k % 8 == 0)):
    if sel.any():
        print(f"sub-block published by k-1 -> barrier passed by k, {name} (ns): median "
              f"{np.median(lat[sel], axis=0)}  p90 {np.percentile(lat[sel], 90, axis=0)}")
steps_by_rank = [np.median(step[(k % 8) == r]) for r in range(8)]
print("median step (ns) by rank in the cluster:", " ".join(f"{v:.0f}" for v in steps_by_rank))
# one SM, one clock: last x consumed (slot 4) -> sub-block solutions published
base = t[1:, 4].astype(np.float64)
marks = []
for st_ in range(4):
    marks.append(np.median(t[1:, 16 + 2 * st_] - base))
    marks.append(np.median(t[1:, 17 + 2 * st_] - base))
print("cycles after the last x was consumed (median): rhs complete +%d; " % np.median(t[1:, 6] - base) +
      "; ".join(f"x{i}: begin +{int(marks[2 * i])}, published +{int(marks[2 * i + 1])}" for i in range(4)) +
      f"; all done +{int(np.median(t[1:, 10] - base))}")
for kk in (33, 34, 35, 36):
    if kk < nb:
        b0 = t[kk, 0]
        print(f"raw clock64 of block {kk} relative to its start: " +
              " ".join(f"s{sl}={int(t[kk, sl] - b0)}" for sl in (1, 2, 3, 4, 6, 16, 17, 18, 19, 20, 21, 22, 23, 10)))
        print(f"raw globaltimer of block {kk} relative to slot 13: " +
              " ".join(f"s{sl}={int(t[kk, sl] - t[kk, 13])}" for sl in (44, 45, 46, 47, 13, 40, 41, 42, 43, 12)))
dg = (t[1:, 12] - t[1:, 13]).astype(np.float64)
dc = (t[1:, 10] - t[1:, 4]).astype(np.float64)
print(f"same interval by both timers: clock64 {np.median(dc):.0f} cycles, globaltimer {np.median(dg):.0f} ns "
      f"-> {np.median(dc) / max(np.median(dg), 1):.2f} GHz")
links = np.diff(pub[1:], axis=1)
print(f"chain links x0->x1->x2->x3 (ns): median {np.median(links, axis=0)}")
print(f"timer granularity: {np.gcd.reduce(np.diff(np.unique(t[:, 13])).astype(np.int64))} ns")
