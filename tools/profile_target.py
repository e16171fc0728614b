"""Short program for `ncu`: one launch of each hot kernel at the BASELINE
sizes after a warm-up launch (so -k regex + -s can pick the warm ones)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import accessor_blas_b200 as ab  # noqa: E402

# usage: profile_target.py [gemv] [dot] [trsv] [key=value ...]   (key=value -> accblas_tune)
which = {a for a in sys.argv[1:] if "=" not in a} or {"gemv", "dot", "trsv", "fixtures"}
for a in sys.argv[1:]:
    if "=" in a:
        key, value = a.split("=")
        ab.tune(key, int(value))
dev = torch.device("cuda:0")
h = ab.Handle(0)
f64, f32, f16 = torch.float64, torch.float32, torch.float16

if "gemv" in which:
    m = n = 16384
    for st in (f32, f16, f64):
        A = torch.empty(m * n, dtype=st, device=dev)
        x = torch.empty(n, dtype=st, device=dev)
        y = torch.zeros(m, dtype=st, device=dev)
        h.fill_uniform(m, n, A, n, 42, 0)
        h.fill_uniform(n, 1, x, 1, 42, m * n)
        for ar in (f64, f32):
            for _ in range(2):
                h.gemv(ar, m, n, 1.0, A, n, x, 1, 1.0, y, 1)
        torch.cuda.synchronize()
        del A

if "dot" in which:
    n = 2 ** 28
    for st in (f32, f16, f64):
        x = torch.empty(n, dtype=st, device=dev)
        y = torch.empty(n, dtype=st, device=dev)
        h.fill_uniform(1, n, x, n, 42, 0)
        h.fill_uniform(1, n, y, n, 42, n)
        for ar in (f64, f32):
            res = torch.zeros(1, dtype=ar, device=dev)
            for _ in range(2):
                h.dot(ar, n, x, 1, y, 1, res)
        torch.cuda.synchronize()
        del x, y

if "trsv" in which:
    n = 16384
    g = torch.empty(n * n, dtype=torch.float64, device=dev)
    h.fill_uniform(n, n, g, n, 42, 0)
    LU, _ = torch.linalg.lu_factor(g.view(n, n))
    del g
    import os
    pairs = os.environ.get("ACCBLAS_PROFILE_TRSV", "f64:f32").split(",")
    DT = {"f64": f64, "f32": f32, "f16": f16}
    for pair in pairs:
        ar_s, st_s = pair.split(":")
        A = LU.contiguous().view(-1).to(DT[st_s])
        b = torch.empty(n, dtype=DT[st_s], device=dev)
        h.fill_uniform(n, 1, b, 1, 42, n * n)
        for _ in range(2):
            x = b.clone()
            h.trsv(DT[ar_s], ab.LOWER, ab.UNIT, n, A, n, x, 1)
        torch.cuda.synchronize()
    del LU
if "fixtures" in which:
    cnt = 2 ** 28
    src = torch.empty(cnt, dtype=f64, device=dev)
    for _ in range(2):
        h.fill_uniform(1, cnt, src, cnt, 42, 0)
    for st in (f32, f16):
        out = torch.empty(cnt, dtype=st, device=dev)
        for _ in range(2):
            h.fill_uniform(1, cnt, out, cnt, 42, 0)
            h.convert(1, cnt, src, cnt, out, cnt)
        torch.cuda.synchronize()
        del out
    del src
print("profile target done")
