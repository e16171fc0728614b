"""Same-GPU comparison the north-star asks for (run under gpurun):

  accblas kernels  vs  the reference's own CUDA kernels (oracle/_ref, compiled
  from /root/reference/cuda)  vs  cuBLAS,

on identical uniform(-1,1) inputs, with the reference's timing protocol
(1 warm-up, min of 10) and its error metric (relative to the plain fp64 kernel
of the reference).  Writes gpurun_out/compare_reference.md / .json.
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import accessor_blas_b200 as ab  # noqa: E402
from bench import dot_bytes, gemv_bytes, min_of_10, trsv_bytes  # noqa: E402
from oracle_binding import RefKernels  # noqa: E402

f64, f32, f16 = torch.float64, torch.float32, torch.float16
NAME = {f64: "fp64", f32: "fp32", f16: "fp16"}
dev = torch.device("cuda:0")
h = ab.Handle(0)
ref = RefKernels()
rows = []


def add(op, size, impl, pair, ms, nbytes, err):
    rows.append({"op": op, "size": size, "impl": impl, "pair": pair, "ms": ms,
                 "GBps": nbytes / ms / 1e6, "rel_error": err})
    print(f"{op:5s} {size:>10s} {impl:10s} {pair:16s} {ms:9.4f} ms "
          f"{nbytes / ms / 1e6:8.1f} GB/s  err {err:.3e}", flush=True)


# ------------------------------------------------------------------ GEMV 16384^2
m = n = 16384
A64 = torch.empty(m * n, dtype=f64, device=dev)
x64 = torch.empty(n, dtype=f64, device=dev)
y64 = torch.empty(m, dtype=f64, device=dev)
h.fill_uniform(m, n, A64, n, 42, 0)
h.fill_uniform(n, 1, x64, 1, 42, m * n)
h.fill_uniform(m, 1, y64, 1, 42, m * n + n)
yref = y64.clone()
ref.gemv(f64, m, n, 1.0, A64, n, x64, 1, 1.0, yref, 1, plain=True)   # the reference's reference
ref.sync()
for st in (f64, f32, f16):
    if st == f64:
        A, x, y0 = A64, x64, y64
    else:
        A = torch.empty(m * n, dtype=st, device=dev)
        x = torch.empty(n, dtype=st, device=dev)
        y0 = torch.empty(m, dtype=st, device=dev)
        h.convert(m, n, A64, n, A, n)
        h.convert(n, 1, x64, 1, x, 1)
        h.convert(m, 1, y64, 1, y0, 1)
    nb = gemv_bytes(m, n, A.element_size())
    for ar in (f64, f32):
        pair = f"Acc<{NAME[ar]},{NAME[st]}>"
        y = y0.clone()
        h.gemv(ar, m, n, 1.0, A, n, x, 1, 1.0, y, 1)
        err = h.l1_error(m, yref, 1, y, 1)
        ms = min_of_10(lambda: h.gemv(ar, m, n, 1.0, A, n, x, 1, 0.0, y, 1), torch)
        add("GEMV", "16384^2", "accblas", pair, ms, nb, err)
        y = y0.clone()
        ref.gemv(ar, m, n, 1.0, A, n, x, 1, 1.0, y, 1)
        ref.sync()
        err = h.l1_error(m, yref, 1, y, 1)
        ms = min_of_10(lambda: ref.gemv(ar, m, n, 1.0, A, n, x, 1, 0.0, y, 1), torch)
        add("GEMV", "16384^2", "reference", pair, ms, nb, err)
    if st != f16:
        y = y0.clone()
        ref.cublas_gemv(m, n, 1.0, A, n, x, 1, 1.0, y, 1)
        ref.sync()
        err = h.l1_error(m, yref, 1, y, 1)
        ms = min_of_10(lambda: ref.cublas_gemv(m, n, 1.0, A, n, x, 1, 0.0, y, 1), torch)
        add("GEMV", "16384^2", "cuBLAS", NAME[st], ms, nb, err)
    if st != f64:
        del A
del A64
torch.cuda.empty_cache()

# ------------------------------------------------------------------ DOT 2^28 and 2^20
for nd, label in ((2 ** 28, "2^28"), (2 ** 20, "2^20")):
    x64 = torch.empty(nd, dtype=f64, device=dev)
    y64 = torch.empty(nd, dtype=f64, device=dev)
    h.fill_uniform(1, nd, x64, nd, 42, 0)
    h.fill_uniform(1, nd, y64, nd, 42, nd)
    rres = torch.zeros(1, dtype=f64, device=dev)
    h.dot(f64, nd, x64, 1, y64, 1, rres)        # deterministic fp64 result as the yardstick
    ref_v = rres.item()
    for st in (f64, f32, f16):
        if st == f64:
            x, y = x64, y64
        else:
            x = torch.empty(nd, dtype=st, device=dev)
            y = torch.empty(nd, dtype=st, device=dev)
            h.convert(1, nd, x64, nd, x, nd)
            h.convert(1, nd, y64, nd, y, nd)
        for ar in (f64, f32):
            pair = f"Acc<{NAME[ar]},{NAME[st]}>"
            res_t = f32 if st == f32 else ar
            res = torch.zeros(1, dtype=res_t, device=dev)
            nb = dot_bytes(nd, x.element_size(), res.element_size())
            h.dot(ar, nd, x, 1, y, 1, res)
            err = abs(res.item() - ref_v) / abs(ref_v)
            ms = min_of_10(lambda: h.dot(ar, nd, x, 1, y, 1, res), torch)
            add("DOT", label, "accblas", pair, ms, nb, err)
            ref.dot(ar, nd, x, 1, y, 1, res)
            ref.sync()
            err = abs(res.item() - ref_v) / abs(ref_v)
            ms = min_of_10(lambda: ref.dot(ar, nd, x, 1, y, 1, res), torch)
            add("DOT", label, "reference", pair, ms, nb, err)
        if st != f16:
            res = torch.zeros(1, dtype=st, device=dev)
            nb = dot_bytes(nd, x.element_size(), res.element_size())
            ref.cublas_dot(nd, x, 1, y, 1, res)
            ref.sync()
            err = abs(res.item() - ref_v) / abs(ref_v)
            ms = min_of_10(lambda: ref.cublas_dot(nd, x, 1, y, 1, res), torch)
            add("DOT", label, "cuBLAS", NAME[st], ms, nb, err)
        if st != f64:
            del x, y
    del x64, y64
    torch.cuda.empty_cache()

# ------------------------------------------------------------------ TRSV 16384
nt = 16384
g = torch.empty(nt * nt, dtype=f64, device=dev)
h.fill_uniform(nt, nt, g, nt, 42, 0)
LU, _ = torch.linalg.lu_factor(g.view(nt, nt))
del g
b64 = torch.empty(nt, dtype=f64, device=dev)
h.fill_uniform(nt, 1, b64, 1, 42, nt * nt)
for tri, upper, unit, mat in (("lower/unit (L)", False, True, LU.contiguous().view(-1)),
                              ("upper/unit (L^T)", True, True, LU.t().contiguous().view(-1)),
                              ("lower/non-unit (U^T)", False, False, LU.t().contiguous().view(-1))):
    xref = b64.clone()
    ref.trsv(f64, upper, unit, nt, mat, nt, xref, 1, plain=True)
    ref.sync()
    A32 = mat.to(f32)
    b32 = b64.to(f32)
    for label, ar, A, b in (("Acc<fp64,fp64>", f64, mat, b64), ("Acc<fp64,fp32>", f64, A32, b32),
                            ("Acc<fp32,fp32>", f32, A32, b32)):
        nb = trsv_bytes(nt, A.element_size())
        for impl, call in (
                ("accblas", lambda xw: h.trsv(ar, ab.UPPER if upper else ab.LOWER,
                                              ab.UNIT if unit else ab.NON_UNIT, nt, A, nt, xw, 1)),
                ("reference", lambda xw: ref.trsv(ar, upper, unit, nt, A, nt, xw, 1))):
            xw = b.clone()
            call(xw)
            torch.cuda.synchronize()
            err = h.l1_error(nt, xref, 1, xw, 1)

            def timed():
                xw.copy_(b)
                call(xw)
            ms = min_of_10(timed, torch) - min_of_10(lambda: xw.copy_(b), torch)
            add("TRSV", f"16384 {tri}", impl, label, ms, nb, err)
    for label, A, b in (("fp64", mat, b64), ("fp32", A32, b32)):
        xw = b.clone()
        ref.cublas_trsv(upper, unit, nt, A, nt, xw, 1)
        torch.cuda.synchronize()
        err = h.l1_error(nt, xref, 1, xw, 1)

        def timed():
            xw.copy_(b)
            ref.cublas_trsv(upper, unit, nt, A, nt, xw, 1)
        ms = min_of_10(timed, torch) - min_of_10(lambda: xw.copy_(b), torch)
        add("TRSV", f"16384 {tri}", "cuBLAS", label, ms, trsv_bytes(nt, A.element_size()), err)

out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
tag = sys.argv[1] if len(sys.argv) > 1 else "compare_reference"
(out / f"{tag}.json").write_text(json.dumps(rows, indent=1))
lines = ["| op | size | implementation | pair | ms (min of 10) | GB/s | rel. error |",
         "|---|---|---|---|---|---|---|"]
for r in rows:
    lines.append(f"| {r['op']} | {r['size']} | {r['impl']} | {r['pair']} | {r['ms']:.4f} | "
                 f"{r['GBps']:.0f} | {r['rel_error']:.3e} |")
(out / f"{tag}.md").write_text("\n".join(lines) + "\n")
