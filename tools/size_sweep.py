"""Size sweep on one GPU: accblas against the reference's own CUDA kernels and
cuBLAS over the sizes the reference's drivers walk (not only the headline
points), min of 10, microseconds.  Flags every point where accblas is slower
than either.  Run under gpurun; writes gpurun_out/size_sweep.md."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import accessor_blas_b200 as ab  # noqa: E402
from bench import min_of_10  # noqa: E402
from oracle_binding import RefKernels  # noqa: E402

f64, f32, f16 = torch.float64, torch.float32, torch.float16
NAME = {f64: "fp64", f32: "fp32", f16: "fp16"}
dev = torch.device("cuda:0")
h = ab.Handle(0)
ref = RefKernels()
lines = ["| op | size | pair | accblas us | reference us | cuBLAS us | slower than |", "|---|---|---|---|---|---|---|"]
flagged = 0


def row(op, size, pair, t_acc, t_ref, t_cub):
    global flagged
    worse = [name for name, t in (("reference", t_ref), ("cuBLAS", t_cub)) if t is not None and t_acc > 1.02 * t]
    flagged += bool(worse)
    line = (f"| {op} | {size} | {pair} | {t_acc * 1e3:.1f} | {t_ref * 1e3:.1f} | "
            f"{'' if t_cub is None else f'{t_cub * 1e3:.1f}'} | {', '.join(worse)} |")
    lines.append(line)
    print(line, flush=True)


# ---------------------------------------------------------------- GEMV (square)
for n in (512, 1024, 2048, 4096, 8192, 12288, 24500):
    A64 = torch.empty(n * n, dtype=f64, device=dev)
    x64 = torch.empty(n, dtype=f64, device=dev)
    h.fill_uniform(n, n, A64, n, 42, 0)
    h.fill_uniform(n, 1, x64, 1, 42, n * n)
    for st in (f64, f32, f16):
        A, x = A64.to(st), x64.to(st)
        y = torch.zeros(n, dtype=st, device=dev)
        for ar in (f64, f32):
            t_acc = min_of_10(lambda: h.gemv(ar, n, n, 1.0, A, n, x, 1, 0.0, y, 1), torch)
            t_ref = min_of_10(lambda: ref.gemv(ar, n, n, 1.0, A, n, x, 1, 0.0, y, 1), torch)
            t_cub = None
            if ar == st and st != f16:
                t_cub = min_of_10(lambda: ref.cublas_gemv(n, n, 1.0, A, n, x, 1, 0.0, y, 1), torch)
            row("GEMV", f"{n}^2", f"Acc<{NAME[ar]},{NAME[st]}>", t_acc, t_ref, t_cub)
        del A, x, y
    del A64, x64

# ---------------------------------------------------------------- DOT
for e in (12, 16, 20, 22, 24, 26):
    n = 2 ** e
    x64 = torch.empty(n, dtype=f64, device=dev)
    y64 = torch.empty(n, dtype=f64, device=dev)
    h.fill_uniform(1, n, x64, n, 42, 0)
    h.fill_uniform(1, n, y64, n, 42, n)
    for st in (f64, f32, f16):
        x, y = x64.to(st), y64.to(st)
        for ar in (f64, f32):
            res = torch.zeros(1, dtype=ar, device=dev)
            t_acc = min_of_10(lambda: h.dot(ar, n, x, 1, y, 1, res), torch)
            t_ref = min_of_10(lambda: ref.dot(ar, n, x, 1, y, 1, res), torch)
            t_cub = None
            if ar == st and st != f16:
                t_cub = min_of_10(lambda: ref.cublas_dot(n, x, 1, y, 1, res), torch)
            row("DOT", f"2^{e}", f"Acc<{NAME[ar]},{NAME[st]}>", t_acc, t_ref, t_cub)
    del x64, y64

# ---------------------------------------------------------------- TRSV (L of LU, lower / unit)
for n in (256, 512, 1024, 2048, 4096, 8192):
    g = torch.empty(n * n, dtype=f64, device=dev)
    h.fill_uniform(n, n, g, n, 42, 0)
    LU, _ = torch.linalg.lu_factor(g.view(n, n))
    LU = LU.contiguous().view(-1)
    b64 = torch.empty(n, dtype=f64, device=dev)
    h.fill_uniform(n, 1, b64, 1, 42, n * n)
    for st in (f64, f32, f16):
        A, b = LU.to(st), b64.to(st)
        for ar in (f64, f32):
            xs = b.clone()

            def run_acc():
                h.trsv(ar, ab.LOWER, ab.UNIT, n, A, n, xs, 1)

            def run_ref():
                ref.trsv(ar, False, True, n, A, n, xs, 1)

            t_acc = min_of_10(run_acc, torch)
            t_ref = min_of_10(run_ref, torch)
            t_cub = None
            if ar == st and st != f16:
                t_cub = min_of_10(lambda: ref.cublas_trsv(False, True, n, A, n, xs, 1), torch)
            row("TRSV", str(n), f"Acc<{NAME[ar]},{NAME[st]}>", t_acc, t_ref, t_cub)
    del g, LU
lines.append(f"\n{flagged} points where accblas is more than 2 % slower than the reference kernel or cuBLAS.")
print(lines[-1])
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "size_sweep.md").write_text("\n".join(lines) + "\n")
