"""Same-box sweeps (run under gpurun): DOT launch shape per pair, the
integer-widening experiment, fill_uniform / convert throughput.

    python tools/sweep_dot_fill.py
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import accessor_blas_b200 as ab  # noqa: E402
from bench import dot_bytes, min_of_10  # noqa: E402

dev = torch.device("cuda:0")
NAME = {torch.float64: "fp64", torch.float32: "fp32", torch.float16: "fp16"}
h = ab.Handle(0)

print("== fill_uniform / convert (GB/s written / read+written, min of 10) ==", flush=True)
cnt = 2 ** 28
for st in (torch.float64, torch.float32, torch.float16):
    out = torch.empty(cnt, dtype=st, device=dev)
    for generic in (0, 1):
        ab.tune("fill_generic", generic)
        ms = min_of_10(lambda: h.fill_uniform(1, cnt, out, cnt, 42, 12345), torch)
        print(f"fill {NAME[st]} {'per-row kernel' if generic else 'linear kernel'}: {ms:8.3f} ms "
              f"{cnt * out.element_size() / ms / 1e6:8.1f} GB/s", flush=True)
    ab.tune("fill_generic", 0)
    del out
src = torch.empty(cnt, dtype=torch.float64, device=dev)
h.fill_uniform(1, cnt, src, cnt, 42, 0)
for st in (torch.float32, torch.float16):
    out = torch.empty(cnt, dtype=st, device=dev)
    ms = min_of_10(lambda: h.convert(1, cnt, src, cnt, out, cnt), torch)
    print(f"convert fp64 -> {NAME[st]}: {ms:8.3f} ms {cnt * (8 + out.element_size()) / ms / 1e6:8.1f} GB/s",
          flush=True)
    del out
del src
torch.cuda.empty_cache()

print("== GEMV 16384^2 fp16 storage, row stride 16384 + pad elements (GB/s, min of 10) ==", flush=True)
from bench import gemv_bytes  # noqa: E402
mg = ng = 16384
for pad in (0, 4, 2, 1):
    lda = ng + pad
    Ah = torch.empty(mg * lda, dtype=torch.float16, device=dev)
    xh = torch.empty(ng, dtype=torch.float16, device=dev)
    yh = torch.zeros(mg, dtype=torch.float16, device=dev)
    h.fill_uniform(mg, lda, Ah, lda, 42, 0)
    h.fill_uniform(ng, 1, xh, 1, 42, mg * lda)
    for ar in (torch.float64, torch.float32):
        ms = min_of_10(lambda: h.gemv(ar, mg, ng, 1.0, Ah, lda, xh, 1, 0.0, yh, 1), torch)
        print(f"gemv Acc<{NAME[ar]},fp16> lda = n + {pad}: {gemv_bytes(mg, ng, 2) / ms / 1e6:7.0f}", flush=True)
    del Ah, xh, yh
torch.cuda.empty_cache()

print("== DOT n = 2^28: block x unroll per pair (GB/s) ==", flush=True)
nd = 2 ** 28
x64 = torch.empty(nd, dtype=torch.float64, device=dev)
y64 = torch.empty(nd, dtype=torch.float64, device=dev)
h.fill_uniform(1, nd, x64, nd, 42, 0)
h.fill_uniform(1, nd, y64, nd, 42, nd)
for st in (torch.float64, torch.float32, torch.float16):
    if st == torch.float64:
        x, y = x64, y64
    else:
        x = torch.empty(nd, dtype=st, device=dev)
        y = torch.empty(nd, dtype=st, device=dev)
        h.convert(1, nd, x64, nd, x, nd)
        h.convert(1, nd, y64, nd, y, nd)
    for ar in (torch.float64, torch.float32):
        res = torch.zeros(1, dtype=ar, device=dev)
        row = []
        for rep in range(2):
            for block in (256, 512, 1024):
                for unroll in (1, 2, 4):
                    mixes = (0, 1) if (st == torch.float32 and ar == torch.float64) else (0,)
                    for mix in mixes:
                        ab.tune("dot_block", block)
                        ab.tune("dot_unroll", unroll)
                        ab.tune("dot_intmix", mix)
                        ms = min_of_10(lambda: h.dot(ar, nd, x, 1, y, 1, res), torch)
                        row.append((block, unroll, mix, dot_bytes(nd, x.element_size(), res.element_size()) / ms / 1e6))
        best = {}
        for block, unroll, mix, g in row:
            best[(block, unroll, mix)] = max(best.get((block, unroll, mix), 0.0), g)
        print(f"dot Acc<{NAME[ar]},{NAME[st]}>: " + "  ".join(
            f"b{b}u{u}{'m' if m else ''}={g:6.0f}" for (b, u, m), g in sorted(best.items())), flush=True)
    if st != torch.float64:
        del x, y
ab.tune("dot_block", 0)
ab.tune("dot_unroll", 0)
ab.tune("dot_intmix", 0)

print("== DOT operand layouts, Acc<fp64,fp32>, n = 2^26 (GB/s) ==", flush=True)
n2 = 2 ** 26
base = torch.empty(2 * n2 + 64, dtype=torch.float32, device=dev)
h.fill_uniform(1, base.numel(), base, base.numel(), 42, 0)
res = torch.zeros(1, dtype=torch.float64, device=dev)
for label, xo, yo, inc in (("aligned", 0, 0, 1), ("same misalignment", 1, 1, 1), ("x+1 y+2", 1, 2, 1),
                           ("x+0 y+3", 0, 3, 1), ("x+2 y+0", 2, 0, 1), ("stride 2", 0, 0, 2)):
    cnt2 = n2 // inc
    xv = base[xo:]
    yv = base[n2 + 32 + yo:]
    ms = min_of_10(lambda: h.dot(torch.float64, cnt2, xv, inc, yv, inc, res), torch)
    print(f"dot {label:20s}: {ms * 1e3:8.1f} us {dot_bytes(cnt2, 4, 8) / ms / 1e6:8.1f} GB/s (algorithmic)", flush=True)
