"""TRSV: cluster kernel (variant 0) against the one-CTA-per-block-row kernel
(variant 1) in ONE process -- results against a fp64 substitution on the
device, and interleaved timings (latency-bound kernels differ by tens of
percent between gpurun boxes, so versions are only compared inside one run).

    python tools/trsv_check.py [quick]
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import accessor_blas_b200 as ab  # noqa: E402

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
dev = torch.device("cuda:0")
NAME = {torch.float64: "fp64", torch.float32: "fp32", torch.float16: "fp16"}
h = ab.Handle(0)


def fixture(n, lda, st):
    T = torch.empty(n * lda, dtype=torch.float64, device=dev)
    b = torch.empty(n, dtype=torch.float64, device=dev)
    h.fill_uniform(n, lda, T, lda, 42, 0)
    h.fill_uniform(n, 1, b, 1, 42, n * lda)
    T.mul_(0.5 / max(n, 1) ** 0.5)
    T.view(n, lda)[:, :n].diagonal().fill_(1.25)
    return T.to(st), b.to(st)


def reference(T, n, lda, b, upper, unit):
    M = T.view(n, lda)[:, :n].double()
    M = torch.triu(M) if upper else torch.tril(M)
    if unit:
        M = M - torch.diag(torch.diagonal(M)) + torch.eye(n, dtype=torch.float64, device=dev)
    return torch.linalg.solve_triangular(M, b.double().unsqueeze(1), upper=upper).squeeze(1)


def run(variant, ar, uplo, diag, n, T, lda, b):
    ab.tune("trsv_variant", variant)
    x = b.clone()
    h.trsv(ar, uplo, diag, n, T, lda, x, 1)
    torch.cuda.synchronize()
    return x


print("== correctness (relative L1 error against fp64 substitution) ==", flush=True)
sizes = [1, 31, 128, 129, 300, 1000, 1024, 1153, 4096] + ([] if quick else [16384, 20000])
worst = 0.0
for n in sizes:
    for pad in (0, 1, 2):
        lda = n + pad
        for st in (torch.float32, torch.float64, torch.float16):
            T, b = fixture(n, lda, st)
            for ar in (torch.float64, torch.float32):
                for uplo, diag in ((ab.LOWER, ab.UNIT), (ab.UPPER, ab.NON_UNIT), (ab.LOWER, ab.NON_UNIT),
                                   (ab.UPPER, ab.UNIT)):
                    want = reference(T, n, lda, b, uplo == ab.UPPER, diag == ab.UNIT)
                    errs = []
                    xs = []
                    for variant in (0, 1):
                        x = run(variant, ar, uplo, diag, n, T, lda, b)
                        x2 = run(variant, ar, uplo, diag, n, T, lda, b)
                        assert torch.equal(x, x2), ("not reproducible", variant, n, pad, st, ar, uplo, diag)
                        xs.append(x)
                        errs.append(float((x.double() - want).abs().sum() / want.abs().sum()))
                    bar = {torch.float64: 1e-13, torch.float32: 2e-6, torch.float16: 2e-2}[st] \
                        if ar == torch.float64 else {torch.float64: 2e-5, torch.float32: 2e-5,
                                                     torch.float16: 2e-2}[st]
                    flag = "" if errs[0] <= max(2.0 * errs[1], bar) else "   <-- WORSE"
                    if flag or pad == 0 and uplo == ab.LOWER and diag == ab.UNIT:
                        print(f"n={n:6d} lda+{pad} Acc<{NAME[ar]},{NAME[st]}> uplo={uplo} diag={diag}: "
                              f"cluster {errs[0]:.3e}  single {errs[1]:.3e}{flag}", flush=True)
                    worst = max(worst, errs[0] / max(errs[1], 1e-300))
                    assert not flag
            del T
print(f"worst cluster/single error ratio: {worst:.3f}", flush=True)


def timed(variant, ar, st, uplo, diag, n, T, lda, b):
    ab.tune("trsv_variant", variant)
    x = b.clone()
    best = 1e9
    for _ in range(8):
        x.copy_(b)
        torch.cuda._sleep(40_000)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        h.trsv(ar, uplo, diag, n, T, lda, x, 1)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3


print("== timings (us, min of 8, interleaved) ==", flush=True)
for n in ([4096, 16384] if quick else [1024, 4096, 16384, 32768]):
    for st in (torch.float32, torch.float64, torch.float16):
        T, b = fixture(n, n, st)
        for ar in (torch.float64, torch.float32):
            for uplo, diag in ((ab.LOWER, ab.UNIT), (ab.UPPER, ab.NON_UNIT)):
                res = {0: 1e9, 1: 1e9}
                for rep in range(2):
                    for variant in (0, 1):
                        res[variant] = min(res[variant], timed(variant, ar, st, uplo, diag, n, T, n, b))
                print(f"n={n:6d} Acc<{NAME[ar]},{NAME[st]}> {'lower' if uplo == ab.LOWER else 'upper'}/"
                      f"{'unit' if diag == ab.UNIT else 'nonunit'}: cluster {res[0]:7.1f}  single {res[1]:7.1f}  "
                      f"({res[1] / res[0]:.2f}x)", flush=True)
        del T
ab.tune("trsv_variant", -1)
