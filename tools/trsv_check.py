"""TRSV variants in ONE process -- results against a fp64 substitution on the
device, and interleaved timings (latency-bound kernels differ by tens of
percent between gpurun boxes, so versions are only compared inside one run).

  cluster   thread-block clusters, hand-off through distributed shared memory
  single    one CTA per block row, hand-off through L2

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
VARIANTS = {"cluster": 0, "single": 1}


def select(name):
    ab.tune("trsv_variant", VARIANTS[name])


def fixture(n, lda, st):
    T = torch.empty(n * lda, dtype=torch.float64, device=dev)
    b = torch.empty(n, dtype=torch.float64, device=dev)
    h.fill_uniform(n, lda, T, lda, 42, 0)
    h.fill_uniform(n, 1, b, 1, 42, n * lda)
    T.mul_(0.5 / max(n, 1) ** 0.5)
    T.view(n, lda)[:, :n].diagonal().fill_(1.25)
    return T.to(st), b.to(st)


def lu_fixture(n, st, transposed):
    g = torch.empty(n * n, dtype=torch.float64, device=dev)
    h.fill_uniform(n, n, g, n, 42, 0)
    LU, _ = torch.linalg.lu_factor(g.view(n, n))
    del g
    M = (LU.t().contiguous() if transposed else LU.contiguous()).view(-1)
    b = torch.empty(n, dtype=torch.float64, device=dev)
    h.fill_uniform(n, 1, b, 1, 42, n * n)
    return M.to(st), b.to(st)


def reference(T, n, lda, b, upper, unit):
    M = T.view(n, lda)[:, :n].double()
    M = torch.triu(M) if upper else torch.tril(M)
    if unit:
        M = M - torch.diag(torch.diagonal(M)) + torch.eye(n, dtype=torch.float64, device=dev)
    return torch.linalg.solve_triangular(M, b.double().unsqueeze(1), upper=upper).squeeze(1)


def run(name, ar, uplo, diag, n, T, lda, b):
    select(name)
    x = b.clone()
    h.trsv(ar, uplo, diag, n, T, lda, x, 1)
    torch.cuda.synchronize()
    return x


def check(label, n, lda, T, b, ar, st, uplo, diag, show):
    want = reference(T, n, lda, b, uplo == ab.UPPER, diag == ab.UNIT)
    errs = {}
    for name in VARIANTS:
        x = run(name, ar, uplo, diag, n, T, lda, b)
        x2 = run(name, ar, uplo, diag, n, T, lda, b)
        assert torch.equal(x, x2), ("not reproducible", name, n, st, ar, uplo, diag)
        errs[name] = float((x.double() - want).abs().sum() / want.abs().sum())
    bar = {torch.float64: 1e-13, torch.float32: 2e-6, torch.float16: 2e-2}[st] \
        if ar == torch.float64 else {torch.float64: 2e-5, torch.float32: 2e-5, torch.float16: 2e-2}[st]
    worse = [nm for nm in ("cluster",) if errs[nm] > max(2.0 * errs["single"], bar)]
    if show or worse:
        print(f"{label} n={n:6d} Acc<{NAME[ar]},{NAME[st]}> uplo={uplo} diag={diag}: " +
              "  ".join(f"{nm} {e:.3e}" for nm, e in errs.items()) +
              ("   <-- WORSE: " + ",".join(worse) if worse else ""), flush=True)
    return errs, worse


print("== correctness (relative L1 error against fp64 substitution) ==", flush=True)
bad = []
sizes = [1, 31, 128, 129, 300, 1000, 1024, 1153, 4096] + ([] if quick else [16384, 20000])
for n in sizes:
    for pad in (0, 1, 2):
        lda = n + pad
        for st in (torch.float32, torch.float64, torch.float16):
            T, b = fixture(n, lda, st)
            for ar in (torch.float64, torch.float32):
                for uplo, diag in ((ab.LOWER, ab.UNIT), (ab.UPPER, ab.NON_UNIT), (ab.LOWER, ab.NON_UNIT),
                                   (ab.UPPER, ab.UNIT)):
                    _, worse = check("scaled", n, lda, T, b, ar, st, uplo, diag,
                                     pad == 0 and uplo == ab.LOWER and diag == ab.UNIT and n in (300, 4096))
                    bad += worse
            del T
# the LU factors of the uniform(-1,1) fixture: L (unit, well conditioned), its
# transpose, and U^T (non-unit, conditioned like the random matrix itself)
for n in ([1000, 4096] if quick else [1000, 4096, 16384]):
    for st in (torch.float32, torch.float64):
        for transposed, uplo, diag in ((False, ab.LOWER, ab.UNIT), (True, ab.UPPER, ab.UNIT),
                                       (True, ab.LOWER, ab.NON_UNIT), (False, ab.UPPER, ab.NON_UNIT)):
            T, b = lu_fixture(n, st, transposed)
            for ar in (torch.float64, torch.float32):
                _, worse = check("LU", n, n, T, b, ar, st, uplo, diag, True)
                bad += worse
            del T
print("WORSE cases:", bad, flush=True)


def timed(name, ar, st, uplo, diag, n, T, lda, b):
    select(name)
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


print("== timings (us, min of 8, interleaved), LU factors of the uniform(-1,1) fixture ==", flush=True)
for n in ([4096, 16384] if quick else [1024, 4096, 16384, 32768]):
    for st in (torch.float32, torch.float64, torch.float16):
        for transposed, uplo, diag, tri in ((False, ab.LOWER, ab.UNIT, "lower/unit L"),
                                            (True, ab.LOWER, ab.NON_UNIT, "lower/nonunit U^T")):
            T, b = lu_fixture(n, st, transposed)
            for ar in (torch.float64, torch.float32):
                res = {nm: 1e9 for nm in VARIANTS}
                for rep in range(2):
                    for nm in VARIANTS:
                        res[nm] = min(res[nm], timed(nm, ar, st, uplo, diag, n, T, n, b))
                print(f"n={n:6d} Acc<{NAME[ar]},{NAME[st]}> {tri}: " +
                      "  ".join(f"{nm} {v:7.1f}" for nm, v in res.items()), flush=True)
            del T
ab.tune("trsv_variant", -1)
