"""Launch-shape sweep on a B200 (development tool; run under gpurun).

    python tools/tune.py [dot] [gemv] [trsv]

Prints GB/s (min of 10, the reference's timing protocol) for every storage /
arithmetic pair at the BASELINE sizes over the tunable launch parameters.
"""
import itertools
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import accessor_blas_b200 as ab  # noqa: E402
from bench import dot_bytes, gemv_bytes, min_of_10  # noqa: E402

NAME = {torch.float64: "fp64", torch.float32: "fp32", torch.float16: "fp16"}
which = set(sys.argv[1:]) or {"dot", "gemv"}
dev = torch.device("cuda:0")
h = ab.Handle(0)
results = {}

if "dot" in which:
    n = 2 ** 28
    for st in (torch.float32, torch.float16, torch.float64):
        x = torch.empty(n, dtype=st, device=dev)
        y = torch.empty(n, dtype=st, device=dev)
        h.fill_uniform(1, n, x, n, 42, 0)
        h.fill_uniform(1, n, y, n, 42, n)
        for ar in (torch.float64, torch.float32):
            res = torch.zeros(1, dtype=ar, device=dev)
            for unroll, cps in itertools.product((2, 4), (0, 2, 3, 4, 5, 6)):
                ab.tune("dot_unroll", unroll)
                ab.tune("dot_ctas_per_sm", cps)
                ms = min_of_10(lambda: h.dot(ar, n, x, 1, y, 1, res), torch)
                gbs = dot_bytes(n, x.element_size(), res.element_size()) / ms / 1e6
                key = f"dot Acc<{NAME[ar]},{NAME[st]}> unroll={unroll} ctas/sm={cps}"
                results[key] = round(gbs, 1)
                print(key, f"{gbs:8.1f} GB/s", flush=True)
        del x, y
    ab.tune("dot_unroll", 0)
    ab.tune("dot_ctas_per_sm", 0)

if "gemv" in which:
    # ACCBLAS_TUNE_GEMV="st:ar:unroll:variant:stages:intwords:taper[:pipe],..." restricts the sweep
    import os
    DT = {"f64": torch.float64, "f32": torch.float32, "f16": torch.float16}
    spec = os.environ.get("ACCBLAS_TUNE_GEMV", "")
    if spec:
        cases = []
        for item in spec.split(","):
            st_s, ar_s, *rest = item.split(":")
            cases.append((DT[st_s], DT[ar_s], *map(int, rest)))
    else:
        cases = [(st, ar, 2, v, 0, 3, 1)
                 for st in (torch.float32, torch.float16, torch.float64)
                 for ar in (torch.float64, torch.float32)
                 for v in (4, 5, 2, 3, 4)]
    m = k = 16384
    for st in (torch.float32, torch.float16, torch.float64):
        mine = [c for c in cases if c[0] == st]
        if not mine:
            continue
        A = torch.empty(m * k, dtype=st, device=dev)
        x = torch.empty(k, dtype=st, device=dev)
        y = torch.zeros(m, dtype=st, device=dev)
        h.fill_uniform(m, k, A, k, 42, 0)
        h.fill_uniform(k, 1, x, 1, 42, m * k)
        for _, ar, unroll, variant, stages, occ, taper, *more in mine:
            pipe = more[0] if more else -1
            ab.tune("gemv_pipe", pipe)
            ab.tune("gemv_unroll", unroll)
            ab.tune("gemv_variant", variant)
            ab.tune("gemv_stages", stages)
            ab.tune("gemv_intwords", occ)
            ab.tune("gemv_taper", taper)
            ms = min_of_10(lambda: h.gemv(ar, m, k, 1.0, A, k, x, 1, 0.0, y, 1), torch)
            gbs = gemv_bytes(m, k, A.element_size()) / ms / 1e6
            key = (f"gemv Acc<{NAME[ar]},{NAME[st]}> unroll={unroll} variant={variant} "
                   f"stages={stages} intwords={occ} taper={taper} pipe={pipe}")
            while key in results:
                key += " (repeat)"
            results[key] = round(gbs, 1)
            print(key, f"{gbs:8.1f} GB/s", flush=True)
        del A
    ab.tune("gemv_unroll", 2)
    ab.tune("gemv_variant", 0)
    ab.tune("gemv_stages", 0)
    ab.tune("gemv_intwords", 2)
    ab.tune("gemv_taper", 1)
    ab.tune("gemv_pipe", -1)

if "trsv" in which:
    # min-of-10 time of every (uplo, diag) at n = 16384 on a well-conditioned
    # triangle (scaled uniform entries, dominant diagonal)
    n = 16384
    for st in (torch.float32, torch.float64, torch.float16):
        T = torch.empty(n * n, dtype=st, device=dev)
        b = torch.empty(n, dtype=st, device=dev)
        h.fill_uniform(n, n, T, n, 42, 0)
        h.fill_uniform(n, 1, b, 1, 42, n * n)
        T.mul_(0.01)
        T.view(n, n).diagonal().fill_(1.0)
        import os
        aheads = [int(v) for v in os.environ.get("ACCBLAS_TUNE_TRSV_AHEAD", "3").split(",")]
        for ar, ahead in itertools.product((torch.float64, torch.float32), aheads):
            ab.tune("trsv_l2_ahead", ahead)
            for uplo, diag in ((ab.LOWER, ab.UNIT), (ab.UPPER, ab.NON_UNIT)):
                xw = b.clone()
                ms = min_of_10(lambda: h.trsv(ar, uplo, diag, n, T, n, xw, 1), torch)
                key = (f"trsv Acc<{NAME[ar]},{NAME[st]}> "
                       f"{'lower' if uplo == ab.LOWER else 'upper'}/"
                       f"{'unit' if diag == ab.UNIT else 'nonunit'} l2_ahead={ahead}")
                results[key] = round(ms * 1e3, 1)
                print(key, f"{ms * 1e3:8.1f} us", flush=True)
        del T

if "b2b" in which:
    # back-to-back launches on one stream (what bench.py's `value` and an
    # iterative solver see): programmatic dependent launch on / off
    m = k = 16384
    for st in (torch.float32, torch.float16, torch.float64):
        A = torch.empty(m * k, dtype=st, device=dev)
        x = torch.empty(k, dtype=st, device=dev)
        y = torch.zeros(m, dtype=st, device=dev)
        h.fill_uniform(m, k, A, k, 42, 0)
        h.fill_uniform(k, 1, x, 1, 42, m * k)
        for ar in (torch.float64, torch.float32):
            for pdl in (0, 1, 0, 1):
                ab.tune("gemv_pdl", pdl)
                for _ in range(5):
                    h.gemv(ar, m, k, 1.0, A, k, x, 1, 0.0, y, 1)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda._sleep(400_000)
                e0.record()
                for _ in range(50):
                    h.gemv(ar, m, k, 1.0, A, k, x, 1, 0.0, y, 1)
                e1.record()
                e1.synchronize()
                ms = e0.elapsed_time(e1) / 50
                gbs = gemv_bytes(m, k, A.element_size()) / ms / 1e6
                key = f"gemv b2b Acc<{NAME[ar]},{NAME[st]}> pdl={pdl}"
                while key in results:
                    key += " (repeat)"
                results[key] = round(gbs, 1)
                print(key, f"{ms * 1e3:8.1f} us/call {gbs:8.1f} GB/s", flush=True)
        del A
    ab.tune("gemv_pdl", 1)
    n = 2 ** 28
    for st in (torch.float32, torch.float16, torch.float64):
        x = torch.empty(n, dtype=st, device=dev)
        y = torch.empty(n, dtype=st, device=dev)
        h.fill_uniform(1, n, x, n, 42, 0)
        h.fill_uniform(1, n, y, n, 42, n)
        for ar in (torch.float64, torch.float32):
            res = torch.zeros(1, dtype=ar, device=dev)
            for pdl in (0, 1, 0, 1):
                ab.tune("dot_pdl", pdl)
                for _ in range(5):
                    h.dot(ar, n, x, 1, y, 1, res)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda._sleep(400_000)
                e0.record()
                for _ in range(50):
                    h.dot(ar, n, x, 1, y, 1, res)
                e1.record()
                e1.synchronize()
                ms = e0.elapsed_time(e1) / 50
                gbs = dot_bytes(n, x.element_size(), res.element_size()) / ms / 1e6
                key = f"dot b2b Acc<{NAME[ar]},{NAME[st]}> pdl={pdl}"
                while key in results:
                    key += " (repeat)"
                results[key] = round(gbs, 1)
                print(key, f"{ms * 1e3:8.1f} us/call {gbs:8.1f} GB/s", flush=True)
        del x, y
    ab.tune("dot_pdl", 1)

if "misaligned" in which:
    # rows that are only 8-byte aligned (the reference driver's stride 24500 for
    # fp16; 16384 + 4 here): register pipeline with 64-bit loads vs cp.async ring
    m = k = 16384
    for st, pad in ((torch.float16, 4), (torch.float16, 1), (torch.float32, 2), (torch.float32, 1), (torch.float64, 1)):
        lda = k + pad
        A = torch.empty(m * lda, dtype=st, device=dev)
        x = torch.empty(k, dtype=st, device=dev)
        y = torch.zeros(m, dtype=st, device=dev)
        h.fill_uniform(m, lda, A, lda, 42, 0)
        h.fill_uniform(k, 1, x, 1, 42, m * lda)
        for ar in (torch.float64, torch.float32):
            for pipe in ((1, 3, 1, 3) if (lda * A.element_size()) % 4 == 0 else (1, 1)):
                ab.tune("gemv_pipe", pipe)
                ms = min_of_10(lambda: h.gemv(ar, m, k, 1.0, A, lda, x, 1, 0.0, y, 1), torch)
                gbs = gemv_bytes(m, k, A.element_size()) / ms / 1e6
                key = f"gemv lda={lda} Acc<{NAME[ar]},{NAME[st]}> pipe={pipe}"
                while key in results:
                    key += " (repeat)"
                results[key] = round(gbs, 1)
                print(key, f"{gbs:8.1f} GB/s", flush=True)
        del A
    ab.tune("gemv_pipe", -1)
    # aligned data: default path vs the 64-bit-load pipeline
    m = k = 16384
    for st in (torch.float16, torch.float32, torch.float64):
        A = torch.empty(m * k, dtype=st, device=dev)
        x = torch.empty(k, dtype=st, device=dev)
        y = torch.zeros(m, dtype=st, device=dev)
        h.fill_uniform(m, k, A, k, 42, 0)
        h.fill_uniform(k, 1, x, 1, 42, m * k)
        for ar in (torch.float64, torch.float32):
            for force in (-1, 8, -1, 8):
                ab.tune("gemv_force_pieces", force)
                ms = min_of_10(lambda: h.gemv(ar, m, k, 1.0, A, k, x, 1, 0.0, y, 1), torch)
                gbs = gemv_bytes(m, k, A.element_size()) / ms / 1e6
                key = f"gemv aligned Acc<{NAME[ar]},{NAME[st]}> force_pieces={force}"
                while key in results:
                    key += " (repeat)"
                results[key] = round(gbs, 1)
                print(key, f"{gbs:8.1f} GB/s", flush=True)
        del A
    ab.tune("gemv_force_pieces", 0)

out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
(out / "tune.json").write_text(json.dumps(results, indent=1))
