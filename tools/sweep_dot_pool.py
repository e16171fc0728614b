"""DOT n = 2^28: share of the tiles handed out dynamically (dot_pool_pct) and
chunk size (dot_chunk_tiles) per pair, against the static partition (pool 0)
and the reference kernel; same box, interleaved, min of 10, GB/s."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import accessor_blas_b200 as ab  # noqa: E402
from bench import dot_bytes, min_of_10  # noqa: E402
from oracle_binding import REF_LIB, RefKernels  # noqa: E402

dev = torch.device("cuda:0")
h = ab.Handle(0)
refk = RefKernels() if REF_LIB.exists() else None
import os
nd = 2 ** int(os.environ.get("DOT_LOG2N", "28"))
NAME = {torch.float64: "fp64", torch.float32: "fp32", torch.float16: "fp16"}
src = torch.empty(2 * nd, dtype=torch.float64, device=dev)
h.fill_uniform(1, 2 * nd, src, 2 * nd, 42, 0)
for st in (torch.float64, torch.float32, torch.float16):
    x = src[:nd].to(st)
    y = src[nd:].to(st)
    for ar in (torch.float64, torch.float32):
        res = torch.zeros(1, dtype=ar, device=dev)
        nb = dot_bytes(nd, x.element_size(), res.element_size())
        best, vals = {}, {}
        configs = [(0, 4), (6, 4), (12, 2), (12, 4), (12, 8), (25, 4), (100, 4)]
        for rep in range(3):
            for pct, ct in configs:
                ab.tune("dot_pool_pct", pct)
                ab.tune("dot_chunk_tiles", ct)
                ms = min_of_10(lambda: h.dot(ar, nd, x, 1, y, 1, res), torch)
                key = f"p{pct}c{ct}"
                best[key] = max(best.get(key, 0), nb / ms / 1e6)
                vals.setdefault(key, set()).add(float(res.item()))
            if refk is not None:
                ms = min_of_10(lambda: refk.dot(ar, nd, x, 1, y, 1, res), torch)
                best["ref"] = max(best.get("ref", 0), nb / ms / 1e6)
        print(f"dot Acc<{NAME[ar]},{NAME[st]}>: " + "  ".join(f"{k}={v:.0f}" for k, v in best.items()), flush=True)
        print("   repeatable per config:", all(len(v) == 1 for v in vals.values()),
              " spread over configs:", max(max(v) for v in vals.values()) - min(min(v) for v in vals.values()), flush=True)
    del x, y
ab.tune("dot_pool_pct", 12)
ab.tune("dot_chunk_tiles", 4)

