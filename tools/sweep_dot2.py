"""DOT, fp64 / fp32 storage: does the reference's launch shape (allocating loads,
several waves of CTAs) explain its 2 % lead for fp64 storage?  Same box,
interleaved, reference kernel next to it."""
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
nd = 2 ** 28
for st, ar in ((torch.float64, torch.float64), (torch.float64, torch.float32), (torch.float32, torch.float64)):
    x = torch.empty(nd, dtype=st, device=dev)
    y = torch.empty(nd, dtype=st, device=dev)
    h.fill_uniform(1, nd, x, nd, 42, 0)
    h.fill_uniform(1, nd, y, nd, 42, nd)
    res = torch.zeros(1, dtype=ar, device=dev)
    nb = dot_bytes(nd, x.element_size(), res.element_size())
    best = {}
    for rep in range(3):
        for block in (256, 1024):
            for l1 in (0, 1):
                for cps in (0, 4, 8, 16, 32):
                    if cps and block == 1024 and cps > 16:
                        continue
                    ab.tune("dot_block", block)
                    ab.tune("dot_unroll", 4)
                    ab.tune("dot_l1", l1)
                    ab.tune("dot_waves", 1 if cps else 0)
                    ab.tune("dot_ctas_per_sm", cps)
                    ms = min_of_10(lambda: h.dot(ar, nd, x, 1, y, 1, res), torch)
                    key = (block, l1, cps)
                    best[key] = max(best.get(key, 0), nb / ms / 1e6)
        if refk is not None:
            ms = min_of_10(lambda: refk.dot(ar, nd, x, 1, y, 1, res), torch)
            best[("ref", 0, 0)] = max(best.get(("ref", 0, 0), 0), nb / ms / 1e6)
    print(f"dot ar={ar} st={st}: " + "  ".join(f"b{k[0]}l{k[1]}c{k[2]}={v:.0f}" for k, v in sorted(best.items(), key=str)),
          flush=True)
    del x, y
for key in ("dot_block", "dot_unroll", "dot_l1", "dot_waves", "dot_ctas_per_sm"):
    ab.tune(key, 0)
