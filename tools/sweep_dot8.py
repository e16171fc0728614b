"""DOT, fp64 storage: 8-byte loads (one element per lane and load, the request
shape of the reference's scalar kernel) against 16-byte loads and against the
reference kernel, same box, interleaved."""
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
x = torch.empty(nd, dtype=torch.float64, device=dev)
y = torch.empty(nd, dtype=torch.float64, device=dev)
h.fill_uniform(1, nd, x, nd, 42, 0)
h.fill_uniform(1, nd, y, nd, 42, nd)
for ar in (torch.float64, torch.float32):
    res = torch.zeros(1, dtype=ar, device=dev)
    nb = dot_bytes(nd, 8, res.element_size())
    best = {}
    vals = {}
    for rep in range(3):
        for vb, shapes in ((16, ((256, 4), (1024, 4))), (8, ((256, 4), (256, 8), (512, 4), (512, 8), (1024, 4), (1024, 8)))):
            for block, unroll in shapes:
                ab.tune("dot_vecbytes", vb)
                ab.tune("dot_block", block)
                ab.tune("dot_unroll", unroll)
                ms = min_of_10(lambda: h.dot(ar, nd, x, 1, y, 1, res), torch)
                key = (vb, block, unroll)
                best[key] = max(best.get(key, 0), nb / ms / 1e6)
                vals[key] = float(res.item())
        if refk is not None:
            ms = min_of_10(lambda: refk.dot(ar, nd, x, 1, y, 1, res), torch)
            best[("ref", 0, 0)] = max(best.get(("ref", 0, 0), 0), nb / ms / 1e6)
    print(f"dot fp64 storage, ar={ar}: " + "  ".join(f"v{k[0]}b{k[1]}u{k[2]}={v:.0f}" for k, v in sorted(best.items(), key=str)), flush=True)
    print("   values:", sorted(set(round(v, 6) for v in vals.values())), flush=True)
for key, val in (("dot_vecbytes", 16), ("dot_block", 0), ("dot_unroll", 0)):
    ab.tune(key, val)
