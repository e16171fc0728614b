"""Generates tests/golden/reference_kernels_b200.npz: outputs of the REFERENCE'S
OWN CUDA kernels (oracle/_ref/libref_kernels.so, compiled from
/root/reference/cuda) on a B200, for small seeded inputs that the CPU test
suite can regenerate (inputs are the closed-form uniform(-1,1) stream, so only
seeds / sizes / outputs are stored).

Run on a GPU box from the repo root:
    python tests/golden/make_reference_golden.py        # writes gpurun_out/reference_kernels_b200.npz
then copy the file to tests/golden/.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle_binding import Oracle, RefKernels  # noqa: E402

NP = {torch.float64: np.float64, torch.float32: np.float32}
TAG = {torch.float64: "f64", torch.float32: "f32"}
orc, ref = Oracle(), RefKernels()
dev = "cuda:0"
out = {"sm_count": np.array([ref.sm_count()])}


def stored(count, st, seed, first=0):
    return orc.convert(orc.uniform(count, seed=seed, first_draw=first), NP[st])


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


# GEMV: (m, n, lda), seed 42, draw order matrix / x / res, alpha = beta = 1
for m, n, lda in ((100, 100, 24500), (257, 1000, 1000), (64, 3001, 3008)):
    for ar, st, plain in ((torch.float64, torch.float64, True), (torch.float32, torch.float32, True),
                          (torch.float64, torch.float32, False), (torch.float32, torch.float32, False),
                          (torch.float64, torch.float64, False)):
        A = stored(m * lda, st, 42)
        x = stored(n, st, 42, m * lda)
        y = stored(m, st, 42, m * lda + n)
        yd = to_dev(y)
        ref.gemv(ar, m, n, 1.0, to_dev(A), lda, to_dev(x), 1, 1.0, yd, 1, plain=plain)
        ref.sync()
        out[f"gemv_{m}_{n}_{lda}_{TAG[ar]}_{TAG[st]}_{int(plain)}"] = yd.cpu().numpy()

# DOT: n, seed 42, x then y
for n in (1000, 1_000_000):
    for ar, st, res_t, plain in ((torch.float64, torch.float64, torch.float64, True),
                                 (torch.float64, torch.float32, torch.float32, False),
                                 (torch.float64, torch.float32, torch.float64, False),
                                 (torch.float32, torch.float32, torch.float32, False)):
        x = stored(n, st, 42)
        y = stored(n, st, 42, n)
        res = torch.full((1,), -999.0, dtype=res_t, device=dev)
        ref.dot(ar, n, to_dev(x), 1, to_dev(y), 1, res, plain=plain)
        ref.sync()
        out[f"dot_{n}_{TAG[ar]}_{TAG[st]}_{TAG[res_t]}_{int(plain)}"] = res.cpu().numpy()

# TRSV: well-conditioned triangles built on the CPU (reproducible without a GPU):
# T = I + 0.02 * uniform (strictly triangular part), diagonal 1 + 0.5 * uniform
for n in (100, 517):
    base = orc.uniform(n * n, seed=7).reshape(n, n) * 0.02
    diag = 1.0 + 0.5 * orc.uniform(n, seed=8)
    T = base.copy()
    np.fill_diagonal(T, diag)
    b = orc.uniform(n, seed=9)
    for upper in (False, True):
        for unit in (False, True):
            for ar, st, plain in ((torch.float64, torch.float64, True),
                                  (torch.float64, torch.float32, False),
                                  (torch.float32, torch.float32, False)):
                A = orc.convert(T.reshape(-1), NP[st])
                xs = orc.convert(b, NP[st])
                xd = to_dev(xs)
                ref.trsv(ar, upper, unit, n, to_dev(A), n, xd, 1, plain=plain)
                ref.sync()
                out[f"trsv_{n}_{int(upper)}_{int(unit)}_{TAG[ar]}_{TAG[st]}_{int(plain)}"] = \
                    xd.cpu().numpy()

dst = ROOT / "gpurun_out"
dst.mkdir(exist_ok=True)
np.savez_compressed(dst / "reference_kernels_b200.npz", **out)
print("wrote", len(out), "arrays")
