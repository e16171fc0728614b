"""Generates tests/golden/fp16_cuda_host.npz: binary16 conversions computed by
CUDA's own cuda_fp16.h HOST routines (oracle/fp16_pin.cu, compiled with nvcc,
runs without a GPU).  Run from the repo root in the build container:
    python tests/golden/make_fp16_golden.py
"""
import subprocess
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]


def inputs() -> np.ndarray:
    rng = np.random.default_rng(1234)
    vals = [rng.uniform(-1, 1, 20000), rng.uniform(-70000, 70000, 4000),
            rng.uniform(-1e-4, 1e-4, 4000), rng.uniform(-7e-8, 7e-8, 2000)]
    # every half value, its neighbours' midpoints (ties) and a hair either side
    h = np.arange(0, 0x7c00, dtype=np.uint16).view(np.float16).astype(np.float64)
    mid = (h[:-1] + h[1:]) / 2
    vals += [h, -h, mid, -mid, np.nextafter(mid, np.inf), np.nextafter(mid, -np.inf)]
    # the probe of SURVEY.md section 7: double rounding would give 0x3c00
    vals.append(np.array([1 + 2.0 ** -11 + 2.0 ** -30, 65504.0, 65519.99, 65520.0,
                          1e6, -1e6, 2.0 ** -24, 2.0 ** -25, 2.0 ** -25 * (1 + 2 ** -20),
                          0.0, -0.0, np.inf, -np.inf, 5e-324, 1e-310]))
    return np.concatenate(vals)


def main():
    exe = ROOT / "oracle" / "_ref" / "fp16_pin"
    exe.parent.mkdir(exist_ok=True)
    subprocess.run(["nvcc", "-std=c++17", "-O2", str(ROOT / "oracle" / "fp16_pin.cu"),
                    "-o", str(exe)], check=True)
    x = inputs()
    with tempfile.TemporaryDirectory() as d:
        d = Path(d)
        x.tofile(d / "in")
        subprocess.run([str(exe), str(d / "in"), str(x.size), str(d / "a"),
                        str(d / "b"), str(d / "c")], check=True)
        a = np.fromfile(d / "a", dtype=np.uint16)
        b = np.fromfile(d / "b", dtype=np.uint16)
        c = np.fromfile(d / "c", dtype=np.float32)
    np.savez_compressed(ROOT / "tests" / "golden" / "fp16_cuda_host.npz",
                        x=x, half_from_f64=a, half_from_f32=b, f32_from_half=c)
    print("wrote", x.size, "vectors")


if __name__ == "__main__":
    main()
