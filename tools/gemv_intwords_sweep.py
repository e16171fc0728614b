import sys, torch
sys.path.insert(0, "/root/repo")
import accessor_blas_b200 as ab
from bench import min_of_10, gemv_bytes
h = ab.Handle(0); dev = torch.device("cuda:0"); m = k = 16384
A = torch.empty(m * k, dtype=torch.float16, device=dev); x = torch.empty(k, dtype=torch.float16, device=dev); y = torch.zeros(m, dtype=torch.float16, device=dev)
h.fill_uniform(m, k, A, k, 42, 0); h.fill_uniform(k, 1, x, 1, 42, m * k)
for iw in (2, 1, 3, 2, 1, 3):
    ab.tune("gemv_intwords", iw)
    ms = min_of_10(lambda: h.gemv(torch.float64, m, k, 1.0, A, k, x, 1, 0.0, y, 1), torch)
    print(f"Acc<fp64,fp16> 64-bit-load pipeline, words on the integer pipes = {iw}: {gemv_bytes(m, k, 2) / ms / 1e6:.1f} GB/s", flush=True)
