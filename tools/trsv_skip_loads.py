import sys, torch
sys.path.insert(0, "/root/repo")
import accessor_blas_b200 as ab
from bench import min_of_10
h = ab.Handle(0); dev = torch.device("cuda:0"); n = 16384
NAME = {torch.float64: "fp64", torch.float32: "fp32", torch.float16: "fp16"}
for st in (torch.float32, torch.float64, torch.float16):
    T = torch.empty(n * n, dtype=st, device=dev); b = torch.empty(n, dtype=st, device=dev)
    h.fill_uniform(n, n, T, n, 42, 0); h.fill_uniform(n, 1, b, 1, 42, n * n)
    T.mul_(0.01); T.view(n, n).diagonal().fill_(1.0)
    for ar in (torch.float64, torch.float32):
        out = []
        for ahead in (1024, -1, 1024, -1):
            ab.tune("trsv_l2_ahead", ahead)
            x = b.clone()
            out.append(min_of_10(lambda: h.trsv(ar, ab.LOWER, ab.UNIT, n, T, n, x, 1), torch) * 1e3)
        print(f"Acc<{NAME[ar]},{NAME[st]}>: normal {out[0]:.1f} / {out[2]:.1f} us, panel loads skipped {out[1]:.1f} / {out[3]:.1f} us", flush=True)
    del T
