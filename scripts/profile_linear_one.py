"""One UNet linear shape launched a few times for `ncu --set full --import-source on -k regex:gemm_tc`:
  python scripts/profile_linear_one.py --m 4096 --n 2560 --k 320 --geglu      (FeedForward GEGLU projection, level 0)
  python scripts/profile_linear_one.py --m 4096 --n 320 --k 320 --residual    (to_out / proj_out)"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=4096); ap.add_argument("--n", type=int, default=2560)
ap.add_argument("--k", type=int, default=320); ap.add_argument("--geglu", action="store_true")
ap.add_argument("--residual", action="store_true"); ap.add_argument("--f16", action="store_true")
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
dev = "cuda"
x16 = torch.randn(1, a.m, a.k, device=dev).half()
w = torch.randn(a.n, a.k, device=dev) * 0.05
w16 = ops.pack_linear_weight(w, 1, geglu=a.geglu)
bias = torch.randn(a.n, device=dev)
for _ in range(a.reps):
    if a.geglu:
        out = torch.empty(1, a.m, a.n // 2, dtype=torch.float16, device=dev)
        ops.linear(x16, w16, a.n, 1, out_f16=out, bias=ops.geglu_interleave(bias), geglu=True)
    elif a.f16:
        out = torch.empty(1, a.m, a.n, dtype=torch.float16, device=dev)
        ops.linear(x16, w16, a.n, 1, out_f16=out)
    else:
        out = torch.empty(a.m, a.n, device=dev)
        res = torch.randn(a.m, a.n, device=dev) if a.residual else None
        ops.linear(x16, w16, a.n, 1, out_f32=out, bias=bias, residual=res)
torch.cuda.synchronize()
print("done")
