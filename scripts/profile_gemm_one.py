"""A single representative contraction (UNet level-0 3x3 conv, M=4096 N=320 K=2880, or any --m/--n/--cin) launched a few
times, for `ncu --set full -k regex:gemm_tc -s 2 -c 1 ...`.  Also usable with attention: --attn."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--hw", type=int, default=64); ap.add_argument("--cin", type=int, default=320)
ap.add_argument("--cout", type=int, default=320); ap.add_argument("--prec", type=int, default=1)
ap.add_argument("--attn", action="store_true"); ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--tune", default="0,0,0", help="block_n,splits,stages[,kernel] (kernel 2 = persistent CTA pairs)")
a = ap.parse_args()
tune = tuple(int(v) for v in a.tune.split(","))
dev = "cuda"
if a.attn:
    B, heads, N = 1, 5, 4096
    qkv = torch.randn(a.prec, B * N, 3 * heads * 64, device=dev).half()
    out = torch.empty(a.prec, B * N, heads * 64, dtype=torch.float16, device=dev)
    for _ in range(a.reps):
        ops.attention(qkv, 0, qkv, heads * 64, qkv, 2 * heads * 64, B, heads, N, N, 0.125, out)
else:
    P = ops.planes_of(a.prec)
    H = W = a.hw
    x16 = torch.randn(P * a.batch, H, W, a.cin, device=dev).half()
    w16 = ops.pack_conv_weight(torch.randn(a.cout, a.cin, 3, 3, device=dev) * 0.02, P)
    out = torch.empty(a.batch * H * W, a.cout, device=dev)
    bias = torch.zeros(a.cout, device=dev)
    for _ in range(a.reps):
        ops.conv(x16, w16, a.cout, a.prec, (a.batch, H, W), ops.taps_3x3_s1(), tune=tune, out_f32=out, bias=bias)
torch.cuda.synchronize()
print("done")
