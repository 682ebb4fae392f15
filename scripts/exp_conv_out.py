"""conv_out (320 -> 4, 3x3) of the UNet at batch 1..8: the one-warp-per-pixel variant against the variant that stages the
46 KB of weights in shared memory once per CTA (DFU_CONV_OUT_SMEM=0/1).  Usage: DFU_CONV_OUT_SMEM=1 python scripts/exp_conv_out.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import ops
w = torch.randn(4, 320, 3, 3, device="cuda") * 0.02
pk = ops.Packer("cuda"); wp = pk.small_out(w); pk.run()
b = torch.zeros(4, device="cuda")
for B in (1, 2, 4, 8, 16):
    x = torch.randn(B, 64, 64, 320, device="cuda")
    out = torch.empty(B, 4, 64, 64, device="cuda")
    for _ in range(5):
        ops.conv_small_out(x, wp, b, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        ops.conv_small_out(x, wp, b, out)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us  (DFU_CONV_OUT_SMEM={os.environ.get('DFU_CONV_OUT_SMEM', 'auto')})")

# conv_in (9 -> 320, 3x3, three gathered NCHW sources), 32 or 16 pixels per CTA (DFU_CONV_IN_PIX)
wi = torch.randn(320, 9, 3, 3, device="cuda") * 0.1
pk = ops.Packer("cuda"); wt = pk.small_in(wi); pk.run()
bi = torch.zeros(320, device="cuda")
for B in (1, 2, 8):
    lat, msk, ml = (torch.randn(B, c, 64, 64, device="cuda") for c in (4, 1, 4))
    o = torch.empty(B, 64, 64, 320, device="cuda")
    for _ in range(5):
        ops.conv_small_in([lat, msk, ml], wt, bi, o, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        ops.conv_small_in([lat, msk, ml], wt, bi, o, B)
    e1.record(); torch.cuda.synchronize()
    print(f"conv_in B={B}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us  (DFU_CONV_IN_PIX={os.environ.get('DFU_CONV_IN_PIX', 'auto')})")
