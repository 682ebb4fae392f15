"""The memory-bound kernels at the shapes of UNet level 0 (and batch 8), a few launches each, for `ncu --set full`:
GroupNorm+SiLU 64x64x320 (gn_cluster_kernel), LayerNorm 4096x320, the nearest-2x / space-to-depth casts, conv_out+DDIM.
  ncu --set full --clock-control none -k regex:'gn_cluster|layernorm|cast_kernel|conv_small_out' -s 8 -c 8 -o gpurun_out/prof_norm \
      python scripts/profile_norm_one.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = "cuda"
x = torch.randn(B, 64, 64, 320, device=dev)
g, b = torch.ones(320, device=dev), torch.zeros(320, device=dev)
o16 = torch.empty(1, B, 64, 64, 320, dtype=torch.float16, device=dev)
t = torch.randn(B * 4096, 320, device=dev)
l16 = torch.empty(1, B * 4096, 320, dtype=torch.float16, device=dev)
up = torch.empty(1, B, 128, 128, 320, dtype=torch.float16, device=dev)
s2d = torch.empty(1, 4 * B, 32, 32, 320, dtype=torch.float16, device=dev)
wo = ops.pack_small_out_weight(torch.randn(4, 320, 3, 3, device=dev) * 0.02)
lat, prev = torch.randn(B, 4, 64, 64, device=dev), torch.empty(B, 4, 64, 64, device=dev)
eps_out = torch.empty(B, 4, 64, 64, device=dev)
coef = torch.tensor([0.99, -0.05], device=dev)
for _ in range(3):
    ops.groupnorm(x, g, b, 1e-5, True, 1, out16=o16)
    ops.layernorm(t, g, b, 1e-5, l16)
    ops.cast_f16(x, ops.CAST_UP2X, up)
    ops.cast_f16(x, ops.CAST_S2D, s2d)
    ops.conv_small_out(x, wo, None, eps_out, sample=lat, prev=prev, coef=coef)
torch.cuda.synchronize()
print("done")
