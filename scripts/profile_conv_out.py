"""The UNet's conv_out (320 -> 4, 3x3, fused DDIM update) at batch 1 or 8, a few launches, for
  ncu --set full --import-source on -k regex:conv_out4 -s 2 -c 1 ...   Usage: python scripts/profile_conv_out.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w = torch.randn(4, 320, 3, 3, device="cuda") * 0.02
pk = ops.Packer("cuda"); wp = pk.small_out(w); pk.run()
b = torch.zeros(4, device="cuda")
x = torch.randn(B, 64, 64, 320, device="cuda")
lat = torch.randn(B, 4, 64, 64, device="cuda")
out, prev = torch.empty_like(lat), torch.empty_like(lat)
coef = torch.tensor([1.01, -0.03], device="cuda")
for _ in range(4):
    ops.conv_small_out(x, wp, b, out, sample=lat, prev=prev, coef=coef)
torch.cuda.synchronize()
print("done")
