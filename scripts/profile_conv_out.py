import sys, os
sys.path.insert(0, "/root/repo")
import torch
from diffute_b200 import ops
x = torch.randn(1, 512, 512, 128, device="cuda")
w = torch.randn(3, 128, 3, 3, device="cuda") * 0.05
wp = ops.pack_small_out_weight(w)
b = torch.zeros(3, device="cuda")
out = torch.empty(1, 3, 512, 512, device="cuda")
for _ in range(3):
    ops.conv_small_out(x, wp, b, out)
torch.cuda.synchronize()
print("done")
