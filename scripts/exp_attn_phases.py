"""Where a softmax warpgroup of attn_fwd_kernel spends its cycles (diagnostic -DDFU_TRACE build): per launch the median
over CTAs of thread 0's cycles waiting for S (Q K^T round trip), for the previous P V, and computing, per own block.
Usage: DFU_TRACE=1 python scripts/exp_attn_phases.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from diffute_b200 import ops, trace

assert os.environ.get("DFU_TRACE") == "1"
buf = trace.enable(1 << 16)
dev = "cuda"
for (B, heads, Nq, Nk) in [(1, 5, 4096, 4096), (1, 5, 4096, 577), (1, 10, 1024, 1024), (1, 10, 1024, 577),
                           (1, 20, 256, 577), (1, 20, 256, 256), (8, 5, 4096, 4096), (8, 10, 1024, 577)]:
    C = heads * 64
    q = torch.randn(1, B * Nq, C, device=dev).half()
    kv = torch.randn(1, B * Nk, 2 * C, device=dev).half()
    out = torch.empty(1, B * Nq, C, dtype=torch.float16, device=dev)
    for _ in range(3):
        ops.attention(q, 0, kv, 0, kv, C, B, heads, Nq, Nk, 0.125, out)
    torch.cuda.synchronize()
    trace.reset()
    ops.attention(q, 0, kv, 0, kv, C, B, heads, Nq, Nk, 0.125, out)
    torch.cuda.synchronize()
    n = int(buf[0].item())
    r = buf[8:8 + n * 16].view(n, 16).cpu().numpy()
    r = r[(r[:, 1] & 0xFF) == 3]
    nb = np.maximum(r[:, 15], 1)
    tot = (r[:, 10] - r[:, 4])
    print(f"B{B} h{heads} Nq{Nq} Nk{Nk}: ctas {len(r)}  own blocks/CTA med {np.median(r[:,15]):.0f}  per own block (cycles, median CTA): "
          f"wait S {np.median(r[:,12]/nb):7.0f}  wait PV {np.median(r[:,13]/nb):6.0f}  compute {np.median(r[:,14]/nb):7.0f}   "
          f"CTA lifetime {np.median(tot):8.0f} cyc = {np.median(tot)/1965:.1f} us")
