"""GroupNorm variants timed with the in-kernel tracer (diagnostic library): cluster shapes (DFU_GN_FORCE=S,T) against
the two-launch path (DFU_GN_CLUSTER=0 needs a fresh process, so it is selected with argv).
  DFU_TRACE=1 python scripts/bench_gn.py [H W C [B]]"""
import os, sys
os.environ["DFU_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import ops, trace

H, W, C = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 64, 320)
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1
x = torch.randn(B, H, W, C, device="cuda")
g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
out = torch.empty(1, B, H, W, C, dtype=torch.float16, device="cuda")
print(f"GroupNorm B={B} {H}x{W}x{C}: {B * H * W * C * 6 / 1e6:.1f} MB algorithmic (4 B read + 2 B write per element)")
w = torch.randn(1 << 20, device="cuda")
trace.enable(1 << 16)


def run(label):
    for _ in range(3):
        ops.groupnorm(x, g, b, 1e-5, True, 1, out16=out)
    torch.cuda.synchronize()
    res = []
    for _ in range(5):
        w.mul_(1.0001)  # a preceding kernel, like in the step
        trace.reset()
        ops.groupnorm(x, g, b, 1e-5, True, 1, out16=out)
        recs = trace.collect()
        t0 = min(r.get("wait_first", r["start_first"]) for r in recs)
        t1 = max(r.get("end_last", 0) for r in recs)
        res.append((t1 - t0, [(r["kernel"], r["nctas"], round(r.get("p7_med", 0) - r.get("wait_med", 0), 2),
                               round(r.get("end_clk_med", 0) - r.get("wait_med", 0), 2)) for r in recs]))
    res.sort(key=lambda t: t[0])
    print(f"{label:24s} median {res[2][0]:6.2f} us  {res[2][1]}", flush=True)


run("auto")
for S in (8, 4, 2, 1):
    for T in (128, 256, 512):
        os.environ["DFU_GN_FORCE"] = f"{S},{T}"
        try:
            run(f"S={S} T={T}")
        except Exception as e:
            print(S, T, "failed", repr(e)[:100])
