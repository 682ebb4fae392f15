"""In-graph cost of dependent kernel chains (is PDL engaging? what is the fixed cost per kernel?).
Usage: DFU_PDL=0|1 python scripts/microbench_launch.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import ops

dev = "cuda"
def bench(name, fn, n=200, reps=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"PDL={os.environ.get('DFU_PDL','1')} {name:34s} {e0.elapsed_time(e1)*1e3/(reps*n):7.2f} us per kernel in-graph", flush=True)

x = torch.randn(4096, 320, device=dev); g_ = torch.ones(320, device=dev); b_ = torch.zeros(320, device=dev)
o16 = torch.empty(1, 4096, 320, dtype=torch.float16, device=dev)
bench("layernorm 4096x320", lambda: ops.layernorm(x, g_, b_, 1e-5, o16))
xs = torch.randn(64, 1280, device=dev); gs = torch.ones(1280, device=dev); bs = torch.zeros(1280, device=dev)
o16s = torch.empty(1, 64, 1280, dtype=torch.float16, device=dev)
bench("layernorm 64x1280", lambda: ops.layernorm(xs, gs, bs, 1e-5, o16s))
for (M, N, K, tune) in [(4096, 320, 320, (0, 0, 0)), (1024, 640, 640, (0, 0, 0)), (256, 1280, 1280, (0, 0, 0)),
                        (256, 1280, 1280, (160, 1, 3)), (4096, 320, 2880, (0, 0, 0)), (64, 1280, 11520, (0, 0, 0))]:
    a16 = torch.randn(1, M, K, device=dev).half(); w16 = torch.randn(N, K, device=dev).half()
    out = torch.empty(M, N, device=dev); res = torch.randn(M, N, device=dev); bias = torch.zeros(N, device=dev)
    bench(f"linear {M}x{N}x{K} tune={tune}", lambda: ops.linear(a16, w16, N, 1, tune=tune, out_f32=out, bias=bias, residual=res), n=100)
xg = torch.randn(1, 64, 64, 320, device=dev); gg = torch.ones(320, device=dev); bg = torch.zeros(320, device=dev)
og = torch.empty(1, 1, 64, 64, 320, dtype=torch.float16, device=dev)
bench("groupnorm 64x64x320", lambda: ops.groupnorm(xg, gg, bg, 1e-5, True, 1, out16=og), n=100)
xg2 = torch.randn(1, 16, 16, 1280, device=dev); gg2 = torch.ones(1280, device=dev); bg2 = torch.zeros(1280, device=dev)
og2 = torch.empty(1, 1, 16, 16, 1280, dtype=torch.float16, device=dev)
bench("groupnorm 16x16x1280", lambda: ops.groupnorm(xg2, gg2, bg2, 1e-5, True, 1, out16=og2), n=100)
print(ops.gemm_stats())
