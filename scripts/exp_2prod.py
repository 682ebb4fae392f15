"""Experiment: one vs two TMA producer lanes in gemm_tc_kernel (DFU_GEMM_2PROD=0/1), single shapes, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import ops


def t_us(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def conv_case(B, H, W, Cin, Cout, cfgs):
    x16 = (torch.randn((B, H, W, Cin), device="cuda") * 0.5).half()
    w16 = ops.pack_conv_weight(torch.randn((Cout, Cin, 3, 3), device="cuda") * 0.02, 1)
    out = torch.empty((B * H * W, Cout), device="cuda")
    gf = 2.0 * B * H * W * Cout * Cin * 9 / 1e9
    for cfg in cfgs:
        us = t_us(lambda: ops.conv(x16, w16, Cout, 1, (B, H, W), ops.taps_3x3_s1(), tune=cfg, out_f32=out))
        print(f"conv B{B} {H}x{W} {Cin}->{Cout} {str(cfg):18s} {us:8.1f} us {gf / us * 1e3:7.1f} TF/s", flush=True)


def lin_case(M, N, K, cfgs):
    a16 = (torch.randn((1, M, K), device="cuda") * 0.5).half()
    w16 = ops.pack_linear_weight(torch.randn((N, K), device="cuda") * 0.02, 1)
    out = torch.empty((M, N), device="cuda")
    gf = 2.0 * M * N * K / 1e9
    for cfg in cfgs:
        us = t_us(lambda: ops.linear(a16, w16, N, 1, tune=cfg, out_f32=out))
        print(f"lin {M}x{N}x{K} {str(cfg):18s} {us:8.1f} us {gf / us * 1e3:7.1f} TF/s", flush=True)


print("DFU_GEMM_2PROD =", os.environ.get("DFU_GEMM_2PROD", "0"))
conv_case(1, 64, 64, 320, 320, [(80, 1, 4, 1), (80, 1, 8, 1), (160, 1, 6, 1), (160, 4, 3, 1), (160, 2, 6, 1)])
conv_case(1, 32, 32, 640, 640, [(80, 4, 4, 1), (160, 4, 6, 1), (80, 2, 8, 1)])
conv_case(1, 8, 8, 1280, 1280, [(128, 8, 6, 1), (128, 16, 3, 1)])
conv_case(8, 32, 32, 640, 640, [(160, 1, 3, 1), (160, 1, 6, 1)])
conv_case(8, 64, 64, 320, 320, [(160, 1, 3, 1)])
lin_case(4096, 320, 320, [(64, 1, 4, 1), (160, 1, 3, 1)])
lin_case(32768, 2560, 320, [(256, 1, 2, 1)])
