"""Experiment: the persistent CTA-pair kernel against the tile-per-CTA kernel on single shapes, stage-count scaling.
  python scripts/exp_pair.py            (DFU_G2_DEBUG=1: every CTA signals its own full barrier — timing only)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import ops


def t_ms(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def conv_case(B, H, W, Cin, Cout, cfgs, prec=1):
    planes = ops.planes_of(prec)
    x16 = (torch.randn((planes * B, H, W, Cin), device="cuda") * 0.5).half()
    w16 = ops.pack_conv_weight(torch.randn((Cout, Cin, 3, 3), device="cuda") * 0.02, planes)
    out = torch.empty((B * H * W, Cout), device="cuda")
    gf = 2.0 * B * H * W * Cout * Cin * 9 / 1e9
    print(f"conv B{B} {H}x{W} {Cin}->{Cout} prec{prec}: {gf:.1f} GF")
    for cfg in cfgs:
        try:
            us = t_ms(lambda: ops.conv(x16, w16, Cout, prec, (B, H, W), ops.taps_3x3_s1(), tune=cfg, out_f32=out))
            print(f"   {str(cfg):22s} {us:8.1f} us  {gf / us * 1e3:7.1f} TF/s", flush=True)
        except Exception as e:
            print(f"   {cfg} failed: {str(e)[:100]}")


def lin_case(M, N, K, cfgs, prec=1):
    planes = ops.planes_of(prec)
    a16 = (torch.randn((planes, M, K), device="cuda") * 0.5).half()
    w16 = ops.pack_linear_weight(torch.randn((N, K), device="cuda") * 0.02, planes)
    out = torch.empty((M, N), device="cuda")
    gf = 2.0 * M * N * K / 1e9
    print(f"linear {M}x{N}x{K} prec{prec}: {gf:.1f} GF")
    for cfg in cfgs:
        try:
            us = t_ms(lambda: ops.linear(a16, w16, N, prec, tune=cfg, out_f32=out))
            print(f"   {str(cfg):22s} {us:8.1f} us  {gf / us * 1e3:7.1f} TF/s", flush=True)
        except Exception as e:
            print(f"   {cfg} failed: {str(e)[:100]}")


pair = lambda bn: [(bn, 1, s, 2) for s in (2, 3, 4, 5, 7, 0)]
conv_case(8, 32, 32, 640, 640, [(160, 1, 3, 1)] + pair(160))
lin_case(8192, 640, 5760, [(160, 1, 3, 1)] + pair(160))
lin_case(8192, 512, 4096, [(256, 1, 2, 1)] + pair(256))
conv_case(8, 64, 64, 320, 320, [(160, 1, 3, 1)] + pair(160))
conv_case(1, 512, 512, 128, 128, [(128, 1, 3, 1)] + pair(128))
conv_case(1, 256, 256, 256, 256, [(256, 1, 2, 1)] + pair(256), prec=2)
