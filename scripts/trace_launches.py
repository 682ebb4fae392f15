"""CPU dry run of one UNet step: records every C-ABI call the orchestration makes (no GPU, no kernels), with the
tiling dfu_gemm_plan would choose.  Optionally joins an ncu launch list by launch order.
Usage: python scripts/trace_launches.py [batch] [latent] [launches.csv]"""
import csv, ctypes as C, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import _lib, arch, ops, unet as U

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L = int(sys.argv[2]) if len(sys.argv) > 2 else 64
real = _lib.lib()
calls = []

class Fake:
    def __getattr__(self, name):
        if name == "dfu_gemm_workspace":
            return lambda d: 0
        if name == "dfu_groupnorm_workspace":
            return real.dfu_groupnorm_workspace
        if name == "dfu_last_error":
            return real.dfu_last_error
        def f(*a):
            if name == "dfu_gemm":
                d = a[0]._obj
                out = (C.c_int32 * 8)()
                real.dfu_gemm_plan(C.byref(d), out)
                k = sum(d.g[i].ntaps * d.g[i].k_per_tap for i in range(d.ngroups))
                calls.append(("gemm_tc_kernel", dict(m=d.m, n=d.n, k=k * d.npass, conv=d.conv, epi=d.epi, block_n=out[0],
                                                    splits=out[1], stages=out[2], tiles=out[3] * out[4], kb=out[5])))
                if out[1] > 1:
                    calls.append(("splitk_reduce_kernel", dict(m=d.m, n=d.n, splits=out[1])))
            else:
                calls.append((name, {}))
            return 0
        return f

ops.lib = lambda: Fake()
ops._stream = lambda: 0
torch.set_num_threads(8)
sd = {k: torch.zeros(s) for k, s in arch.unet_param_shapes().items()}
net = U.UNet2DConditionModel(sd, device="cpu", precision="fp16", use_cuda_graph=False)
net.prepare_context(torch.zeros(B, 577, 1024))
calls.clear()
net._forward_impl(B, L, L)
g = [c for c in calls if c[0] in ("gemm_tc_kernel", "splitk_reduce_kernel")]
print(f"{len(calls)} C-ABI calls per step, {sum(1 for c in calls if c[0]=='gemm_tc_kernel')} gemm launches, "
      f"{sum(1 for c in calls if c[0]=='splitk_reduce_kernel')} split-K reduces")
times = None
if len(sys.argv) > 3:
    lines = [l for l in open(sys.argv[3]) if l.startswith('"')]
    times = {}
    for r in csv.DictReader(lines):
        if r["Metric Name"] == "gpu__time_duration.sum":
            times.setdefault(re.sub(r"\(.*", "", r["Kernel Name"]), []).append(float(r["Metric Value"].replace(",", "")) / 1000)
idx = {}
tot = {}
for name, info in g:
    us = None
    if times and name in times:
        i = idx.get(name, 0); idx[name] = i + 1
        us = times[name][i] if i < len(times[name]) else None
    if name == "gemm_tc_kernel":
        fl = 2.0 * info["m"] * info["n"] * info["k"]
        key = (info["m"], info["n"], info["k"], info["block_n"], info["splits"], info["tiles"])
        t = tot.setdefault(key, [0, 0.0, fl])
        t[0] += 1
        t[1] += us or 0.0
print(f"{'M':>6} {'N':>6} {'K':>6} {'bn':>4} {'spl':>4} {'tiles':>5} {'cnt':>4} {'avg_us':>8} {'tot_ms':>7} {'TFLOP/s':>8}")
for key, (cnt, us, fl) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    tf = (fl * cnt / (us * 1e-6) / 1e12) if us else 0.0
    print(f"{key[0]:6d} {key[1]:6d} {key[2]:6d} {key[3]:4d} {key[4]:4d} {key[5]:5d} {cnt:4d} {us/cnt if cnt else 0:8.1f} {us/1000:7.3f} {tf:8.1f}")
