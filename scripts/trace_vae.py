"""Timeline of one VAE encode + decode at 512x512 from in-kernel records (diagnostic library):
  DFU_TRACE=1 python scripts/trace_vae.py [batch] > gpurun_out/trace_vae.txt"""
import os, sys
os.environ["DFU_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import synthetic, trace
from diffute_b200.pipeline import DiffUTEPipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
pipe = DiffUTEPipeline.from_synthetic("fp16", os.environ.get("VAE_PREC", "fp16x2"))
inp = synthetic.make_inputs(B, 512, 512)
x = inp["masked_image"].to(pipe.device)
lat = inp["latents"].to(pipe.device)
for _ in range(2):
    pipe.vae.encode(x); pipe.vae.decode(lat, pre_scale=1 / 0.18215)
torch.cuda.synchronize()
trace.enable(1 << 20)
for name, fn in (("encode", lambda: pipe.vae.encode(x)), ("decode", lambda: pipe.vae.decode(lat, pre_scale=1 / 0.18215))):
    trace.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(4e7))  # let the host queue the launches first
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    L = trace.collect()
    print(f"== {name}: {e0.elapsed_time(e1) * 1e3:.0f} us by events, {len(L)} launches")
    prev = None
    agg = {}
    for i, d in enumerate(L):
        end = d.get("end_last", d["start_last"])
        wait = d.get("wait_first", d["start_first"])
        dt = 0 if prev is None else max(0.0, end - prev)
        prev = end if prev is None else max(prev, end)
        a = agg.setdefault(d["kernel"], [0, 0.0]); a[0] += 1; a[1] += dt
        if dt > 150:
            print(f"  {i:4d} {d['kernel']:14s} ctas {d['nctas']:6d} extra {d['extra']:#x} cost {dt:8.1f} us body {end - wait:8.1f}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:14s} {a[0]:4d} launches {a[1]:9.1f} us")
