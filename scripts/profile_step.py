"""One eager (un-graphed) UNet denoising step between cudaProfilerStart/Stop, for ncu:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python scripts/profile_step.py mixed 1
Usage: python scripts/profile_step.py [mixed|fp16x2|fp16] [batch] [vae]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import synthetic
from diffute_b200.pipeline import DiffUTEPipeline

mode = sys.argv[1] if len(sys.argv) > 1 else "mixed"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
do_vae = len(sys.argv) > 3 and sys.argv[3] == "vae"
up, vp = {"mixed": ("fp16", "fp16x2"), "fp16x2": ("fp16x2", "fp16x2"), "fp16": ("fp16", "fp16")}[mode]
pipe = DiffUTEPipeline.from_synthetic(up, vp, vae_encoder_precision="fp16" if mode == "mixed" else None)
inp = synthetic.make_inputs(B, 512, 512)
dev = pipe.device
h = w = 64
A = pipe.unet.arena
pipe.unet.prepare_context(inp["glyph_embeds"].to(dev))
lat = A.get("pipe.latents", (B, 4, h, w)); lat.copy_(inp["latents"])
mask = A.get("pipe.mask", (B, 1, h, w)); mask.copy_(inp["mask"][:, :, ::8, ::8])
ml = A.get("pipe.masked", (B, 4, h, w)); ml.copy_(inp["latents"] * 0.3)
state = A.get("pipe.state", (B + 2,)); state.fill_(981.0)
for _ in range(2):
    pipe.unet._forward_impl(B, h, w, srcs=[lat, mask, ml], t=state[:B])
if do_vae:
    pipe.vae.decode(lat, pre_scale=1 / 0.18215); pipe.vae.encode(inp["masked_image"].to(dev))
torch.cuda.synchronize()
torch.cuda.profiler.start()
pipe.unet._forward_impl(B, h, w, srcs=[lat, mask, ml], t=state[:B])
if do_vae:
    pipe.vae.encode(inp["masked_image"].to(dev)); pipe.vae.decode(lat, pre_scale=1 / 0.18215)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
from diffute_b200 import ops as _o
print("done", _o.gemm_stats())
