import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import synthetic
from diffute_b200.pipeline import DiffUTEPipeline
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pipe = DiffUTEPipeline.from_synthetic("fp16", "fp16x2", vae_encoder_precision="fp16")
inp = synthetic.make_inputs(B, 512, 512)
d = {k: v.cuda() for k, v in inp.items()}
def call():
    return pipe(masked_image=d["masked_image"], mask_image=d["mask"], glyph_embeds=d["glyph_embeds"], latents=d["latents"],
                posterior_noise=d["posterior_noise"], num_inference_steps=50).images
for i in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); torch.cuda.synchronize()
    print(f"call {i}: events {e0.elapsed_time(e1):.1f} ms, wall {(time.perf_counter() - t0) * 1e3:.1f} ms, "
          f"gen {pipe.unet.buffer_generation()} graphs {len(pipe._graphs)} mem {torch.cuda.memory_allocated() / 1e9:.1f} GB", flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); call(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
