"""VAE encode / decode timing at a given batch (CUDA events), per-kernel-class breakdown by the launch profiler."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import ops, synthetic, arch
from diffute_b200.vae import AutoencoderKL
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
px = int(sys.argv[2]) if len(sys.argv) > 2 else 512
vae = AutoencoderKL.from_synthetic(precision="fp16x2", encoder_precision="fp16")
x = (torch.rand((B, 3, px, px)) * 2 - 1).cuda()
def t_ms(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
z = vae.encode(x).latent_dist.mode()
print(f"B={B} {px}px encode {t_ms(lambda: vae.encode(x).latent_dist.mode()):.2f} ms  decode {t_ms(lambda: vae.decode(z).sample):.2f} ms", flush=True)
for name, fn in (("encode", lambda: vae.encode(x).latent_dist.mode()), ("decode", lambda: vae.decode(z).sample)):
    ops.PROFILE = {}
    torch.cuda._sleep(int(4e8))
    fn(); torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    print("  " + name + ": " + ", ".join(f"{k} {sum(a.elapsed_time(b) for a, b, _, _ in v):.2f}ms/{len(v)}" for k, v in sorted(prof.items())))
