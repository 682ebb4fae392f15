"""CPU study: which contraction-operand precision meets the 1e-3 decoded-RGB bar?

Runs the fp32 oracle and operand-rounded variants (fp32 accumulate, RN operands) on identical
seeded synthetic weights/inputs and prints max|y-y_ref|/max|y_ref| on latents and decoded RGB.
Usage: python scripts/precision_study.py [px] [steps] [modes...]
"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle.unet as ou
from oracle import UNetOracle, VAEOracle, DDIMOracle, sample_loop
from diffute_b200 import arch, synthetic

px = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
modes = sys.argv[3:] or ["fp16", "bf16", "fp16x2"]
torch.set_num_threads(os.cpu_count())
u = UNetOracle(); u.load_state_dict(synthetic.make_state_dict(arch.unet_param_shapes()))
v = VAEOracle(); v.load_state_dict(synthetic.make_state_dict(arch.vae_param_shapes()))
inp = synthetic.make_inputs(1, px, px)

class _Emu:
    """run a module's forward under a given emulation mode"""
    def __init__(self, mod, mode, names):
        self.mod, self.mode, self.config = mod, mode, mod.config
        for n in names:
            setattr(self, n, self._wrap(getattr(mod, n)))
    def _wrap(self, fn):
        def f(*a, **k):
            old = ou.emulate; ou.emulate = self.mode
            try: return fn(*a, **k)
            finally: ou.emulate = old
        return f
    def __call__(self, *a, **k):
        return self._wrap(self.mod.__call__)(*a, **k)

def run(mode):
    # mode "A+B": UNet operands in A, VAE operands in B ("fp32" = no rounding)
    mu, mv = (mode.split("+") + [None])[:2] if mode and "+" in mode else (mode, mode)
    mu = None if mu == "fp32" else mu; mv = None if mv == "fp32" else mv
    uu = _Emu(u, mu, []); vv = _Emu(v, mv, ["encode", "decode"])
    t0 = time.time()
    lat = sample_loop(uu, vv, DDIMOracle(), inp["masked_image"], inp["mask"], inp["glyph_embeds"], inp["latents"],
                      steps, posterior_noise=inp["posterior_noise"], return_latents=True)
    rgb = vv.decode(lat / 0.18215).sample
    return lat, rgb, time.time() - t0

lat0, rgb0, dt = run(None)
print(f"fp32 ref: {dt:.1f}s  lat absmax {lat0.abs().max():.3f} std {lat0.std():.3f}  rgb absmax {rgb0.abs().max():.3f}", flush=True)
for m in modes:
    lat, rgb, dt = run(m)
    el = ((lat - lat0).abs().max() / lat0.abs().max()).item()
    er = ((rgb - rgb0).abs().max() / rgb0.abs().max()).item()
    print(f"{m:8s}: {dt:.1f}s  latents maxrel {el:.3e}   rgb maxrel {er:.3e}", flush=True)
