"""Which half of the VAE needs the 3-pass (hi/lo) contraction?  Full 512x512 / 50-step parity vs the CPU oracle with the
encoder and decoder precisions chosen independently (UNet fp16).  Writes gpurun_out/parity_vae_split.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import arch, synthetic
from diffute_b200.pipeline import DiffUTEPipeline
from diffute_b200.vae import AutoencoderKL
from oracle import DDIMOracle, UNetOracle, VAEOracle, sample_loop

torch.set_num_threads(min(os.cpu_count(), 32))
usd = synthetic.make_state_dict(arch.unet_param_shapes())
vsd = synthetic.make_state_dict(arch.vae_param_shapes())
inp = synthetic.make_inputs(1, 512, 512)
uo, vo = UNetOracle(), VAEOracle()
uo.load_state_dict(usd); vo.load_state_dict(vsd)
ref = sample_loop(uo, vo, DDIMOracle(), inp["masked_image"], inp["mask"], inp["glyph_embeds"], inp["latents"], 50,
                  posterior_noise=inp["posterior_noise"])
pipe = DiffUTEPipeline.from_synthetic("fp16", "fp16x2", state_dicts=(usd, vsd))
v_hi = pipe.vae
v_lo = AutoencoderKL(vsd, precision="fp16")


class Split:
    def __init__(self, enc, dec):
        self.enc, self.dec, self.config = enc, dec, enc.config

    def encode(self, *a, **k):
        return self.enc.encode(*a, **k)

    def decode(self, *a, **k):
        return self.dec.decode(*a, **k)


res = {}
for name, enc, dec in (("enc16x2_dec16x2", v_hi, v_hi), ("enc16_dec16x2", v_lo, v_hi), ("enc16x2_dec16", v_hi, v_lo),
                       ("enc16_dec16", v_lo, v_lo)):
    pipe.vae = Split(enc, dec)
    out = pipe(masked_image=inp["masked_image"], mask_image=inp["mask"], glyph_embeds=inp["glyph_embeds"],
               latents=inp["latents"], posterior_noise=inp["posterior_noise"], num_inference_steps=50).images.cpu()
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    res[name] = err
    print(name, err, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/parity_vae_split.json", "w"), indent=1)
