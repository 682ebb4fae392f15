"""Generate the small golden vectors under tests/golden/ from the CPU oracle (run in the build container).

The reference (chenhaoxing/DiffUTE) ships no tests or golden vectors and its arithmetic lives in an uninstallable
third-party package, so these fixtures pin the ORACLE (KAT-pinned schedulers + param-count-pinned architecture) on
seeded synthetic weights; GPU tests compare the CUDA path to them without needing the oracle to run at full size.
Usage: python tests/golden/make_golden.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from diffute_b200 import arch, synthetic  # noqa: E402
from oracle import DDIMOracle, UNetOracle, VAEOracle, sample_loop  # noqa: E402

torch.set_num_threads(os.cpu_count())
u, v = UNetOracle(), VAEOracle()
u.load_state_dict(synthetic.make_state_dict(arch.unet_param_shapes()))
v.load_state_dict(synthetic.make_state_dict(arch.vae_param_shapes()))
out = {}
inp = synthetic.make_inputs(1, 64, 64)
x = torch.cat([inp["latents"], inp["mask"][:, :, ::8, ::8], inp["latents"] * 0.5], 1)
out["unet_L8_t981"] = u(x, 981, inp["glyph_embeds"]).sample.flatten().tolist()
post = v.encode(inp["masked_image"]).latent_dist
out["vae_moments_64px"] = post.parameters.flatten().tolist()
z = post.sample(noise=inp["posterior_noise"])
out["vae_decode_64px"] = v.decode(z / 0.18215).sample.flatten()[::7].tolist()
rgb = sample_loop(u, v, DDIMOracle(), inp["masked_image"], inp["mask"], inp["glyph_embeds"], inp["latents"], 4,
                  posterior_noise=inp["posterior_noise"])
out["loop_64px_4steps_rgb"] = rgb.flatten()[::7].tolist()
s = DDIMOracle()
s.set_timesteps(50)
out["ddim_sd2_coeffs"] = [list(s.collapsed_coeffs(int(t))) for t in s.timesteps]
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json"), "w") as f:
    json.dump(out, f)
print({k: len(val) for k, val in out.items()})
