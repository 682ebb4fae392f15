"""CPU restatement of the reference sampling loop (TEST INFRASTRUCTURE).

Follows app.ipynb:772-819 line by line, with the two sources of run-to-run
non-determinism in the reference neutralised exactly as SURVEY.md 8c prescribes:
`latent_dist.sample()` takes an explicit eps (or uses mode()), and the initial
latents are passed in (the reference draws them with a CPU generator, seed 0,
app.ipynb:798).  The reference's first, dead `vae.encode(image)` (app.ipynb:781,
result overwritten at :798) is skipped.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F


@torch.no_grad()
def sample_loop(unet, vae, scheduler, masked_image: torch.Tensor, mask: torch.Tensor, glyph_embeds: torch.Tensor,
                latents: torch.Tensor, num_inference_steps: int = 50, posterior_noise: Optional[torch.Tensor] = None,
                guidance_scale: float = 1.0, negative_glyph_embeds: Optional[torch.Tensor] = None,
                return_latents: bool = False):
    """masked_image [B,3,H,W] in [-1,1]; mask [B,1,H,W] in {0,1}; glyph_embeds [B,577,1024]; latents [B,4,H/8,W/8]."""
    sf = vae.config["scaling_factor"]
    vsf = 2 ** (len(vae.config["block_out_channels"]) - 1)
    H, W = mask.shape[-2:]
    mask_l = F.interpolate(mask, size=(H // vsf, W // vsf))  # nearest (app.ipynb:787-790)
    post = vae.encode(masked_image).latent_dist               # app.ipynb:793
    ml = (post.mode() if posterior_noise is None else post.sample(noise=posterior_noise)) * sf
    latents = latents * scheduler.init_noise_sigma            # app.ipynb:800
    scheduler.set_timesteps(num_inference_steps)              # app.ipynb:803
    do_cfg = guidance_scale != 1.0 and negative_glyph_embeds is not None
    for t in scheduler.timesteps:                             # app.ipynb:806
        x = scheduler.scale_model_input(latents, t)           # app.ipynb:810
        x = torch.cat([x, mask_l, ml], dim=1)                 # app.ipynb:811
        if do_cfg:  # the commented-out intent at train_diffute_v1.py:915-917
            eps_c = unet(x, t, glyph_embeds).sample
            eps_u = unet(x, t, negative_glyph_embeds).sample
            eps = eps_u + guidance_scale * (eps_c - eps_u)
        else:
            eps = unet(x, t, glyph_embeds).sample             # app.ipynb:814
        latents = scheduler.step(eps, t, latents).prev_sample  # app.ipynb:816
    if return_latents:
        return latents
    return vae.decode(latents / sf).sample                    # app.ipynb:818-819
