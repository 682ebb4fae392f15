"""Seeded synthetic weights and inputs for the DiffUTE sampling path.

No SD2 / DiffUTE checkpoint exists offline, so benchmarks and parity tests use
random-init weights of the exact SD2-inpainting / SD2-VAE architecture
(SURVEY.md 8d).  Everything is generated on the CPU with `torch.Generator`
(like the reference's own seeded noise, app.ipynb:796-801) and is a pure
function of (key, shape, seed): the CPU oracle and the GPU engine read
identical bytes without sharing code paths.

Init scheme ("variance preserving", stated here because the reference defines
none): conv / linear weights ~ U(-a, a) with a = sqrt(3 / fan_in), biases
~ U(-1, 1) / sqrt(fan_in); GroupNorm / LayerNorm weight = 1 + 0.1 N(0,1),
bias = 0.1 N(0,1).  Unit-variance signal propagation keeps the predicted noise
O(1), so the 50-step trajectory exercises the UNet rather than letting its
output vanish next to the latents.
"""
from __future__ import annotations

import zlib
from typing import Dict, Tuple

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def make_param(key: str, shape: Tuple[int, ...], seed: int = 1234) -> torch.Tensor:
    g = _gen(key, seed)
    is_norm = len(shape) == 1 and (".norm" in key or "norm_out" in key or "group_norm" in key or key.startswith("norm"))
    if is_norm:
        if key.endswith(".weight"):
            return 1.0 + 0.1 * torch.randn(shape, generator=g)
        return 0.1 * torch.randn(shape, generator=g)
    if key.endswith(".weight"):
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        a = (3.0 / fan_in) ** 0.5
        return (torch.rand(shape, generator=g) * 2 - 1) * a
    # bias of conv / linear: fan_in is not recoverable from the shape alone -> fixed small scale
    return (torch.rand(shape, generator=g) * 2 - 1) * 0.05


def make_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 1234) -> Dict[str, torch.Tensor]:
    return {k: make_param(k, s, seed) for k, s in shapes.items()}


def make_inputs(batch: int = 1, height: int = 512, width: int = 512, ctx_tokens: int = 577, ctx_dim: int = 1024,
                seed: int = 0):
    """Glyph-masked inpainting inputs shaped like app.ipynb:663-801 produces them.

    Returns dict(latents [B,4,h,w] (seed+0, like torch.manual_seed(0) app.ipynb:798), image [B,3,H,W] in [-1,1],
    mask [B,1,H,W] in {0,1} (a text-line box), masked_image (masked pixels = -1.0: the reference multiplies the
    uint8 image by (mask<0.5) *before* Normalize(0.5,0.5), app.ipynb:380-383, :332-336), glyph_embeds
    [B,577,1024] ~ N(0,1) (TrOCR's final LayerNorm output is ~unit variance), posterior_noise [B,4,h,w])."""
    def g(s):
        gen = torch.Generator(device="cpu")
        gen.manual_seed(seed + s)
        return gen

    h, w = height // 8, width // 8
    latents = torch.randn((batch, 4, h, w), generator=g(0))
    image = torch.rand((batch, 3, height, width), generator=g(1)) * 2 - 1
    mask = torch.zeros((batch, 1, height, width))
    mask[:, :, height // 4: height // 2, width // 8: 7 * width // 8] = 1.0
    masked_image = torch.where(mask > 0.5, torch.full_like(image, -1.0), image)
    posterior_noise = torch.randn((batch, 4, h, w), generator=g(2))
    glyph = torch.randn((batch, ctx_tokens, ctx_dim), generator=g(3))
    return dict(latents=latents, image=image, mask=mask, masked_image=masked_image,
                posterior_noise=posterior_noise, glyph_embeds=glyph)
