"""fp32 CPU restatement of diffusers' AutoencoderKL (SD2 VAE config).

TEST INFRASTRUCTURE: see oracle/__init__.py.  Reference call sites:
  app.ipynb:781-782, 793-794   vae.encode(x).latent_dist.sample() * scaling_factor
  app.ipynb:818-819            vae.decode(latents / scaling_factor).sample
  train_vae.py:721-722         vae(x)["sample"]
Math follows SURVEY.md Appendix A.2.  Pinned by the parameter identities
34,163,664 (encoder + quant_conv) and 49,490,199 (decoder + post_quant_conv).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import unet as _u
from .unet import ResnetBlock2D, Downsample2D, Upsample2D, conv2d, linear

SD2_VAE_CONFIG = dict(
    in_channels=3,
    out_channels=3,
    block_out_channels=(128, 256, 512, 512),
    layers_per_block=2,
    latent_channels=4,
    norm_num_groups=32,
    act_fn="silu",
    scaling_factor=0.18215,
    sample_size=512,
)


class VAEAttention(nn.Module):
    """Single-head (d=C) spatial self-attention with biased q/k/v/out (Appendix A.2).

    Key names follow diffusers >=0.17 (`to_q,to_k,to_v,to_out.0`); `load_state_dict_compat`
    also accepts the <=0.16 names (`query,key,value,proj_attn`)."""

    def __init__(self, ch: int, groups: int = 32):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6, affine=True)
        self.to_q = nn.Linear(ch, ch)
        self.to_k = nn.Linear(ch, ch)
        self.to_v = nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch), nn.Identity()])

    def forward(self, x):
        B, C, H, W = x.shape
        r = x
        h = F.group_norm(x, self.group_norm.num_groups, self.group_norm.weight, self.group_norm.bias,
                         self.group_norm.eps)
        h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
        q = linear(h, self.to_q.weight, self.to_q.bias)
        k = linear(h, self.to_k.weight, self.to_k.bias)
        v = linear(h, self.to_v.weight, self.to_v.bias)
        o = _u.attention_core(q, k, v, 1, C ** -0.5)
        o = linear(o, self.to_out[0].weight, self.to_out[0].bias)
        return o.reshape(B, H, W, C).permute(0, 3, 1, 2) + r


class _VAEMid(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, None, eps=1e-6), ResnetBlock2D(ch, ch, None, eps=1e-6)])
        self.attentions = nn.ModuleList([VAEAttention(ch)])

    def forward(self, h):
        h = self.resnets[0](h)
        h = self.attentions[0](h)
        return self.resnets[1](h)


class _EncDown(nn.Module):
    def __init__(self, cin, cout, has_down, layers):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, None, eps=1e-6)
                                      for j in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, 0)]) if has_down else None

    def forward(self, h):
        for r in self.resnets:
            h = r(h)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
        return h


class _DecUp(nn.Module):
    def __init__(self, cin, cout, has_up, layers):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, None, eps=1e-6)
                                      for j in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if has_up else None

    def forward(self, h):
        for r in self.resnets:
            h = r(h)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


class Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        boc = list(cfg["block_out_channels"])
        self.conv_in = nn.Conv2d(cfg["in_channels"], boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        prev = boc[0]
        for i, c in enumerate(boc):
            self.down_blocks.append(_EncDown(prev, c, i != len(boc) - 1, cfg["layers_per_block"]))
            prev = c
        self.mid_block = _VAEMid(boc[-1])
        self.conv_norm_out = nn.GroupNorm(cfg["norm_num_groups"], boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * cfg["latent_channels"], 3, padding=1)

    def forward(self, x):
        h = conv2d(x, self.conv_in.weight, self.conv_in.bias)
        for b in self.down_blocks:
            h = b(h)
        h = self.mid_block(h)
        h = F.silu(F.group_norm(h, self.conv_norm_out.num_groups, self.conv_norm_out.weight,
                                self.conv_norm_out.bias, self.conv_norm_out.eps))
        return conv2d(h, self.conv_out.weight, self.conv_out.bias)


class Decoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        boc = list(cfg["block_out_channels"])
        rboc = boc[::-1]
        self.conv_in = nn.Conv2d(cfg["latent_channels"], boc[-1], 3, padding=1)
        self.mid_block = _VAEMid(boc[-1])
        self.up_blocks = nn.ModuleList()
        prev = rboc[0]
        for i, c in enumerate(rboc):
            self.up_blocks.append(_DecUp(prev, c, i != len(boc) - 1, cfg["layers_per_block"] + 1))
            prev = c
        self.conv_norm_out = nn.GroupNorm(cfg["norm_num_groups"], boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], cfg["out_channels"], 3, padding=1)

    def forward(self, z):
        h = conv2d(z, self.conv_in.weight, self.conv_in.bias)
        h = self.mid_block(h)
        for b in self.up_blocks:
            h = b(h)
        h = F.silu(F.group_norm(h, self.conv_norm_out.num_groups, self.conv_norm_out.weight,
                                self.conv_norm_out.bias, self.conv_norm_out.eps))
        return conv2d(h, self.conv_out.weight, self.conv_out.bias)


class DiagonalGaussian:
    """diffusers DiagonalGaussianDistribution (Appendix A.2)."""

    def __init__(self, moments: torch.Tensor):
        self.parameters = moments
        self.mean, logvar = torch.chunk(moments, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None, noise: Optional[torch.Tensor] = None):
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


@dataclass
class EncoderOutput:
    latent_dist: DiagonalGaussian


@dataclass
class DecoderOutput:
    sample: torch.Tensor

    def __getitem__(self, k):
        if k in (0, "sample"):
            return self.sample
        raise KeyError(k)


class VAEOracle(nn.Module):
    def __init__(self, **overrides):
        super().__init__()
        cfg = dict(SD2_VAE_CONFIG)
        cfg.update(overrides)
        self.config = cfg
        self.encoder = Encoder(cfg)
        self.decoder = Decoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg["latent_channels"], 2 * cfg["latent_channels"], 1)
        self.post_quant_conv = nn.Conv2d(cfg["latent_channels"], cfg["latent_channels"], 1)

    @torch.no_grad()
    def encode(self, x):
        h = self.encoder(x)
        moments = conv2d(h, self.quant_conv.weight, self.quant_conv.bias, padding=0)
        return EncoderOutput(DiagonalGaussian(moments))

    @torch.no_grad()
    def decode(self, z):
        z = conv2d(z, self.post_quant_conv.weight, self.post_quant_conv.bias, padding=0)
        return DecoderOutput(self.decoder(z))

    @torch.no_grad()
    def forward(self, x, sample_posterior: bool = False, generator=None):
        post = self.encode(x).latent_dist
        z = post.sample(generator) if sample_posterior else post.mode()
        return self.decode(z)


_OLD2NEW = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}


def remap_legacy_vae_keys(sd: dict) -> dict:
    """Accept diffusers <=0.16 VAE attention key names."""
    out = {}
    for k, v in sd.items():
        parts = k.split(".")
        if "attentions" in parts:
            for old, new in _OLD2NEW.items():
                if parts[-2] == old:
                    parts[-2:-1] = new.split(".")
                    break
        out[".".join(parts)] = v
    return out
