"""fp32 CPU restatement of diffusers' UNet2DConditionModel (SD2-inpainting config).

TEST INFRASTRUCTURE: see oracle/__init__.py.  Reference call sites:
  app.ipynb:811-814        unet(cat([lat, mask, masked_lat], 1), t, ocr_embeddings).sample
  train_diffute_v1.py:912-913
Module math follows SURVEY.md Appendix A.1 (diffusers 0.15-0.2x semantics); the
state-dict key names are diffusers' so released DiffUTE checkpoints load.
Structure is pinned by the parameter identity 865,925,124 (tests/test_oracle_unet.py).

`emulate` (module-level, default None) optionally rounds the operands of every
contraction (conv / linear / attention matmuls) to a 16-bit format while
keeping fp32 accumulation: it is used only by the precision study in
tests/ and DESIGN.md, never by the oracle proper.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

SD2_INPAINT_UNET_CONFIG = dict(
    in_channels=9,
    out_channels=4,
    sample_size=64,
    block_out_channels=(320, 640, 1280, 1280),
    layers_per_block=2,
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
    attention_head_dim=(5, 10, 20, 20),  # diffusers misnomer: these are head COUNTS; head dim is 64
    cross_attention_dim=1024,
    use_linear_projection=True,
    norm_num_groups=32,
    norm_eps=1e-5,
    act_fn="silu",
    flip_sin_to_cos=True,
    freq_shift=0,
    downsample_padding=1,
    mid_block_scale_factor=1,
    upcast_attention=False,
)

# ---- precision emulation hook (study only) -------------------------------------------------
emulate: Optional[str] = None  # None | "fp16" | "bf16" | "fp16x2" | "bf16x2"


def _r(x: torch.Tensor) -> torch.Tensor:
    """Round a contraction operand the way the named 16-bit mode would (RN), keep fp32 container."""
    if emulate is None:
        return x
    if emulate == "fp16":
        return x.to(torch.float16).to(torch.float32)
    if emulate == "bf16":
        return x.to(torch.bfloat16).to(torch.float32)
    if emulate == "fp16x2":
        hi = x.to(torch.float16).to(torch.float32)
        lo = (x - hi).to(torch.float16).to(torch.float32)
        return hi + lo
    if emulate == "bf16x2":
        hi = x.to(torch.bfloat16).to(torch.float32)
        lo = (x - hi).to(torch.bfloat16).to(torch.float32)
        return hi + lo
    raise ValueError(emulate)


def conv2d(x, w, b, stride=1, padding=1):
    return F.conv2d(_r(x), _r(w), b, stride=stride, padding=padding)


def linear(x, w, b=None):
    return F.linear(_r(x), _r(w), b)


def attention_core(q, k, v, heads: int, scale: float):
    """softmax(q k^T * scale) v per head.  q:[B,Nq,C], k,v:[B,Nk,C] -> [B,Nq,C]."""
    B, Nq, C = q.shape
    Nk = k.shape[1]
    d = C // heads
    q = q.reshape(B, Nq, heads, d).permute(0, 2, 1, 3)
    k = k.reshape(B, Nk, heads, d).permute(0, 2, 1, 3)
    v = v.reshape(B, Nk, heads, d).permute(0, 2, 1, 3)
    s = torch.matmul(_r(q), _r(k).transpose(-1, -2)) * scale
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(_r(p), _r(v))
    return o.permute(0, 2, 1, 3).reshape(B, Nq, C)


# ---- building blocks ----------------------------------------------------------------------
class ResnetBlock2D(nn.Module):
    """GN->SiLU->conv3x3 -> +time_emb_proj(SiLU(temb)) -> GN->SiLU->conv3x3 -> + shortcut(x).

    Appendix A.1; shared by the VAE with temb_channels=None and eps=1e-6 (Appendix A.2)."""

    def __init__(self, cin: int, cout: int, temb_channels: Optional[int] = 1280, groups: int = 32, eps: float = 1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, cout) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = F.silu(F.group_norm(x, self.norm1.num_groups, self.norm1.weight, self.norm1.bias, self.norm1.eps))
        h = conv2d(h, self.conv1.weight, self.conv1.bias)
        if self.time_emb_proj is not None:
            h = h + linear(F.silu(temb), self.time_emb_proj.weight, self.time_emb_proj.bias)[:, :, None, None]
        h = F.silu(F.group_norm(h, self.norm2.num_groups, self.norm2.weight, self.norm2.bias, self.norm2.eps))
        h = conv2d(h, self.conv2.weight, self.conv2.bias)
        if self.conv_shortcut is not None:
            x = conv2d(x, self.conv_shortcut.weight, self.conv_shortcut.bias, padding=0)
        return x + h


class Attention(nn.Module):
    """q/k/v without bias, to_out.0 with bias; head dim 64; scale 64^-0.5 (Appendix A.1)."""

    def __init__(self, query_dim: int, heads: int, context_dim: Optional[int] = None):
        super().__init__()
        context_dim = context_dim or query_dim
        self.heads = heads
        self.scale = (query_dim // heads) ** -0.5
        self.to_q = nn.Linear(query_dim, query_dim, bias=False)
        self.to_k = nn.Linear(context_dim, query_dim, bias=False)
        self.to_v = nn.Linear(context_dim, query_dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(query_dim, query_dim), nn.Identity()])

    def forward(self, x, context=None):
        context = x if context is None else context
        q = linear(x, self.to_q.weight)
        k = linear(context, self.to_k.weight)
        v = linear(context, self.to_v.weight)
        o = attention_core(q, k, v, self.heads, self.scale)
        return linear(o, self.to_out[0].weight, self.to_out[0].bias)


class GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        a, g = linear(x, self.proj.weight, self.proj.bias).chunk(2, dim=-1)
        return a * F.gelu(g)  # exact erf GELU


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return linear(self.net[0](x), self.net[2].weight, self.net[2].bias)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, context_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, context_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, h, ctx):
        h = self.attn1(self.norm1(h)) + h
        h = self.attn2(self.norm2(h), ctx) + h
        h = self.ff(self.norm3(h)) + h
        return h


class Transformer2DModel(nn.Module):
    """GN(eps 1e-6) -> [B,HW,C] -> proj_in (Linear) -> block -> proj_out (Linear) -> + residual."""

    def __init__(self, dim: int, heads: int, context_dim: int, groups: int = 32):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, context_dim)])
        self.proj_out = nn.Linear(dim, dim)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        r = x
        h = F.group_norm(x, self.norm.num_groups, self.norm.weight, self.norm.bias, self.norm.eps)
        h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
        h = linear(h, self.proj_in.weight, self.proj_in.bias)
        for blk in self.transformer_blocks:
            h = blk(h, ctx)
        h = linear(h, self.proj_out.weight, self.proj_out.bias)
        return h.reshape(B, H, W, C).permute(0, 3, 1, 2) + r


class Downsample2D(nn.Module):
    def __init__(self, ch: int, padding: int = 1):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=padding)
        self.padding = padding

    def forward(self, x):
        if self.padding == 0:  # VAE encoder: asymmetric right/bottom zero pad (Appendix A.2)
            x = F.pad(x, (0, 1, 0, 1), value=0.0)
        return conv2d(x, self.conv.weight, self.conv.bias, stride=2, padding=self.padding)


class Upsample2D(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return conv2d(x, self.conv.weight, self.conv.bias)


class _DownBlock(nn.Module):
    def __init__(self, cin, cout, heads, ctx_dim, has_attn, has_down, layers=2):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout) for j in range(layers)])
        if has_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, ctx_dim) for _ in range(layers)])
        else:
            self.attentions = None
        self.downsamplers = nn.ModuleList([Downsample2D(cout, 1)]) if has_down else None

    def forward(self, h, temb, ctx, skips):
        for j, res in enumerate(self.resnets):
            h = res(h, temb)
            if self.attentions is not None:
                h = self.attentions[j](h, ctx)
            skips.append(h)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            skips.append(h)
        return h


class _MidBlock(nn.Module):
    def __init__(self, ch, heads, ctx_dim):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch), ResnetBlock2D(ch, ch)])
        self.attentions = nn.ModuleList([Transformer2DModel(ch, heads, ctx_dim)])

    def forward(self, h, temb, ctx):
        h = self.resnets[0](h, temb)
        h = self.attentions[0](h, ctx)
        return self.resnets[1](h, temb)


class _UpBlock(nn.Module):
    def __init__(self, prev_out, cout, skip_chs, heads, ctx_dim, has_attn, has_up):
        super().__init__()
        res = []
        for j, sc in enumerate(skip_chs):
            cin = (prev_out if j == 0 else cout) + sc
            res.append(ResnetBlock2D(cin, cout))
        self.resnets = nn.ModuleList(res)
        if has_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, ctx_dim) for _ in skip_chs])
        else:
            self.attentions = None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if has_up else None

    def forward(self, h, temb, ctx, skips):
        for j, res in enumerate(self.resnets):
            s = skips.pop()
            h = res(torch.cat([h, s], dim=1), temb)  # hidden first, skip second
            if self.attentions is not None:
                h = self.attentions[j](h, ctx)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


class _TimestepEmbedding(nn.Module):
    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        x = F.silu(linear(x, self.linear_1.weight, self.linear_1.bias))
        return linear(x, self.linear_2.weight, self.linear_2.bias)


def timestep_sincos(t: torch.Tensor, dim: int = 320, flip_sin_to_cos: bool = True, freq_shift: float = 0.0):
    """diffusers `Timesteps`: [B] -> [B, dim] fp32 (Appendix A.1)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32) / (half - freq_shift)
    arg = t[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


@dataclass
class UNetOutput:
    sample: torch.Tensor

    def __getitem__(self, k):
        return self.sample if k in (0, "sample") else (_ for _ in ()).throw(KeyError(k))


class UNetOracle(nn.Module):
    def __init__(self, **overrides):
        super().__init__()
        cfg = dict(SD2_INPAINT_UNET_CONFIG)
        cfg.update(overrides)
        self.config = cfg
        boc = list(cfg["block_out_channels"])
        heads = list(cfg["attention_head_dim"])
        ctx = cfg["cross_attention_dim"]
        tdim = boc[0] * 4
        self.conv_in = nn.Conv2d(cfg["in_channels"], boc[0], 3, padding=1)
        self.time_embedding = _TimestepEmbedding(boc[0], tdim)
        # down
        self.down_blocks = nn.ModuleList()
        skip_chs = [boc[0]]
        prev = boc[0]
        for i, (cout, typ) in enumerate(zip(boc, cfg["down_block_types"])):
            last = i == len(boc) - 1
            self.down_blocks.append(_DownBlock(prev, cout, heads[i], ctx, typ.startswith("CrossAttn"), not last,
                                               cfg["layers_per_block"]))
            skip_chs += [cout] * cfg["layers_per_block"] + ([] if last else [cout])
            prev = cout
        self.mid_block = _MidBlock(boc[-1], heads[-1], ctx)
        # up
        self.up_blocks = nn.ModuleList()
        rboc, rheads = boc[::-1], heads[::-1]
        prev = boc[-1]
        for i, (cout, typ) in enumerate(zip(rboc, cfg["up_block_types"])):
            n = cfg["layers_per_block"] + 1
            sk = [skip_chs.pop() for _ in range(n)]
            self.up_blocks.append(_UpBlock(prev, cout, sk, rheads[i], ctx, typ.startswith("CrossAttn"),
                                           i != len(boc) - 1))
            prev = cout
        self.conv_norm_out = nn.GroupNorm(cfg["norm_num_groups"], boc[0], eps=cfg["norm_eps"])
        self.conv_out = nn.Conv2d(boc[0], cfg["out_channels"], 3, padding=1)

    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, return_dict: bool = True):
        B = sample.shape[0]
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.int64 if isinstance(t, int) else torch.float32)
        elif t.dim() == 0:
            t = t[None]
        t = t.expand(B)
        temb = self.time_embedding(timestep_sincos(t, self.config["block_out_channels"][0],
                                                   self.config["flip_sin_to_cos"], self.config["freq_shift"]))
        h = conv2d(sample, self.conv_in.weight, self.conv_in.bias)
        skips = [h]
        for blk in self.down_blocks:
            h = blk(h, temb, encoder_hidden_states, skips)
        h = self.mid_block(h, temb, encoder_hidden_states)
        for blk in self.up_blocks:
            h = blk(h, temb, encoder_hidden_states, skips)
        h = F.silu(F.group_norm(h, self.conv_norm_out.num_groups, self.conv_norm_out.weight,
                                self.conv_norm_out.bias, self.conv_norm_out.eps))
        out = conv2d(h, self.conv_out.weight, self.conv_out.bias)
        return UNetOutput(out) if return_dict else (out,)
