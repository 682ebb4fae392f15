"""Second, structurally independent restatement of the DiffUTE sampling math (TEST INFRASTRUCTURE ONLY).

Purpose: `oracle/` (torch fp32, nn.Module tree, F.conv2d / F.group_norm / F.layer_norm / torch.softmax) is the checker
of the CUDA path, but the reference ships no golden vectors for the UNet / VAE (diffusers is neither
vendored nor installable here).  This file restates the same published architecture a second time in a deliberately
different form, so that a mistake in one restatement cannot hide in the other:

  * numpy float64 instead of torch float32; no autograd modules, a flat functional walk over the state-dict keys;
  * NHWC tensors; convolutions as an explicit im2col gather followed by ONE matrix product;
  * GroupNorm / LayerNorm / softmax / SiLU / erf-GELU / nearest upsample written out from their definitions;
  * the UNet skip bookkeeping derived from the block list, not copied from the oracle's module constructors.

It follows the same reference call sites (app.ipynb:772-819; train_diffute_v1.py:875-913) and SURVEY.md Appendix A.
tests/test_oracle_crosscheck.py compares it with `oracle/` primitive by primitive, block by block and end to end.
"""
from __future__ import annotations

import math

import numpy as np


def f64(t):
    return t.detach().cpu().double().numpy() if hasattr(t, "detach") else np.asarray(t, dtype=np.float64)


# ---------------------------------------------------------------------------------------------
# primitives (NHWC, float64)
# ---------------------------------------------------------------------------------------------
_ERF = np.vectorize(math.erf, otypes=[np.float64])


def silu(x):
    return x / (1.0 + np.exp(-x))


def gelu_erf(x):
    return 0.5 * x * (1.0 + _ERF(x / math.sqrt(2.0)))


def group_norm(x, gamma, beta, groups, eps):
    """x [B,H,W,C]: statistics over (H, W, C/groups) per sample and group, biased variance."""
    B, H, W, C = x.shape
    cpg = C // groups
    y = np.empty_like(x)
    for b in range(B):
        for g in range(groups):
            blk = x[b, :, :, g * cpg:(g + 1) * cpg]
            mu = blk.sum() / blk.size
            var = ((blk - mu) ** 2).sum() / blk.size
            y[b, :, :, g * cpg:(g + 1) * cpg] = (blk - mu) / math.sqrt(var + eps)
    return y * gamma + beta


def layer_norm(x, gamma, beta, eps=1e-5):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * gamma + beta


def conv2d(x, w, b, stride=1, pad=(1, 1, 1, 1)):
    """x [B,H,W,Cin]; w torch layout [Cout,Cin,kh,kw]; pad = (top, bottom, left, right).  im2col + one matmul."""
    B, H, W, Cin = x.shape
    Cout, _, kh, kw = w.shape
    xp = np.zeros((B, H + pad[0] + pad[1], W + pad[2] + pad[3], Cin))
    xp[:, pad[0]:pad[0] + H, pad[2]:pad[2] + W, :] = x
    Ho = (xp.shape[1] - kh) // stride + 1
    Wo = (xp.shape[2] - kw) // stride + 1
    cols = np.empty((B, Ho, Wo, kh, kw, Cin))
    for ky in range(kh):
        for kx in range(kw):
            cols[:, :, :, ky, kx, :] = xp[:, ky:ky + stride * Ho:stride, kx:kx + stride * Wo:stride, :]
    wm = np.transpose(w, (2, 3, 1, 0)).reshape(kh * kw * Cin, Cout)   # [(ky,kx,ci), co]
    y = cols.reshape(B * Ho * Wo, kh * kw * Cin) @ wm
    if b is not None:
        y = y + b
    return y.reshape(B, Ho, Wo, Cout)


def linear(x, w, b=None):
    y = x @ w.T
    return y if b is None else y + b


def softmax_rows(s):
    s = s - s.max(axis=-1, keepdims=True)
    e = np.exp(s)
    return e / e.sum(axis=-1, keepdims=True)


def attention(q, k, v, heads):
    """q [B,Nq,C], k/v [B,Nk,C]; per head softmax(q k^T / sqrt(d)) v, heads laid out as contiguous channel slices."""
    B, Nq, C = q.shape
    d = C // heads
    out = np.empty_like(q)
    for b in range(B):
        for h in range(heads):
            sl = slice(h * d, (h + 1) * d)
            p = softmax_rows(q[b, :, sl] @ k[b, :, sl].T / math.sqrt(d))
            out[b, :, sl] = p @ v[b, :, sl]
    return out


def upsample_nearest2x(x):
    B, H, W, C = x.shape
    idx_y = np.arange(2 * H) // 2
    idx_x = np.arange(2 * W) // 2
    return x[:, idx_y][:, :, idx_x]


def timestep_sincos(t, dim=320, flip_sin_to_cos=True, freq_shift=0.0):
    half = dim // 2
    freqs = np.exp(-math.log(10000.0) * np.arange(half) / (half - freq_shift))
    arg = np.asarray(t, dtype=np.float64)[:, None] * freqs[None, :]
    s, c = np.sin(arg), np.cos(arg)
    return np.concatenate([c, s], axis=1) if flip_sin_to_cos else np.concatenate([s, c], axis=1)


# ---------------------------------------------------------------------------------------------
# blocks, addressed by state-dict prefix
# ---------------------------------------------------------------------------------------------
class Net:
    """A state dict with float64 access by key."""

    def __init__(self, sd):
        self.sd = sd

    def __call__(self, key):
        return f64(self.sd[key])

    def has(self, key):
        return key in self.sd


def resnet(P: Net, k, x, temb=None, eps=1e-5, groups=32):
    h = silu(group_norm(x, P(k + ".norm1.weight"), P(k + ".norm1.bias"), groups, eps))
    h = conv2d(h, P(k + ".conv1.weight"), P(k + ".conv1.bias"))
    if temb is not None:
        h = h + linear(silu(temb), P(k + ".time_emb_proj.weight"), P(k + ".time_emb_proj.bias"))[:, None, None, :]
    h = silu(group_norm(h, P(k + ".norm2.weight"), P(k + ".norm2.bias"), groups, eps))
    h = conv2d(h, P(k + ".conv2.weight"), P(k + ".conv2.bias"))
    if P.has(k + ".conv_shortcut.weight"):
        x = conv2d(x, P(k + ".conv_shortcut.weight"), P(k + ".conv_shortcut.bias"), pad=(0, 0, 0, 0))
    return x + h


def transformer(P: Net, k, x, ctx, heads, groups=32):
    B, H, W, C = x.shape
    h = group_norm(x, P(k + ".norm.weight"), P(k + ".norm.bias"), groups, 1e-6).reshape(B, H * W, C)
    h = linear(h, P(k + ".proj_in.weight"), P(k + ".proj_in.bias"))
    b = k + ".transformer_blocks.0"
    n = layer_norm(h, P(b + ".norm1.weight"), P(b + ".norm1.bias"))
    a = attention(linear(n, P(b + ".attn1.to_q.weight")), linear(n, P(b + ".attn1.to_k.weight")),
                  linear(n, P(b + ".attn1.to_v.weight")), heads)
    h = h + linear(a, P(b + ".attn1.to_out.0.weight"), P(b + ".attn1.to_out.0.bias"))
    n = layer_norm(h, P(b + ".norm2.weight"), P(b + ".norm2.bias"))
    a = attention(linear(n, P(b + ".attn2.to_q.weight")), linear(ctx, P(b + ".attn2.to_k.weight")),
                  linear(ctx, P(b + ".attn2.to_v.weight")), heads)
    h = h + linear(a, P(b + ".attn2.to_out.0.weight"), P(b + ".attn2.to_out.0.bias"))
    n = layer_norm(h, P(b + ".norm3.weight"), P(b + ".norm3.bias"))
    u = linear(n, P(b + ".ff.net.0.proj.weight"), P(b + ".ff.net.0.proj.bias"))
    inner = u.shape[-1] // 2
    h = h + linear(u[..., :inner] * gelu_erf(u[..., inner:]), P(b + ".ff.net.2.weight"), P(b + ".ff.net.2.bias"))
    h = linear(h, P(k + ".proj_out.weight"), P(k + ".proj_out.bias"))
    return x + h.reshape(B, H, W, C)


def unet_forward(sd, sample_nchw, timesteps, ctx, block_out=(320, 640, 1280, 1280), heads=(5, 10, 20, 20),
                 attn_down=(True, True, True, False), layers=2, eps=1e-5):
    """SD2-inpainting UNet2DConditionModel.forward on a flat diffusers state dict; returns NCHW float64."""
    P = Net(sd)
    x = np.transpose(f64(sample_nchw), (0, 2, 3, 1))
    ctx = f64(ctx)
    B = x.shape[0]
    t = np.broadcast_to(np.asarray(timesteps, dtype=np.float64).reshape(-1), (B,))
    e = timestep_sincos(t, block_out[0])
    temb = linear(silu(linear(e, P("time_embedding.linear_1.weight"), P("time_embedding.linear_1.bias"))),
                  P("time_embedding.linear_2.weight"), P("time_embedding.linear_2.bias"))
    h = conv2d(x, P("conv_in.weight"), P("conv_in.bias"))
    stack = [h]
    nlev = len(block_out)
    for i in range(nlev):
        for j in range(layers):
            h = resnet(P, f"down_blocks.{i}.resnets.{j}", h, temb, eps)
            if attn_down[i]:
                h = transformer(P, f"down_blocks.{i}.attentions.{j}", h, ctx, heads[i])
            stack.append(h)
        if i != nlev - 1:
            h = conv2d(h, P(f"down_blocks.{i}.downsamplers.0.conv.weight"), P(f"down_blocks.{i}.downsamplers.0.conv.bias"),
                       stride=2)
            stack.append(h)
    h = resnet(P, "mid_block.resnets.0", h, temb, eps)
    h = transformer(P, "mid_block.attentions.0", h, ctx, heads[-1])
    h = resnet(P, "mid_block.resnets.1", h, temb, eps)
    for i in range(nlev):
        lev = nlev - 1 - i            # up block i mirrors down block `lev`
        for j in range(layers + 1):
            h = np.concatenate([h, stack.pop()], axis=-1)      # hidden first, skip second
            h = resnet(P, f"up_blocks.{i}.resnets.{j}", h, temb, eps)
            if attn_down[lev]:
                h = transformer(P, f"up_blocks.{i}.attentions.{j}", h, ctx, heads[lev])
        if i != nlev - 1:
            h = conv2d(upsample_nearest2x(h), P(f"up_blocks.{i}.upsamplers.0.conv.weight"),
                       P(f"up_blocks.{i}.upsamplers.0.conv.bias"))
    assert not stack
    h = silu(group_norm(h, P("conv_norm_out.weight"), P("conv_norm_out.bias"), 32, eps))
    return np.transpose(conv2d(h, P("conv_out.weight"), P("conv_out.bias")), (0, 3, 1, 2))


# ---------------------------------------------------------------------------------------------
# AutoencoderKL
# ---------------------------------------------------------------------------------------------
def _vae_attention(P: Net, k, x):
    B, H, W, C = x.shape
    n = group_norm(x, P(k + ".group_norm.weight"), P(k + ".group_norm.bias"), 32, 1e-6).reshape(B, H * W, C)
    q = linear(n, P(k + ".to_q.weight"), P(k + ".to_q.bias"))
    kk = linear(n, P(k + ".to_k.weight"), P(k + ".to_k.bias"))
    v = linear(n, P(k + ".to_v.weight"), P(k + ".to_v.bias"))
    a = attention(q, kk, v, 1)
    return x + linear(a, P(k + ".to_out.0.weight"), P(k + ".to_out.0.bias")).reshape(B, H, W, C)


def _vae_mid(P, side, h):
    h = resnet(P, f"{side}.mid_block.resnets.0", h, None, 1e-6)
    h = _vae_attention(P, f"{side}.mid_block.attentions.0", h)
    return resnet(P, f"{side}.mid_block.resnets.1", h, None, 1e-6)


def vae_encode_moments(sd, x_nchw, block_out=(128, 256, 512, 512), layers=2):
    """-> moments [B, 8, h, w] (mean | logvar), NCHW float64."""
    P = Net(sd)
    h = conv2d(np.transpose(f64(x_nchw), (0, 2, 3, 1)), P("encoder.conv_in.weight"), P("encoder.conv_in.bias"))
    for i in range(len(block_out)):
        for j in range(layers):
            h = resnet(P, f"encoder.down_blocks.{i}.resnets.{j}", h, None, 1e-6)
        if i != len(block_out) - 1:   # pad right/bottom by one, stride 2, no symmetric padding
            k = f"encoder.down_blocks.{i}.downsamplers.0.conv"
            h = conv2d(h, P(k + ".weight"), P(k + ".bias"), stride=2, pad=(0, 1, 0, 1))
    h = _vae_mid(P, "encoder", h)
    h = silu(group_norm(h, P("encoder.conv_norm_out.weight"), P("encoder.conv_norm_out.bias"), 32, 1e-6))
    h = conv2d(h, P("encoder.conv_out.weight"), P("encoder.conv_out.bias"))
    h = conv2d(h, P("quant_conv.weight"), P("quant_conv.bias"), pad=(0, 0, 0, 0))
    return np.transpose(h, (0, 3, 1, 2))


def vae_decode(sd, z_nchw, block_out=(128, 256, 512, 512), layers=2):
    P = Net(sd)
    h = conv2d(np.transpose(f64(z_nchw), (0, 2, 3, 1)), P("post_quant_conv.weight"), P("post_quant_conv.bias"),
               pad=(0, 0, 0, 0))
    h = conv2d(h, P("decoder.conv_in.weight"), P("decoder.conv_in.bias"))
    h = _vae_mid(P, "decoder", h)
    n = len(block_out)
    for i in range(n):
        for j in range(layers + 1):
            h = resnet(P, f"decoder.up_blocks.{i}.resnets.{j}", h, None, 1e-6)
        if i != n - 1:
            k = f"decoder.up_blocks.{i}.upsamplers.0.conv"
            h = conv2d(upsample_nearest2x(h), P(k + ".weight"), P(k + ".bias"))
    h = silu(group_norm(h, P("decoder.conv_norm_out.weight"), P("decoder.conv_norm_out.bias"), 32, 1e-6))
    return np.transpose(conv2d(h, P("decoder.conv_out.weight"), P("decoder.conv_out.bias")), (0, 3, 1, 2))


# ---------------------------------------------------------------------------------------------
# DDIM (eta = 0, epsilon prediction, SD2 schedule) straight from the DDIM paper's eq. 12
# ---------------------------------------------------------------------------------------------
def ddim_alphas(num_train=1000, beta_start=0.00085, beta_end=0.012):
    betas = np.linspace(math.sqrt(beta_start), math.sqrt(beta_end), num_train) ** 2
    return np.cumprod(1.0 - betas)


def ddim_timesteps(n, num_train=1000, steps_offset=1):
    ratio = num_train // n
    return [int(i * ratio + steps_offset) for i in range(n)][::-1]


def ddim_step(x, eps, t, n, acp=None, num_train=1000):
    acp = ddim_alphas() if acp is None else acp
    prev = t - num_train // n
    a_t = acp[t]
    a_p = acp[prev] if prev >= 0 else acp[0]      # set_alpha_to_one = False
    x0 = (x - math.sqrt(1.0 - a_t) * eps) / math.sqrt(a_t)
    return math.sqrt(a_p) * x0 + math.sqrt(1.0 - a_p) * eps
