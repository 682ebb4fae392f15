"""Parameter inventory of the two networks on the DiffUTE sampling path.

The engine does not instantiate nn.Modules; it works from this flat inventory
(diffusers state-dict key -> shape).  Key names are the ones diffusers'
`save_pretrained` writes (reference: train_diffute_v1.py:664-669 writes them,
app.ipynb:550-553 reads them), so a released DiffUTE checkpoint maps 1:1.
Totals are checked in tests: UNet 865,925,124; VAE 83,653,863.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

Shape = Tuple[int, ...]

SD2_INPAINT_UNET_CONFIG = dict(
    in_channels=9,
    out_channels=4,
    sample_size=64,
    block_out_channels=(320, 640, 1280, 1280),
    layers_per_block=2,
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
    attention_head_dim=(5, 10, 20, 20),
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

SD2_SCHEDULER_CONFIG = dict(
    num_train_timesteps=1000,
    beta_start=0.00085,
    beta_end=0.012,
    beta_schedule="scaled_linear",
    set_alpha_to_one=False,
    steps_offset=1,
    clip_sample=False,
    prediction_type="epsilon",
)


def _conv(d, k, cin, cout, ks):
    d[k + ".weight"] = (cout, cin, ks, ks)
    d[k + ".bias"] = (cout,)


def _lin(d, k, cin, cout, bias=True):
    d[k + ".weight"] = (cout, cin)
    if bias:
        d[k + ".bias"] = (cout,)


def _norm(d, k, c):
    d[k + ".weight"] = (c,)
    d[k + ".bias"] = (c,)


def _resnet(d, k, cin, cout, temb):
    _norm(d, k + ".norm1", cin)
    _conv(d, k + ".conv1", cin, cout, 3)
    if temb:
        _lin(d, k + ".time_emb_proj", temb, cout)
    _norm(d, k + ".norm2", cout)
    _conv(d, k + ".conv2", cout, cout, 3)
    if cin != cout:
        _conv(d, k + ".conv_shortcut", cin, cout, 1)


def _transformer(d, k, c, ctx):
    _norm(d, k + ".norm", c)
    _lin(d, k + ".proj_in", c, c)
    b = k + ".transformer_blocks.0"
    _norm(d, b + ".norm1", c)
    _lin(d, b + ".attn1.to_q", c, c, False)
    _lin(d, b + ".attn1.to_k", c, c, False)
    _lin(d, b + ".attn1.to_v", c, c, False)
    _lin(d, b + ".attn1.to_out.0", c, c)
    _norm(d, b + ".norm2", c)
    _lin(d, b + ".attn2.to_q", c, c, False)
    _lin(d, b + ".attn2.to_k", ctx, c, False)
    _lin(d, b + ".attn2.to_v", ctx, c, False)
    _lin(d, b + ".attn2.to_out.0", c, c)
    _norm(d, b + ".norm3", c)
    _lin(d, b + ".ff.net.0.proj", c, 8 * c)
    _lin(d, b + ".ff.net.2", 4 * c, c)
    _lin(d, k + ".proj_out", c, c)


def unet_layout(cfg=None):
    """Static block plan: returns dict with down/mid/up block descriptions (channel bookkeeping)."""
    cfg = cfg or SD2_INPAINT_UNET_CONFIG
    boc = list(cfg["block_out_channels"])
    L = cfg["layers_per_block"]
    down = []
    skip = [boc[0]]
    prev = boc[0]
    for i, (c, typ) in enumerate(zip(boc, cfg["down_block_types"])):
        last = i == len(boc) - 1
        down.append(dict(cin=prev, cout=c, attn=typ.startswith("CrossAttn"), down=not last, layers=L,
                         heads=cfg["attention_head_dim"][i]))
        skip += [c] * L + ([] if last else [c])
        prev = c
    up = []
    rboc = boc[::-1]
    rheads = list(cfg["attention_head_dim"])[::-1]
    prev = boc[-1]
    for i, (c, typ) in enumerate(zip(rboc, cfg["up_block_types"])):
        sk = [skip.pop() for _ in range(L + 1)]
        up.append(dict(prev=prev, cout=c, skips=sk, attn=typ.startswith("CrossAttn"), up=i != len(boc) - 1,
                       heads=rheads[i]))
        prev = c
    return dict(down=down, up=up, mid=dict(ch=boc[-1], heads=cfg["attention_head_dim"][-1]))


def unet_param_shapes(cfg=None) -> "OrderedDict[str, Shape]":
    cfg = cfg or SD2_INPAINT_UNET_CONFIG
    d: "OrderedDict[str, Shape]" = OrderedDict()
    boc = list(cfg["block_out_channels"])
    ctx = cfg["cross_attention_dim"]
    temb = boc[0] * 4
    lay = unet_layout(cfg)
    _conv(d, "conv_in", cfg["in_channels"], boc[0], 3)
    _lin(d, "time_embedding.linear_1", boc[0], temb)
    _lin(d, "time_embedding.linear_2", temb, temb)
    for i, b in enumerate(lay["down"]):
        for j in range(b["layers"]):
            _resnet(d, f"down_blocks.{i}.resnets.{j}", b["cin"] if j == 0 else b["cout"], b["cout"], temb)
            if b["attn"]:
                _transformer(d, f"down_blocks.{i}.attentions.{j}", b["cout"], ctx)
        if b["down"]:
            _conv(d, f"down_blocks.{i}.downsamplers.0.conv", b["cout"], b["cout"], 3)
    m = lay["mid"]["ch"]
    _resnet(d, "mid_block.resnets.0", m, m, temb)
    _transformer(d, "mid_block.attentions.0", m, ctx)
    _resnet(d, "mid_block.resnets.1", m, m, temb)
    for i, b in enumerate(lay["up"]):
        for j, sc in enumerate(b["skips"]):
            cin = (b["prev"] if j == 0 else b["cout"]) + sc
            _resnet(d, f"up_blocks.{i}.resnets.{j}", cin, b["cout"], temb)
            if b["attn"]:
                _transformer(d, f"up_blocks.{i}.attentions.{j}", b["cout"], ctx)
        if b["up"]:
            _conv(d, f"up_blocks.{i}.upsamplers.0.conv", b["cout"], b["cout"], 3)
    _norm(d, "conv_norm_out", boc[0])
    _conv(d, "conv_out", boc[0], cfg["out_channels"], 3)
    return d


def _vae_attn(d, k, c):
    _norm(d, k + ".group_norm", c)
    _lin(d, k + ".to_q", c, c)
    _lin(d, k + ".to_k", c, c)
    _lin(d, k + ".to_v", c, c)
    _lin(d, k + ".to_out.0", c, c)


def vae_param_shapes(cfg=None) -> "OrderedDict[str, Shape]":
    cfg = cfg or SD2_VAE_CONFIG
    d: "OrderedDict[str, Shape]" = OrderedDict()
    boc = list(cfg["block_out_channels"])
    L = cfg["layers_per_block"]
    lc = cfg["latent_channels"]
    _conv(d, "encoder.conv_in", cfg["in_channels"], boc[0], 3)
    prev = boc[0]
    for i, c in enumerate(boc):
        for j in range(L):
            _resnet(d, f"encoder.down_blocks.{i}.resnets.{j}", prev if j == 0 else c, c, None)
        if i != len(boc) - 1:
            _conv(d, f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c, 3)
        prev = c
    _resnet(d, "encoder.mid_block.resnets.0", prev, prev, None)
    _vae_attn(d, "encoder.mid_block.attentions.0", prev)
    _resnet(d, "encoder.mid_block.resnets.1", prev, prev, None)
    _norm(d, "encoder.conv_norm_out", prev)
    _conv(d, "encoder.conv_out", prev, 2 * lc, 3)
    _conv(d, "quant_conv", 2 * lc, 2 * lc, 1)
    _conv(d, "post_quant_conv", lc, lc, 1)
    rboc = boc[::-1]
    _conv(d, "decoder.conv_in", lc, rboc[0], 3)
    _resnet(d, "decoder.mid_block.resnets.0", rboc[0], rboc[0], None)
    _vae_attn(d, "decoder.mid_block.attentions.0", rboc[0])
    _resnet(d, "decoder.mid_block.resnets.1", rboc[0], rboc[0], None)
    prev = rboc[0]
    for i, c in enumerate(rboc):
        for j in range(L + 1):
            _resnet(d, f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else c, c, None)
        if i != len(boc) - 1:
            _conv(d, f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3)
        prev = c
    _norm(d, "decoder.conv_norm_out", prev)
    _conv(d, "decoder.conv_out", prev, cfg["out_channels"], 3)
    return d


def count(shapes: Dict[str, Shape]) -> int:
    n = 0
    for s in shapes.values():
        p = 1
        for x in s:
            p *= x
        n += p
    return n
