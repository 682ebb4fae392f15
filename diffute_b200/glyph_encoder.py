"""B200-native TrOCR glyph encoder: the ViT encoder of `trocr-large-printed` behind the transformers call signature.

Drop-in for the object the reference builds at app.ipynb:546-548
    trocr_model = VisionEncoderDecoderModel.from_pretrained('./trocr-large-printed').encoder.cuda()
and calls at app.ipynb:773-776 / train_diffute_v1.py:880-884
    ocr_feature = trocr_model(pixel_values)
    ocr_embeddings = ocr_feature.last_hidden_state            # [B, 577, 1024] = the UNet's encoder_hidden_states
It is a ViT (`ViTModel`): Conv2d(3 -> D, 16x16 / 16) patch embedding + CLS token + learned position embeddings, L pre-LN
blocks (LN -> MHSA(q,k,v; out proj) -> +; LN -> Linear(D->4D) -> erf-GELU -> Linear(4D->D) -> +), final LayerNorm.
Every contraction runs on the same tcgen05 kernels as the UNet: the patch embedding is a GEMM over a patchified operand
(dfu_patchify_f16) whose residual operand carries CLS / bias / position embeddings, q|k|v are one fused projection,
attention is the fused d=64 kernel (16 heads x 577 tokens), the MLP's GELU is a GEMM epilogue.
State-dict keys are transformers' ViTModel keys (a VisionEncoderDecoderModel checkpoint's `encoder.` prefix is stripped).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import ops
from .unet import Arena, _Config, _prec

TROCR_LARGE_VIT_CONFIG = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                              image_size=384, patch_size=16, num_channels=3, qkv_bias=False, layer_norm_eps=1e-12,
                              hidden_act="gelu")


def vit_param_shapes(cfg=None) -> Dict[str, tuple]:
    c = dict(TROCR_LARGE_VIT_CONFIG)
    c.update(cfg or {})
    D, I, P, C = c["hidden_size"], c["intermediate_size"], c["patch_size"], c["num_channels"]
    n_tok = 1 + (c["image_size"] // P) ** 2
    d = {"embeddings.cls_token": (1, 1, D), "embeddings.position_embeddings": (1, n_tok, D),
         "embeddings.patch_embeddings.projection.weight": (D, C, P, P), "embeddings.patch_embeddings.projection.bias": (D,)}
    for i in range(c["num_hidden_layers"]):
        b = f"encoder.layer.{i}"
        for n in ("query", "key", "value"):
            d[f"{b}.attention.attention.{n}.weight"] = (D, D)
            if c["qkv_bias"]:
                d[f"{b}.attention.attention.{n}.bias"] = (D,)
        d[f"{b}.attention.output.dense.weight"] = (D, D)
        d[f"{b}.attention.output.dense.bias"] = (D,)
        d[f"{b}.intermediate.dense.weight"] = (I, D)
        d[f"{b}.intermediate.dense.bias"] = (I,)
        d[f"{b}.output.dense.weight"] = (D, I)
        d[f"{b}.output.dense.bias"] = (D,)
        for n in ("layernorm_before", "layernorm_after"):
            d[f"{b}.{n}.weight"] = (D,)
            d[f"{b}.{n}.bias"] = (D,)
    d["layernorm.weight"] = (D,)
    d["layernorm.bias"] = (D,)
    return d


@dataclass
class GlyphEncoderOutput:
    last_hidden_state: torch.Tensor

    def __getitem__(self, k):
        if k in (0, "last_hidden_state"):
            return self.last_hidden_state
        raise KeyError(k)


class TrOCRGlyphEncoder:
    def __init__(self, state_dict: Dict[str, torch.Tensor], config: Optional[dict] = None, device="cuda",
                 precision="fp16x2"):
        cfg = dict(TROCR_LARGE_VIT_CONFIG)
        if config:
            cfg.update({k: v for k, v in config.items() if k in cfg})
        if cfg["hidden_act"] != "gelu":
            raise ValueError("only the exact-erf GELU MLP of the TrOCR ViT is implemented")
        if cfg["hidden_size"] != 64 * cfg["num_attention_heads"]:
            raise ValueError("attention head dim must be 64 (the fused attention kernel's tile)")
        self.config = _Config(cfg)
        self.device = torch.device(device)
        self.prec = _prec(precision)
        self.planes = ops.planes_of(self.prec)
        sd = dict(state_dict)
        if any(k.startswith("encoder.embeddings.") for k in sd):  # VisionEncoderDecoderModel checkpoint: encoder.<vit key>
            sd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
        shapes = vit_param_shapes(cfg)
        missing = [k for k in shapes if k not in sd]
        if missing:
            raise KeyError(f"ViT state dict is missing {len(missing)} keys, e.g. {missing[:3]}")
        for k, s in shapes.items():
            if tuple(sd[k].shape) != tuple(s):
                raise ValueError(f"{k}: expected shape {s}, got {tuple(sd[k].shape)}")
        self.arena = Arena(self.device)
        self.ws = ops.Workspace(64 << 20, self.device)
        self._pack(sd)

    @classmethod
    def from_pretrained(cls, path, **kw):
        """`path`: a transformers folder (config.json + model.safetensors | pytorch_model.bin) of a ViTModel or of the
        VisionEncoderDecoderModel the reference loads (its `encoder` sub-config / `encoder.` keys are used)."""
        with open(os.path.join(path, "config.json")) as f:
            cfg = json.load(f)
        cfg = cfg.get("encoder", cfg)
        st = os.path.join(path, "model.safetensors")
        if os.path.isfile(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu", weights_only=True)
        return cls({k: v.float() for k, v in sd.items()}, cfg, **kw)

    def _pack(self, sd):
        cfg, P = self.config, self.planes
        pk = ops.Packer(self.device)
        w: Dict[str, torch.Tensor] = {}
        self.w = w
        D = cfg["hidden_size"]
        w["patch.w16"] = pk.weight16(sd["embeddings.patch_embeddings.projection.weight"].reshape(D, -1), P)
        # residual operand of the patch-embedding GEMM: row 0 = cls + pos[0], row t = conv bias + pos[t]
        pos = sd["embeddings.position_embeddings"][0].float()
        emb = pos + sd["embeddings.patch_embeddings.projection.bias"].float()[None, :]
        emb[0] = pos[0] + sd["embeddings.cls_token"].float().reshape(-1)
        w["embed.res"] = pk.f32(emb)
        for i in range(cfg["num_hidden_layers"]):
            b = f"encoder.layer.{i}"
            a = f"{b}.attention.attention"
            w[f"{b}.qkv.w16"] = torch.empty((P * 3 * D, D), dtype=torch.float16, device=self.device)
            for j, n in enumerate(("query", "key", "value")):
                pk.weight16(sd[f"{a}.{n}.weight"], P, into=w[f"{b}.qkv.w16"], row0=j * D, total_rows=3 * D)
            if cfg["qkv_bias"]:
                w[f"{b}.qkv.b"] = torch.empty((3 * D,), dtype=torch.float32, device=self.device)
                for j, n in enumerate(("query", "key", "value")):
                    pk.f32(sd[f"{a}.{n}.bias"].reshape(-1, 1), into=w[f"{b}.qkv.b"], row0=j * D)
            for name, key in (("proj", f"{b}.attention.output.dense"), ("fc1", f"{b}.intermediate.dense"),
                              ("fc2", f"{b}.output.dense")):
                w[f"{b}.{name}.w16"] = pk.weight16(sd[key + ".weight"], P)
                w[f"{b}.{name}.b"] = pk.f32(sd[key + ".bias"])
            for n in ("layernorm_before", "layernorm_after"):
                w[f"{b}.{n}.g"] = pk.f32(sd[f"{b}.{n}.weight"])
                w[f"{b}.{n}.b"] = pk.f32(sd[f"{b}.{n}.bias"])
        w["ln.g"] = pk.f32(sd["layernorm.weight"])
        w["ln.b"] = pk.f32(sd["layernorm.bias"])
        pk.run()

    # nn.Module-ish no-ops the reference touches (app.ipynb:547 `.cuda()`, :557 `.requires_grad_(False)`)
    def cuda(self, *a):
        return self

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def requires_grad_(self, flag=False):
        return self

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor, return_dict: bool = True, **unused):
        cfg, w, P, prec, A = self.config, self.w, self.planes, self.prec, self.arena
        if pixel_values.dim() != 4 or pixel_values.shape[1] != cfg["num_channels"] or \
                pixel_values.shape[2] != cfg["image_size"] or pixel_values.shape[3] != cfg["image_size"]:
            raise ValueError(f"pixel_values must be [B,{cfg['num_channels']},{cfg['image_size']},{cfg['image_size']}], "
                             f"got {tuple(pixel_values.shape)}")
        x = pixel_values.to(device=self.device, dtype=torch.float32).contiguous()
        B = x.shape[0]
        D, I, heads, ps = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_attention_heads"], cfg["patch_size"]
        T = 1 + (cfg["image_size"] // ps) ** 2
        M = B * T
        eps = float(cfg["layer_norm_eps"])
        op16 = lambda name, shape: A.get(name, (P, *shape), torch.float16)
        patches = op16("patches", (M, cfg["num_channels"] * ps * ps))
        ops.patchify_f16(x, ps, patches)
        res = A.get("embed.res", (B, T, D))
        res.copy_(w["embed.res"].unsqueeze(0).expand(B, T, D))
        h = A.get("h0", (M, D))
        ops.linear(patches, w["patch.w16"], D, prec, ws=self.ws, out_f32=h, residual=res.view(M, D))
        for i in range(cfg["num_hidden_layers"]):
            b = f"encoder.layer.{i}"
            ln = op16("ln", (M, D))
            ops.layernorm(h, w[f"{b}.layernorm_before.g"], w[f"{b}.layernorm_before.b"], eps, ln)
            qkv = op16("qkv", (M, 3 * D))
            ops.linear(ln, w[f"{b}.qkv.w16"], 3 * D, prec, ws=self.ws, out_f16=qkv, bias=w.get(f"{b}.qkv.b"))
            at = op16("attn", (M, D))
            ops.attention(qkv, 0, qkv, D, qkv, 2 * D, B, heads, T, T, 0.125, at, ws=self.ws)
            h1 = A.get("h1", (M, D))
            ops.linear(at, w[f"{b}.proj.w16"], D, prec, ws=self.ws, out_f32=h1, bias=w[f"{b}.proj.b"], residual=h)
            ops.layernorm(h1, w[f"{b}.layernorm_after.g"], w[f"{b}.layernorm_after.b"], eps, ln)
            mid = op16("mlp", (M, I))
            ops.linear(ln, w[f"{b}.fc1.w16"], I, prec, ws=self.ws, out_f16=mid, bias=w[f"{b}.fc1.b"], gelu=True)
            ops.linear(mid, w[f"{b}.fc2.w16"], D, prec, ws=self.ws, out_f32=h, bias=w[f"{b}.fc2.b"], residual=h1)
        out = torch.empty((B, T, D), dtype=torch.float32, device=self.device)
        ops.layernorm(h, w["ln.g"], w["ln.b"], eps, out32=out.view(M, D))
        return GlyphEncoderOutput(out) if return_dict else (out,)

    __call__ = forward


class _ProcessorOutput(dict):
    """BatchFeature-like: `.pixel_values` and `["pixel_values"]`."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class TrOCRGlyphProcessor:
    """Image side of `TrOCRProcessor` (app.ipynb:546, 773-774: `processor(images=ttf_imgs, return_tensors="pt")
    .pixel_values`) on the GPU: each uint8 glyph image goes to the device once and `dfu_glyph_preprocess` produces its
    [3, 384, 384] slice of `pixel_values` -- PIL's antialiased bilinear resize, rescale 1/255 and normalise 0.5 / 0.5,
    bit-identical to transformers' PIL-backend ViTImageProcessor (trocr-large-printed's preprocessor_config.json:
    size 384, resample 2, image_mean = image_std = 0.5).  The text side (tokenizer) is not on the sampling path."""

    def __init__(self, size: int = 384, device="cuda"):
        self.size, self.device = int(size), torch.device(device)

    def __call__(self, images, return_tensors: str = "pt", **unused):
        import numpy as np
        if return_tensors != "pt":
            raise NotImplementedError("return_tensors='pt' only")
        if not isinstance(images, (list, tuple)):
            images = [images]
        out = torch.empty((len(images), 3, self.size, self.size), dtype=torch.float32, device=self.device)
        for i, im in enumerate(images):
            a = im if isinstance(im, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(im)))
            if a.dim() == 2:  # PIL 'L' image: ViTImageProcessor converts to RGB
                a = a[:, :, None].expand(-1, -1, 3)
            if a.dtype != torch.uint8 or a.dim() != 3 or a.shape[2] != 3:
                raise ValueError("glyph images must be uint8 [h, w, 3] (PIL RGB image, numpy array or tensor)")
            ops.glyph_preprocess(a.to(self.device).contiguous(), out[i])
        return _ProcessorOutput(pixel_values=out)
