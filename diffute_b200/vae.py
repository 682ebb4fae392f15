"""B200-native AutoencoderKL (SD2 VAE architecture) behind the diffusers call signatures.

Drop-in for the object built at app.ipynb:550 / train_diffute_v1.py:632 and called as
    vae.encode(x).latent_dist.sample() * vae.config.scaling_factor      (app.ipynb:781-782, :793-794)
    vae.decode(latents / vae.config.scaling_factor).sample               (app.ipynb:818-819)
    vae(x)["sample"]                                                     (train_vae.py:721-722)
All device arithmetic goes through libdiffute_b200.so; math follows SURVEY.md Appendix A.2.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import arch, ops
from .unet import Arena, _Config, _prec


class DiagonalGaussianDistribution:
    """diffusers' DiagonalGaussianDistribution over NCHW moments [B, 2*C, h, w] held on the device."""

    def __init__(self, moments: torch.Tensor):
        self.parameters = moments
        self._c = moments.shape[1] // 2

    @property
    def mean(self):
        return self.parameters[:, : self._c]

    @property
    def logvar(self):
        return torch.clamp(self.parameters[:, self._c:], -30.0, 20.0)

    @property
    def std(self):
        return torch.exp(0.5 * self.logvar)

    @property
    def var(self):
        return torch.exp(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None, noise: Optional[torch.Tensor] = None,
               scale: float = 1.0) -> torch.Tensor:
        """mean + std * eps.  Like the reference (app.ipynb:439-480 randn_tensor) eps is drawn where the generator
        lives (CPU generator -> CPU draw -> copy), so seeded runs match a CPU run bit for bit in eps."""
        B, C2, h, w = self.parameters.shape
        if noise is None:
            dev = generator.device if generator is not None else self.parameters.device
            noise = torch.randn((B, self._c, h, w), generator=generator, device=dev, dtype=torch.float32)
        noise = noise.to(self.parameters.device, torch.float32).contiguous()
        z = torch.empty((B, self._c, h, w), device=self.parameters.device, dtype=torch.float32)
        ops.gaussian_sample(self.parameters, noise, scale, z)
        return z

    def mode(self, scale: float = 1.0) -> torch.Tensor:
        B, C2, h, w = self.parameters.shape
        z = torch.empty((B, self._c, h, w), device=self.parameters.device, dtype=torch.float32)
        ops.gaussian_sample(self.parameters, None, scale, z)
        return z


@dataclass
class AutoencoderKLOutput:
    latent_dist: DiagonalGaussianDistribution

    def __getitem__(self, k):
        if k in (0, "latent_dist"):
            return self.latent_dist
        raise KeyError(k)


@dataclass
class DecoderOutput:
    sample: torch.Tensor

    def __getitem__(self, k):
        if k in (0, "sample"):
            return self.sample
        raise KeyError(k)


class AutoencoderKL:
    def __init__(self, state_dict: Dict[str, torch.Tensor], config: Optional[dict] = None, device="cuda",
                 precision="fp16x2", encoder_precision=None):
        """precision: contraction mode of the decoder (and of the encoder unless `encoder_precision` is given).
        The decoded RGB is what the 1e-3 parity bar is measured on: the decoder keeps the 3-pass hi/lo split
        (fp16x2), while the encoder may run one fp16 pass — measured at 512x512 / 50 steps the decoded-RGB error is
        3.1e-4 with an fp16 encoder vs 3.3e-4 with fp16x2 (UNet-dominated either way; scripts/parity_vae_split.py)."""
        from .checkpoint import remap_legacy_vae_keys
        cfg = dict(arch.SD2_VAE_CONFIG)
        if config:
            cfg.update({k: v for k, v in config.items() if not k.startswith("_")})
        self.config = _Config(cfg)
        self.device = torch.device(device)
        self.dec_prec = _prec(precision)
        self.enc_prec = _prec(encoder_precision) if encoder_precision is not None else self.dec_prec
        self._use(self.dec_prec)
        self.dtype = torch.float32
        sd = remap_legacy_vae_keys(state_dict)
        shapes = arch.vae_param_shapes(cfg)
        missing = [k for k in shapes if k not in sd]
        if missing:
            raise KeyError(f"VAE state dict is missing {len(missing)} keys, e.g. {missing[:3]}")
        for k, s in shapes.items():
            t = sd[k]
            if t.dim() == 4 and len(s) == 2:  # <=0.16 checkpoints store the attention projections as 1x1 convs
                t = sd[k] = t.reshape(s)
            if tuple(t.shape) != tuple(s):
                raise ValueError(f"{k}: expected shape {s}, got {tuple(t.shape)}")
        self.arena = Arena(self.device)
        self.ws = ops.Workspace(256 << 20, self.device)
        self._sd = {k: sd[k].detach().to(torch.float32) for k in shapes}
        self._pack(self._sd)

    @classmethod
    def from_pretrained(cls, path, subfolder: Optional[str] = "vae", revision=None, **kw):
        from .checkpoint import load_diffusers_folder
        cfg, sd = load_diffusers_folder(path, subfolder)
        return cls(sd, cfg, **kw)

    @classmethod
    def from_synthetic(cls, seed: int = 1234, **kw):
        from . import synthetic
        return cls(synthetic.make_state_dict(arch.vae_param_shapes(), seed), **kw)

    # nn.Module-ish no-ops used by the reference scripts
    def eval(self):
        return self

    def requires_grad_(self, flag=False):
        return self

    def to(self, *a, **k):
        return self

    def cuda(self, *a):
        return self

    def _use(self, prec):
        """select the contraction mode of the half (encoder / decoder) that runs next"""
        self.prec = prec
        self.planes = ops.planes_of(prec)

    def _planes_for(self, key: str) -> int:
        enc = key.startswith("encoder.") or key.startswith("quant_conv")
        return ops.planes_of(self.enc_prec if enc else self.dec_prec)

    def _pack(self, sd):
        """diffusers state dict -> kernel layouts, ONE dfu_pack_weights launch (see ops.Packer)."""
        w: Dict[str, torch.Tensor] = {}
        self.w = w
        pk = ops.Packer(self.device)
        for k, t in sd.items():
            P = self._planes_for(k)
            if k.endswith(".weight") and t.dim() == 4 and t.shape[0] > 8 and t.shape[1] > 16:
                w[k[:-7] + ".w16"] = pk.weight16(t, P)     # tensor-core convs
            elif k.endswith(".weight") and t.dim() == 2:
                w[k[:-7] + ".w16"] = pk.weight16(t, P)     # attention projections
            elif t.dim() == 1:
                w[k] = pk.f32(t)                            # biases, GroupNorm affine parameters
        # fused 1x1 shortcut bias folds into conv2's bias
        for k in sd:
            if k.endswith(".conv_shortcut.bias"):
                base = k[: -len(".conv_shortcut.bias")]
                w[base + ".conv2.bias_sc"] = pk.f32(sd[base + ".conv2.bias"], add=sd[k])
        for side in ("encoder", "decoder"):
            a = f"{side}.mid_block.attentions.0"
            P = self._planes_for(a)
            C = sd[f"{a}.to_q.weight"].shape[0]
            w[a + ".qkv.w16"] = torch.empty((P * 3 * C, C), dtype=torch.float16, device=self.device)
            w[a + ".qkv.b"] = torch.empty((3 * C,), dtype=torch.float32, device=self.device)
            for i, x in enumerate("qkv"):
                pk.weight16(sd[f"{a}.to_{x}.weight"], P, into=w[a + ".qkv.w16"], row0=i * C, total_rows=3 * C)
                pk.f32(sd[f"{a}.to_{x}.bias"].reshape(-1, 1), into=w[a + ".qkv.b"], row0=i * C)
        w["encoder.conv_out.wp"] = pk.small_out(sd["encoder.conv_out.weight"])
        w["decoder.conv_out.wp"] = pk.small_out(sd["decoder.conv_out.weight"])
        # decoder output conv (128 -> 3 over the full-resolution map, 1.8 GFLOP at 512x512) on the tensor cores: output
        # channels zero-padded to one 32-column tile; a 1x1 "selection" pass then adds the bias and writes NCHW
        wo = sd["decoder.conv_out.weight"]
        Pd = ops.planes_of(self.dec_prec)
        w["decoder.conv_out.w16"] = torch.zeros((Pd * 32, wo.shape[1] * 9), dtype=torch.float16, device=self.device)
        pk.weight16(wo, Pd, into=w["decoder.conv_out.w16"], row0=0, total_rows=32)
        sel = torch.zeros((wo.shape[0], 1, 32), dtype=torch.float32)
        for i in range(wo.shape[0]):
            sel[i, 0, i] = 1.0
        w["decoder.conv_out.sel"] = sel.to(self.device)
        for k in ("encoder.conv_in", "post_quant_conv", "decoder.conv_in"):
            w[k + ".wt"] = pk.small_in(sd[k + ".weight"])
        qw = sd["quant_conv.weight"]
        w["quant_conv.w2"] = pk.f32(qw.reshape(qw.shape[0], -1))
        pk.run()

    # nn.Module surface (train_vae.py / train_diffute_v1.py touch these on the frozen VAE)
    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: self._sd[k] for k in arch.vae_param_shapes(self.config)}

    def parameters(self):
        return iter(self.state_dict().values())

    def save_pretrained(self, save_directory: str, safe_serialization: bool = True, **kw):
        from .checkpoint import save_diffusers_folder
        save_diffusers_folder(save_directory, None, dict(self.config), self.state_dict(), "AutoencoderKL",
                              safe_serialization=safe_serialization)

    # ------------------------------------------------------------------------------------------
    def _op16(self, name, shape):
        return self.arena.get(name, (self.planes, *shape), torch.float16)

    def _fp32(self, avoid, shape):
        """one of three rotating fp32 activation buffers not in `avoid`"""
        for n in ("v.0", "v.1", "v.2"):
            if all(self.arena.bufs.get(n) is None or a is None or
                   self.arena.bufs[n].data_ptr() != a.untyped_storage().data_ptr() for a in avoid):
                return self.arena.get(n, shape)
        raise RuntimeError("no free activation buffer")

    def _resnet(self, k, x, cout):
        w, prec, P = self.w, self.prec, self.planes
        B, H, W, cin = x.shape
        has_sc = (k + ".conv_shortcut.w16") in w
        a16 = self._op16("opA", (B, H, W, cin))
        raw16 = self._op16("opRaw", (B, H, W, cin)) if has_sc else None
        ops.groupnorm(x, w[k + ".norm1.weight"], w[k + ".norm1.bias"], 1e-6, True, prec, out16=a16, raw16=raw16,
                      ws=self.ws)
        hmid = self._fp32([x], (B, H, W, cout))
        ops.conv(a16.view(P * B, H, W, cin), w[k + ".conv1.w16"], cout, prec, (B, H, W), ops.taps_3x3_s1(), ws=self.ws,
                 out_f32=hmid.view(-1, cout), bias=w[k + ".conv1.bias"])
        b16 = self._op16("opA", (B, H, W, cout))
        ops.groupnorm(hmid, w[k + ".norm2.weight"], w[k + ".norm2.bias"], 1e-6, True, prec, out16=b16, ws=self.ws)
        out = self._fp32([x, hmid], (B, H, W, cout))
        if has_sc:
            ops.conv(b16.view(P * B, H, W, cout), w[k + ".conv2.w16"], cout, prec, (B, H, W), ops.taps_3x3_s1(),
                     shortcut=(raw16.view(P * B, H, W, cin), w[k + ".conv_shortcut.w16"]), ws=self.ws,
                     out_f32=out.view(-1, cout), bias=w[k + ".conv2.bias_sc"])
        else:
            ops.conv(b16.view(P * B, H, W, cout), w[k + ".conv2.w16"], cout, prec, (B, H, W), ops.taps_3x3_s1(),
                     ws=self.ws, out_f32=out.view(-1, cout), bias=w[k + ".conv2.bias"], residual=x.view(-1, cout))
        return out

    def _attention(self, k, x):
        """Single-head d=C spatial self-attention: S = Q K^T on the contraction core, fp32 row softmax,
        O = P V with V transposed once; all samples of the batch in one launch per stage (N = H*W keys per sample)."""
        w, prec, P = self.w, self.prec, self.planes
        B, H, W, C = x.shape
        N = H * W
        g16 = self._op16("opA", (B * N, C))
        ops.groupnorm(x, w[k + ".group_norm.weight"], w[k + ".group_norm.bias"], 1e-6, False, prec,
                      out16=g16.view(P, B, H, W, C), ws=self.ws)
        qkv = self._op16("qkv", (B * N, 3 * C))
        ops.linear(g16, w[k + ".qkv.w16"], 3 * C, prec, ws=self.ws, out_f16=qkv, bias=w[k + ".qkv.b"])
        # every sample's S = Q K^T, row softmax and O = P V in ONE batched launch each (no per-sample host loop):
        # product b reads rows [b*N, (b+1)*N) of the fused q|k|v matrix; V is transposed once per sample into [C, N]
        s = self.arena.get("att.s", (B * N, N))
        p16 = self._op16("att.p", (B * N, N))
        vt = self._op16("att.vt", (B * C, N))
        o16 = self._op16("att.o", (B * N, C))
        q, kk, v = qkv[:, :, 0:C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:3 * C]
        ops.linear(q, kk, N, prec, ws=self.ws, batch=B, a_batch_rows=N, b_batch_rows=N, out_f32=s)
        ops.softmax_rows(s, C ** -0.5, p16)
        for pl in range(P):
            ops.transpose_f16(v[pl].view(B, N, C), vt[pl].view(B, C, N))
        ops.linear(p16, vt, C, prec, ws=self.ws, batch=B, a_batch_rows=N, b_batch_rows=C, out_f16=o16)
        out = self._fp32([x], (B, H, W, C))
        ops.linear(o16, w[k + ".to_out.0.w16"], C, prec, ws=self.ws, out_f32=out.view(-1, C),
                   bias=w[k + ".to_out.0.bias"], residual=x.view(-1, C))
        return out

    def _mid(self, side, h):
        C = h.shape[-1]
        h = self._resnet(f"{side}.mid_block.resnets.0", h, C)
        h = self._attention(f"{side}.mid_block.attentions.0", h)
        return self._resnet(f"{side}.mid_block.resnets.1", h, C)

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        self._use(self.enc_prec)
        cfg, w, P = self.config, self.w, self.planes
        if x.dim() != 4 or x.shape[1] != cfg["in_channels"]:
            raise ValueError(f"encode expects [B,{cfg['in_channels']},H,W], got {tuple(x.shape)}")
        B, _, H, W = x.shape
        if H % 8 or W % 8:
            raise ValueError("image height/width must be multiples of 8")
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        boc = list(cfg["block_out_channels"])
        h = self._fp32([], (B, H, W, boc[0]))
        ops.conv_small_in([x], w["encoder.conv_in.wt"], w["encoder.conv_in.bias"], h, B)
        for i, c in enumerate(boc):
            for j in range(cfg["layers_per_block"]):
                h = self._resnet(f"encoder.down_blocks.{i}.resnets.{j}", h, c)
            if i != len(boc) - 1:
                Bh, Hh, Wh, Ch = h.shape
                s16 = self._op16("opA", (4 * B, Hh // 2, Wh // 2, Ch))
                ops.cast_f16(h, ops.CAST_S2D, s16)
                out = self._fp32([h], (B, Hh // 2, Wh // 2, Ch))
                k = f"encoder.down_blocks.{i}.downsamplers.0.conv"
                ops.conv(s16.view(P * 4 * B, Hh // 2, Wh // 2, Ch), w[k + ".w16"], Ch, self.prec,
                         (B, Hh // 2, Wh // 2), ops.taps_3x3_s2(B, 0), imgs_per_plane=4 * B, ws=self.ws,
                         out_f32=out.view(-1, Ch), bias=w[k + ".bias"])
                h = out
        h = self._mid("encoder", h)
        Bh, Hh, Wh, Ch = h.shape
        o32 = self._fp32([h], (Bh, Hh, Wh, Ch))
        ops.groupnorm(h, w["encoder.conv_norm_out.weight"], w["encoder.conv_norm_out.bias"], 1e-6, True, self.prec,
                      out32=o32, ws=self.ws)
        moments = torch.empty((B, 2 * cfg["latent_channels"], Hh, Wh), device=self.device, dtype=torch.float32)
        ops.conv_small_out(o32, w["encoder.conv_out.wp"], w["encoder.conv_out.bias"], moments, w2=w["quant_conv.w2"],
                           b2=w["quant_conv.bias"])
        dist = DiagonalGaussianDistribution(moments)
        return AutoencoderKLOutput(dist) if return_dict else (dist,)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True, pre_scale: float = 1.0):
        """z [B, latent_channels, h, w] (already divided by scaling_factor by the caller, app.ipynb:818), or pass
        pre_scale = 1/scaling_factor to fold that division into the first kernel."""
        self._use(self.dec_prec)
        cfg, w, P = self.config, self.w, self.planes
        if z.dim() != 4 or z.shape[1] != cfg["latent_channels"]:
            raise ValueError(f"decode expects [B,{cfg['latent_channels']},h,w], got {tuple(z.shape)}")
        z = z.to(device=self.device, dtype=torch.float32).contiguous()
        B, lc, hh, ww = z.shape
        boc = list(cfg["block_out_channels"])[::-1]
        zq = self.arena.get("dec.zq", (B, hh, ww, lc))
        ops.conv_small_in([z], w["post_quant_conv.wt"], w["post_quant_conv.bias"], zq, B, pre_scale=pre_scale)
        h = self._fp32([], (B, hh, ww, boc[0]))
        ops.conv_small_in([zq], w["decoder.conv_in.wt"], w["decoder.conv_in.bias"], h, B, nhwc=True)
        h = self._mid("decoder", h)
        for i, c in enumerate(boc):
            for j in range(cfg["layers_per_block"] + 1):
                h = self._resnet(f"decoder.up_blocks.{i}.resnets.{j}", h, c)
            if i != len(boc) - 1:
                Bh, Hh, Wh, Ch = h.shape
                u16 = self._op16("opA", (B, 2 * Hh, 2 * Wh, Ch))
                ops.cast_f16(h, ops.CAST_UP2X, u16)
                out = self._fp32([h], (B, 2 * Hh, 2 * Wh, Ch))
                k = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                ops.conv(u16.view(P * B, 2 * Hh, 2 * Wh, Ch), w[k + ".w16"], Ch, self.prec, (B, 2 * Hh, 2 * Wh),
                         ops.taps_3x3_s1(), ws=self.ws, out_f32=out.view(-1, Ch), bias=w[k + ".bias"])
                h = out
        Bh, Hh, Wh, Ch = h.shape
        img = torch.empty((B, cfg["out_channels"], Hh, Wh), device=self.device, dtype=torch.float32)
        if Ch % 64 == 0 and cfg["out_channels"] <= 8:
            o16 = self._op16("opA", (B, Hh, Wh, Ch))
            ops.groupnorm(h, w["decoder.conv_norm_out.weight"], w["decoder.conv_norm_out.bias"], 1e-6, True, self.prec,
                          out16=o16, ws=self.ws)
            t32 = self._fp32([h], (Bh, Hh, Wh, 32))
            ops.conv(o16.view(P * B, Hh, Wh, Ch), w["decoder.conv_out.w16"], 32, self.prec, (B, Hh, Wh), ops.taps_3x3_s1(),
                     ws=self.ws, out_f32=t32.view(-1, 32))
            ops.conv_small_out(t32, w["decoder.conv_out.sel"], w["decoder.conv_out.bias"], img)
        else:
            o32 = self._fp32([h], (Bh, Hh, Wh, Ch))
            ops.groupnorm(h, w["decoder.conv_norm_out.weight"], w["decoder.conv_norm_out.bias"], 1e-6, True, self.prec,
                          out32=o32, ws=self.ws)
            ops.conv_small_out(o32, w["decoder.conv_out.wp"], w["decoder.conv_out.bias"], img)
        return DecoderOutput(img) if return_dict else (img,)

    @torch.no_grad()
    def forward(self, sample, sample_posterior: bool = False, return_dict: bool = True, generator=None):
        post = self.encode(sample).latent_dist
        z = post.sample(generator) if sample_posterior else post.mode()
        return self.decode(z, return_dict=return_dict)

    __call__ = forward
