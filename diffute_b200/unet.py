"""B200-native UNet2DConditionModel (SD2-inpainting architecture) behind the diffusers call signature.

Drop-in for the object the reference builds at app.ipynb:551-553 / train_diffute_v1.py:633-635 and calls at
app.ipynb:814 / train_diffute_v1.py:913:

    noise_pred = unet(latent_model_input, t, ocr_embeddings).sample

`forward` launches only kernels from libdiffute_b200.so (tcgen05 implicit-GEMM convs / linears, fused attention,
GroupNorm / LayerNorm / cast kernels); torch supplies device memory, streams and CUDA-graph capture.  The residual
stream stays fp32 NHWC in HBM; every contraction operand is written once as fp16 (hi [, lo]) by the kernel that
produces it.  Math follows SURVEY.md Appendix A.1; state-dict key names are diffusers'.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import arch, ops
from .ops import PREC_FP16, PREC_FP16X2


@dataclass
class UNet2DConditionOutput:
    sample: torch.Tensor

    def __getitem__(self, k):
        if k in (0, "sample"):
            return self.sample
        raise KeyError(k)

    def to_tuple(self):
        return (self.sample,)


class _Config(dict):
    """dict with attribute access, like diffusers' FrozenDict (`unet.config.in_channels`, `unet.config["..."]`)."""
    __getattr__ = dict.__getitem__


def _prec(precision) -> int:
    if precision in (PREC_FP16, "fp16"):
        return PREC_FP16
    if precision in (PREC_FP16X2, "fp16x2", "fp32", "parity"):
        return PREC_FP16X2
    raise ValueError(f"unknown precision {precision!r}")


class Arena:
    """Named static device buffers: the same name always returns the same storage, so a captured CUDA graph replays
    against fixed addresses and no allocation happens on the sampling path after the first call."""

    def __init__(self, device):
        self.device = device
        self.generation = 0  # bumped whenever a buffer is (re)allocated: graphs captured before are stale
        self.bufs: Dict[str, torch.Tensor] = {}

    def get(self, name: str, shape, dtype=torch.float32) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        t = self.bufs.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(max(n, 1), dtype=dtype, device=self.device)
            self.bufs[name] = t
            self.generation += 1
        return t[:n].view(*shape)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


class UNet2DConditionModel:
    def __init__(self, state_dict: Dict[str, torch.Tensor], config: Optional[dict] = None, device="cuda",
                 precision="fp16", use_cuda_graph: bool = True):
        cfg = dict(arch.SD2_INPAINT_UNET_CONFIG)
        if config:
            cfg.update({k: v for k, v in config.items() if not k.startswith("_")})
        boc = list(cfg["block_out_channels"])
        if isinstance(cfg["attention_head_dim"], int):  # diffusers accepts one int for all blocks
            cfg["attention_head_dim"] = (cfg["attention_head_dim"],) * len(boc)
        if len(cfg["attention_head_dim"]) != len(boc):
            raise ValueError("attention_head_dim must be an int or one entry per block")
        self.config = _Config(cfg)
        self.device = torch.device(device)
        self.prec = _prec(precision)
        self.planes = ops.planes_of(self.prec)
        self.use_cuda_graph = use_cuda_graph
        self.dtype = torch.float32
        shapes = arch.unet_param_shapes(cfg)
        missing = [k for k in shapes if k not in state_dict]
        if missing:
            raise KeyError(f"UNet state dict is missing {len(missing)} keys, e.g. {missing[:3]}")
        for k, s in shapes.items():
            if tuple(state_dict[k].shape) != tuple(s):
                raise ValueError(f"{k}: expected shape {s}, got {tuple(state_dict[k].shape)}")
        self.layout = arch.unet_layout(cfg)
        for blk in self.layout["down"] + [dict(self.layout["mid"], cout=self.layout["mid"]["ch"], attn=True)]:
            if blk.get("attn") and blk["cout"] != 64 * blk["heads"]:
                raise ValueError(f"attention head dim must be 64 (the fused attention kernel's tile): got {blk['cout']} "
                                 f"channels over {blk['heads']} heads; attention_head_dim holds head COUNTS per block")
        self.ws = ops.Workspace(256 << 20, self.device)
        self.arena = Arena(self.device)
        self._graphs = {}
        self._ctx_key = None
        self._weights_generation = 0
        self._sd = {k: state_dict[k].detach().to(torch.float32) for k in shapes}
        self._pack(self._sd)

    # ------------------------------------------------------------------------------------------
    # loading
    # ------------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, path, subfolder: Optional[str] = "unet", revision=None, **kw):
        from .checkpoint import load_diffusers_folder
        cfg, sd = load_diffusers_folder(path, subfolder)
        return cls(sd, cfg, **kw)

    @classmethod
    def from_synthetic(cls, seed: int = 1234, **kw):
        from . import synthetic
        return cls(synthetic.make_state_dict(arch.unet_param_shapes(), seed), **kw)

    def _pack(self, sd):
        """diffusers state dict -> kernel layouts, ONE dfu_pack_weights launch for the whole network."""
        P = self.planes
        pk = ops.Packer(self.device)
        w: Dict[str, torch.Tensor] = {}
        self.w = w

        def conv3(k):
            w[k + ".w16"] = pk.weight16(sd[k + ".weight"], P)
            w[k + ".b"] = pk.f32(sd[k + ".bias"])

        def lin(k, bias=True):
            w[k + ".w16"] = pk.weight16(sd[k + ".weight"], P)
            if bias:
                w[k + ".b"] = pk.f32(sd[k + ".bias"])

        def norm(k):
            w[k + ".g"] = pk.f32(sd[k + ".weight"])
            w[k + ".be"] = pk.f32(sd[k + ".bias"])

        def stacked(keys):
            """several [n_i, K] linears stacked along N into one packed matrix (fused q|k|v / k|v projections)"""
            rows = sum(sd[k].shape[0] for k in keys)
            K = sd[keys[0]].shape[1]
            dst = torch.empty((P * rows, K), dtype=torch.float16, device=self.device)
            r0 = 0
            for k in keys:
                pk.weight16(sd[k], P, into=dst, row0=r0, total_rows=rows)
                r0 += sd[k].shape[0]
            return dst

        lay = self.layout
        res_keys, tr_keys, conv_keys = [], [], []
        self.attn_layers = []
        for i, blk in enumerate(lay["down"]):
            for j in range(blk["layers"]):
                res_keys.append(f"down_blocks.{i}.resnets.{j}")
                if blk["attn"]:
                    tr_keys.append(f"down_blocks.{i}.attentions.{j}")
                    self.attn_layers.append((tr_keys[-1], blk["cout"]))
            if blk["down"]:
                conv_keys.append(f"down_blocks.{i}.downsamplers.0.conv")
        res_keys.append("mid_block.resnets.0")
        tr_keys.append("mid_block.attentions.0")
        self.attn_layers.append(("mid_block.attentions.0", lay["mid"]["ch"]))
        res_keys.append("mid_block.resnets.1")
        for i, blk in enumerate(lay["up"]):
            for j in range(len(blk["skips"])):
                res_keys.append(f"up_blocks.{i}.resnets.{j}")
                if blk["attn"]:
                    tr_keys.append(f"up_blocks.{i}.attentions.{j}")
                    self.attn_layers.append((tr_keys[-1], blk["cout"]))
            if blk["up"]:
                conv_keys.append(f"up_blocks.{i}.upsamplers.0.conv")

        # the 22 time_emb_proj layers are stacked into one [sum(Cout), 1280] fp32 matrix (one GEMV per step)
        self.temb_off, off = {}, 0
        for k in res_keys:
            self.temb_off[k] = off
            off += sd[k + ".time_emb_proj.weight"].shape[0]
        self.temb_total = off
        tdim = sd[res_keys[0] + ".time_emb_proj.weight"].shape[1]
        w["temb_proj.w"] = torch.empty((off, tdim), dtype=torch.float32, device=self.device)
        w["temb_proj.b"] = torch.empty((off,), dtype=torch.float32, device=self.device)
        for k in res_keys:
            norm(k + ".norm1")
            conv3(k + ".conv1")
            norm(k + ".norm2")
            w[k + ".conv2.w16"] = pk.weight16(sd[k + ".conv2.weight"], P)
            if k + ".conv_shortcut.weight" in sd:  # fused 1x1 shortcut: second operand group, bias folded into conv2's
                w[k + ".sc.w16"] = pk.weight16(sd[k + ".conv_shortcut.weight"], P)
                w[k + ".conv2.b"] = pk.f32(sd[k + ".conv2.bias"], add=sd[k + ".conv_shortcut.bias"])
            else:
                w[k + ".conv2.b"] = pk.f32(sd[k + ".conv2.bias"])
            pk.f32(sd[k + ".time_emb_proj.weight"], into=w["temb_proj.w"], row0=self.temb_off[k])
            pk.f32(sd[k + ".time_emb_proj.bias"].reshape(-1, 1), into=w["temb_proj.b"], row0=self.temb_off[k])
        for k in tr_keys:
            norm(k + ".norm")
            lin(k + ".proj_in")
            lin(k + ".proj_out")
            b = k + ".transformer_blocks.0"
            for n in ("norm1", "norm2", "norm3"):
                norm(f"{b}.{n}")
            w[b + ".attn1.qkv.w16"] = stacked([f"{b}.attn1.to_{x}.weight" for x in "qkv"])
            lin(b + ".attn1.to_out.0")
            lin(b + ".attn2.to_q", bias=False)
            w[b + ".attn2.kv.w16"] = stacked([f"{b}.attn2.to_{x}.weight" for x in "kv"])
            lin(b + ".attn2.to_out.0")
            w[b + ".ff.net.0.proj.w16"] = pk.weight16(sd[b + ".ff.net.0.proj.weight"], P, geglu=True)
            w[b + ".ff.net.0.proj.b"] = pk.f32(sd[b + ".ff.net.0.proj.bias"].reshape(-1, 1), geglu=True).reshape(-1)
            lin(b + ".ff.net.2")
        for k in conv_keys:
            conv3(k)
        w["conv_in.w"] = pk.small_in(sd["conv_in.weight"])
        w["conv_in.b"] = pk.f32(sd["conv_in.bias"])
        for n in ("time_embedding.linear_1", "time_embedding.linear_2"):
            w[n + ".w"] = pk.f32(sd[n + ".weight"])
            w[n + ".b"] = pk.f32(sd[n + ".bias"])
        norm("conv_norm_out")
        w["conv_out.wp"] = pk.small_out(sd["conv_out.weight"])
        w["conv_out.b"] = pk.f32(sd["conv_out.bias"])
        pk.run()

    # ------------------------------------------------------------------------------------------
    # nn.Module surface the reference's training / checkpoint hooks touch (train_diffute_v1.py:664-687)
    # ------------------------------------------------------------------------------------------
    def state_dict(self) -> Dict[str, torch.Tensor]:
        """diffusers key -> fp32 tensor (the tensors this model was loaded from, in inventory order)."""
        return {k: self._sd[k] for k in arch.unet_param_shapes(self.config)}

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True):
        shapes = arch.unet_param_shapes(self.config)
        missing = [k for k in shapes if k not in state_dict]
        unexpected = [k for k in state_dict if k not in shapes]
        if missing or (strict and unexpected):
            raise KeyError(f"UNet state dict: {len(missing)} missing keys (e.g. {missing[:3]}), "
                           f"{len(unexpected)} unexpected (e.g. {unexpected[:3]})")
        for k, s in shapes.items():
            if tuple(state_dict[k].shape) != tuple(s):
                raise ValueError(f"{k}: expected shape {s}, got {tuple(state_dict[k].shape)}")
        self._sd = {k: state_dict[k].detach().to(torch.float32) for k in shapes}
        self._pack(self._sd)
        self._graphs.clear()          # captured steps hold the old packed-weight addresses
        self._weights_generation += 1
        self._ctx_key = None
        return self

    def parameters(self):
        return iter(self.state_dict().values())

    def named_parameters(self):
        return iter(self.state_dict().items())

    def register_to_config(self, **kwargs):
        """diffusers ConfigMixin.register_to_config (train_diffute_v1.py:687 copies a loaded model's config over).
        Architecture-defining entries may not change under a loaded model."""
        new = dict(self.config)
        new.update({k: v for k, v in kwargs.items() if not k.startswith("_")})
        if dict(arch.unet_param_shapes(new)) != dict(arch.unet_param_shapes(self.config)):
            raise ValueError("register_to_config: the new config describes a different architecture than the loaded "
                             "weights; build a new UNet2DConditionModel instead")
        self.config = _Config(new)

    def save_pretrained(self, save_directory: str, safe_serialization: bool = True, **kw):
        """Writes `<save_directory>/config.json` + `diffusion_pytorch_model.{safetensors|bin}` — the layout
        `from_pretrained(parent, subfolder=basename)` reads back (train_diffute_v1.py:669 saves to `<out>/unet`)."""
        from .checkpoint import save_diffusers_folder
        save_diffusers_folder(save_directory, None, dict(self.config), self.state_dict(), "UNet2DConditionModel",
                              safe_serialization=safe_serialization)

    # nn.Module-ish conveniences the reference scripts touch (train_diffute_v1.py:657, :696, :859)
    def eval(self):
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("diffute_b200 is a sampling engine: training kernels are out of scope (DESIGN.md)")
        return self

    def requires_grad_(self, flag: bool = False):
        return self

    def to(self, *a, **k):
        return self

    def cuda(self, *a):
        return self

    def enable_xformers_memory_efficient_attention(self, *a, **k):
        return None  # attention is always the fused tcgen05 kernel

    def enable_gradient_checkpointing(self):
        return None

    # ------------------------------------------------------------------------------------------
    # building blocks (each launches C-ABI kernels on the current stream)
    # ------------------------------------------------------------------------------------------
    def _op16(self, name, shape):
        return self.arena.get(name, (self.planes, *shape), torch.float16)

    def _resnet(self, k, x0, x1, cout, tproj, tag):
        w, A, prec = self.w, self.arena, self.prec
        B, H, W, c0 = x0.shape
        cin = c0 + (0 if x1 is None else x1.shape[-1])
        has_sc = (k + ".sc.w16") in w
        a16 = self._op16("opA", (B, H, W, cin))
        raw16 = self._op16("opRaw", (B, H, W, cin)) if has_sc else None
        ops.groupnorm(x0, w[k + ".norm1.g"], w[k + ".norm1.be"], self.config["norm_eps"], True, prec, src1=x1,
                      out16=a16, raw16=raw16, ws=self.ws)
        hmid = A.get("res_mid", (B, H, W, cout))
        off = self.temb_off[k]
        ops.conv(a16.view(self.planes * B, H, W, cin), w[k + ".conv1.w16"], cout, prec, (B, H, W), ops.taps_3x3_s1(),
                 ws=self.ws, out_f32=hmid.view(B * H * W, cout), bias=w[k + ".conv1.b"],
                 rowvec=tproj[:, off:off + cout], rows_per_sample=H * W)
        b16 = self._op16("opA", (B, H, W, cout))
        ops.groupnorm(hmid, w[k + ".norm2.g"], w[k + ".norm2.be"], self.config["norm_eps"], True, prec, out16=b16,
                      ws=self.ws)
        out = A.get(tag, (B, H, W, cout))
        if has_sc:
            ops.conv(b16.view(self.planes * B, H, W, cout), w[k + ".conv2.w16"], cout, prec, (B, H, W),
                     ops.taps_3x3_s1(), shortcut=(raw16.view(self.planes * B, H, W, cin), w[k + ".sc.w16"]),
                     ws=self.ws, out_f32=out.view(B * H * W, cout), bias=w[k + ".conv2.b"])
        else:
            ops.conv(b16.view(self.planes * B, H, W, cout), w[k + ".conv2.w16"], cout, prec, (B, H, W),
                     ops.taps_3x3_s1(), ws=self.ws, out_f32=out.view(B * H * W, cout), bias=w[k + ".conv2.b"],
                     residual=x0.view(B * H * W, cout))
        return out

    def _transformer(self, k, x, heads, ctx_kv, n_ctx, tag):
        w, A, prec, P = self.w, self.arena, self.prec, self.planes
        B, H, W, C = x.shape
        M = B * H * W
        N = H * W
        b = k + ".transformer_blocks.0"
        g16 = self._op16("opA", (M, C))
        ops.groupnorm(x, w[k + ".norm.g"], w[k + ".norm.be"], 1e-6, False, prec, out16=g16.view(P, B, H, W, C),
                      ws=self.ws)
        t0 = A.get("tok0", (M, C))
        ops.linear(g16, w[k + ".proj_in.w16"], C, prec, ws=self.ws, out_f32=t0, bias=w[k + ".proj_in.b"])
        # self-attention
        ln = self._op16("ln", (M, C))
        ops.layernorm(t0, w[b + ".norm1.g"], w[b + ".norm1.be"], 1e-5, ln)
        qkv = self._op16("qkv", (M, 3 * C))
        ops.linear(ln, w[b + ".attn1.qkv.w16"], 3 * C, prec, ws=self.ws, out_f16=qkv)
        at = self._op16("attn", (M, C))
        ops.attention(qkv, 0, qkv, C, qkv, 2 * C, B, heads, N, N, 0.125, at, ws=self.ws)
        t1 = A.get("tok1", (M, C))
        ops.linear(at, w[b + ".attn1.to_out.0.w16"], C, prec, ws=self.ws, out_f32=t1, bias=w[b + ".attn1.to_out.0.b"],
                   residual=t0)
        # cross-attention over the glyph tokens (K/V projected once per encoder_hidden_states)
        ops.layernorm(t1, w[b + ".norm2.g"], w[b + ".norm2.be"], 1e-5, ln)
        q = self._op16("qkv", (M, C))
        ops.linear(ln, w[b + ".attn2.to_q.w16"], C, prec, ws=self.ws, out_f16=q)
        ops.attention(q, 0, ctx_kv, 0, ctx_kv, C, B, heads, N, n_ctx, 0.125, at, ws=self.ws)
        t2 = A.get("tok0", (M, C))
        ops.linear(at, w[b + ".attn2.to_out.0.w16"], C, prec, ws=self.ws, out_f32=t2, bias=w[b + ".attn2.to_out.0.b"],
                   residual=t1)
        # GEGLU feed-forward; its second linear writes the fp16 operand of proj_out directly
        ops.layernorm(t2, w[b + ".norm3.g"], w[b + ".norm3.be"], 1e-5, ln)
        gg = self._op16("ffh", (M, 4 * C))
        ops.linear(ln, w[b + ".ff.net.0.proj.w16"], 8 * C, prec, ws=self.ws, out_f16=gg, bias=w[b + ".ff.net.0.proj.b"],
                   geglu=True)
        t3 = self._op16("ln", (M, C))
        ops.linear(gg, w[b + ".ff.net.2.w16"], C, prec, ws=self.ws, out_f16=t3, bias=w[b + ".ff.net.2.b"], residual=t2)
        out = A.get(tag, (B, H, W, C))
        ops.linear(t3, w[k + ".proj_out.w16"], C, prec, ws=self.ws, out_f32=out.view(M, C), bias=w[k + ".proj_out.b"],
                   residual=x.view(M, C))
        return out

    def _downsample(self, k, x, tag):
        B, H, W, C = x.shape
        s16 = self._op16("opA", (4 * B, H // 2, W // 2, C))
        ops.cast_f16(x, ops.CAST_S2D, s16)
        out = self.arena.get(tag, (B, H // 2, W // 2, C))
        ops.conv(s16.view(self.planes * 4 * B, H // 2, W // 2, C), self.w[k + ".w16"], C, self.prec,
                 (B, H // 2, W // 2), ops.taps_3x3_s2(B, self.config["downsample_padding"]), imgs_per_plane=4 * B,
                 ws=self.ws, out_f32=out.view(-1, C), bias=self.w[k + ".b"])
        return out

    def _upsample(self, k, x, tag):
        B, H, W, C = x.shape
        u16 = self._op16("opA", (B, 2 * H, 2 * W, C))
        ops.cast_f16(x, ops.CAST_UP2X, u16)
        out = self.arena.get(tag, (B, 2 * H, 2 * W, C))
        ops.conv(u16.view(self.planes * B, 2 * H, 2 * W, C), self.w[k + ".w16"], C, self.prec, (B, 2 * H, 2 * W),
                 ops.taps_3x3_s1(), ws=self.ws, out_f32=out.view(-1, C), bias=self.w[k + ".b"])
        return out

    # ------------------------------------------------------------------------------------------
    # glyph-token K/V projections: step-invariant, hoisted out of the sampling loop
    # ------------------------------------------------------------------------------------------
    def prepare_context(self, ehs: torch.Tensor):
        """ehs [B, T, cross_attention_dim] fp32 -> per-layer fp16 K|V operands [planes, B*T, 2C] (static buffers)."""
        B, T, D = ehs.shape
        e = self.arena.get("ehs", (B, T, D))
        e.copy_(ehs)
        e16 = self._op16("ehs16", (B * T, D))
        ops.cast_f16(e.view(1, 1, B * T, D), ops.CAST_PLAIN, e16.view(self.planes, 1, 1, B * T, D))
        self.ctx = {}
        for k, C in self.attn_layers:
            kv = self._op16("ctxkv." + k, (B * T, 2 * C))
            ops.linear(e16, self.w[k + ".transformer_blocks.0.attn2.kv.w16"], 2 * C, self.prec, ws=self.ws, out_f16=kv)
            self.ctx[k] = kv
        self.n_ctx = T
        self.ctx_batch = B

    def _ensure_context(self, ehs: torch.Tensor):
        """Re-project K/V only when `encoder_hidden_states` is a different tensor object or was modified in place.
        The tensor is kept referenced so its storage cannot be recycled under the same identity."""
        key = self._ctx_key
        if key is None or key[0] is not ehs or key[1] != ehs._version:
            self.prepare_context(ehs.to(device=self.device, dtype=torch.float32))
            self._ctx_key = (ehs, ehs._version)

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    def time_projections(self, t: torch.Tensor, out: torch.Tensor):
        """All time-dependent inputs of one step, tproj [B, sum(Cout)] = time_emb_proj_j(SiLU(time_embedding(t))) for
        the 22 resnets.  They depend on t only, so a sampler may compute them once per timestep and reuse them."""
        A, w, cfg = self.arena, self.w, self.config
        B = t.shape[0]
        c0 = cfg["block_out_channels"][0]
        te = A.get("t.sincos", (B, c0))
        ops.timestep_embedding(t, c0, cfg["flip_sin_to_cos"], cfg["freq_shift"], te)
        t1 = A.get("t.h1", (B, 4 * c0))
        ops.gemv(te, w["time_embedding.linear_1.w"], w["time_embedding.linear_1.b"], t1, silu_out=True)
        temb = A.get("t.emb", (B, 4 * c0))
        ops.gemv(t1, w["time_embedding.linear_2.w"], w["time_embedding.linear_2.b"], temb)
        ops.gemv(temb, w["temb_proj.w"], w["temb_proj.b"], out, silu_in=True)

    def _forward_impl(self, B, H, W, step_io=None, srcs=None, t=None, tproj=None):
        """All launches of one denoising step against static buffers: `in.sample` [B,9,H,W] (or `srcs`, up to three
        NCHW tensors whose channel concat is the UNet input — the cat of app.ipynb:811 is then never materialised),
        `in.t` [B] and the prepared glyph context."""
        A, w, cfg, lay = self.arena, self.w, self.config, self.layout
        if srcs is None:
            srcs = [A.get("in.sample", (B, cfg["in_channels"], H, W))]
        if t is None:
            t = A.get("in.t", (B,))
        c0 = cfg["block_out_channels"][0]
        if tproj is None:
            # time embedding: sincos -> linear/SiLU -> linear, then all 22 time_emb_proj(SiLU(temb)) in one launch
            tproj = A.get("t.proj", (B, self.temb_total))
            self.time_projections(t, tproj)

        h = A.get("h.in", (B, H, W, c0))
        ops.conv_small_in(srcs, w["conv_in.w"], w["conv_in.b"], h, B)
        skips = [h]
        n = 0

        def tag():
            nonlocal n
            n += 1
            return f"h.{n}"

        for i, blk in enumerate(lay["down"]):
            for j in range(blk["layers"]):
                h = self._resnet(f"down_blocks.{i}.resnets.{j}", h, None, blk["cout"], tproj, tag())
                if blk["attn"]:
                    k = f"down_blocks.{i}.attentions.{j}"
                    h = self._transformer(k, h, blk["heads"], self.ctx[k], self.n_ctx, tag())
                skips.append(h)
            if blk["down"]:
                h = self._downsample(f"down_blocks.{i}.downsamplers.0.conv", h, tag())
                skips.append(h)
        mid = lay["mid"]
        h = self._resnet("mid_block.resnets.0", h, None, mid["ch"], tproj, tag())
        h = self._transformer("mid_block.attentions.0", h, mid["heads"], self.ctx["mid_block.attentions.0"], self.n_ctx,
                              tag())
        h = self._resnet("mid_block.resnets.1", h, None, mid["ch"], tproj, tag())
        for i, blk in enumerate(lay["up"]):
            for j in range(len(blk["skips"])):
                s = skips.pop()
                h = self._resnet(f"up_blocks.{i}.resnets.{j}", h, s, blk["cout"], tproj, tag())
                if blk["attn"]:
                    k = f"up_blocks.{i}.attentions.{j}"
                    h = self._transformer(k, h, blk["heads"], self.ctx[k], self.n_ctx, tag())
            if blk["up"]:
                h = self._upsample(f"up_blocks.{i}.upsamplers.0.conv", h, tag())
        Bh, Hh, Wh, Ch = h.shape
        o32 = A.get("out.norm", (Bh, Hh, Wh, Ch))
        ops.groupnorm(h, w["conv_norm_out.g"], w["conv_norm_out.be"], cfg["norm_eps"], True, self.prec, out32=o32,
                      ws=self.ws)
        out = A.get("out.eps", (B, cfg["out_channels"], H, W))
        if step_io is None:
            ops.conv_small_out(o32, w["conv_out.wp"], w["conv_out.b"], out)
        else:  # fused scheduler update: prev = coef[0]*latents + coef[1]*eps written by the same kernel
            lat, prev, coef = step_io[:3]  # (+ seed words: the ancestral DDPM noise is added by the same kernel)
            ops.conv_small_out(o32, w["conv_out.wp"], w["conv_out.b"], out, sample=lat, prev=prev, coef=coef,
                               seed=step_io[3] if len(step_io) > 3 else None)
        return out

    def _run(self, B, H, W):
        # the glyph-context length and batch are kernel arguments (Nk, K/V row strides) baked into a captured graph
        key = (B, H, W, self.n_ctx, self.ctx_batch)
        if not self.use_cuda_graph:
            return self._forward_impl(B, H, W)
        g = self._graphs.get(key)
        if g is None or g[2] != self.buffer_generation():
            # warm-up run allocates every static buffer and sets function attributes; then capture
            self._forward_impl(B, H, W)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._forward_impl(B, H, W)
            g = (graph, out, self.buffer_generation())
            self._graphs[key] = g
        g[0].replay()
        return g[1]

    def buffer_generation(self):
        return (self.arena.generation, self.ws.generation, self._weights_generation)

    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, timestep_cond=None,
                attention_mask=None, cross_attention_kwargs=None, down_block_additional_residuals=None,
                mid_block_additional_residual=None, return_dict: bool = True):
        if any(x is not None for x in (class_labels, timestep_cond, attention_mask, down_block_additional_residuals,
                                       mid_block_additional_residual)):
            raise NotImplementedError("class_labels / timestep_cond / attention_mask / ControlNet residuals are not "
                                      "used by DiffUTE (app.ipynb:814) and are not implemented")
        if sample.dim() != 4 or sample.shape[1] != self.config["in_channels"]:
            raise ValueError(f"sample must be [B,{self.config['in_channels']},H,W], got {tuple(sample.shape)}")
        B, _, H, W = sample.shape
        if H % 8 or W % 8:
            raise ValueError("latent height/width must be multiples of 8 (three stride-2 stages)")
        if encoder_hidden_states.shape[0] != B or encoder_hidden_states.shape[2] != self.config["cross_attention_dim"]:
            raise ValueError(f"encoder_hidden_states must be [{B},T,{self.config['cross_attention_dim']}]")
        dev = self.device
        A = self.arena
        A.get("in.sample", (B, self.config["in_channels"], H, W)).copy_(sample.to(device=dev, dtype=torch.float32))
        tbuf = A.get("in.t", (B,))
        if torch.is_tensor(timestep):
            tt = timestep.to(device=dev, dtype=torch.float32).reshape(-1)
            tbuf.copy_(tt.expand(B) if tt.numel() == 1 else tt)
        else:
            tbuf.fill_(float(timestep))
        self._ensure_context(encoder_hidden_states)
        out = self._run(B, H, W).clone()
        return UNet2DConditionOutput(out) if return_dict else (out,)

    __call__ = forward

    def state_dict_keys(self):
        return list(arch.unet_param_shapes(self.config).keys())
