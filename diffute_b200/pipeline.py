"""DiffUTEPipeline — the reference's `text_editing` sampling core (app.ipynb:772-819) as one call.

The reference has no pipeline class; its loop is: TrOCR-encode the glyph image, VAE-encode the masked image,
draw seeded noise, N x (cat -> unet -> scheduler.step), VAE-decode.  This class runs exactly that sequence on the
native kernels, with the invariants hoisted: glyph K/V projections once per request, mask and masked-image latents
gathered by conv_in instead of being concatenated each step, the scheduler update fused into conv_out, and the
whole UNet step replayed as one CUDA graph.  The signature follows diffusers' StableDiffusionInpaintPipeline with
`prompt_embeds` replaced by `glyph_embeds`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import ops
from .schedulers import DDIMScheduler, DDPMScheduler


@dataclass
class DiffUTEPipelineOutput:
    images: object
    latents: Optional[torch.Tensor] = None

    def __getitem__(self, k):
        if k in (0, "images"):
            return self.images
        raise KeyError(k)


class DiffUTEPipeline:
    def __init__(self, vae, unet, scheduler, glyph_encoder=None, glyph_processor=None):
        self.vae, self.unet, self.scheduler = vae, unet, scheduler
        self.glyph_encoder, self.glyph_processor = glyph_encoder, glyph_processor
        self.device = unet.device
        self.fuse_scheduler_step = True  # False: always run scheduler.step() as its own kernel (tests, custom schedulers)
        self._graphs = {}
        self._rows_cache = {}

    @classmethod
    def from_pretrained(cls, path, unet_precision="fp16", vae_precision="fp16x2", scheduler_cls=DDIMScheduler,
                        vae_encoder_precision=None, **kw):
        """Folder layout of the reference: <path>/{unet,vae,scheduler}/ (app.ipynb:545-553)."""
        from .unet import UNet2DConditionModel
        from .vae import AutoencoderKL
        unet = UNet2DConditionModel.from_pretrained(path, "unet", precision=unet_precision)
        vae = AutoencoderKL.from_pretrained(path, "vae", precision=vae_precision,
                                            encoder_precision=vae_encoder_precision)
        return cls(vae, unet, scheduler_cls.from_pretrained(path, "scheduler"), **kw)

    @classmethod
    def from_synthetic(cls, unet_precision="fp16", vae_precision="fp16x2", seed: int = 1234, state_dicts=None,
                       vae_encoder_precision=None):
        from . import arch, synthetic
        from .unet import UNet2DConditionModel
        from .vae import AutoencoderKL
        if state_dicts is None:
            state_dicts = (synthetic.make_state_dict(arch.unet_param_shapes(), seed),
                           synthetic.make_state_dict(arch.vae_param_shapes(), seed))
        unet = UNet2DConditionModel(state_dicts[0], precision=unet_precision)
        vae = AutoencoderKL(state_dicts[1], precision=vae_precision, encoder_precision=vae_encoder_precision)
        return cls(vae, unet, DDIMScheduler())

    # ------------------------------------------------------------------------------------------
    def encode_glyph(self, glyph_images):
        """TrOCR encoder last_hidden_state [B,577,1024] (app.ipynb:773-776) through the attached `glyph_processor` /
        `glyph_encoder`: glyph_encoder.TrOCRGlyphProcessor + TrOCRGlyphEncoder (native kernels), or transformers' own
        objects -- both have the same call signatures.  Once per request, ~1% of the FLOPs."""
        if self.glyph_encoder is None or self.glyph_processor is None:
            raise ValueError("no glyph encoder attached: pass glyph_embeds [B,577,1024] instead of text/glyph images")
        pv = self.glyph_processor(images=glyph_images, return_tensors="pt").pixel_values.to(self.device)
        with torch.no_grad():
            return self.glyph_encoder(pv).last_hidden_state.detach().float()

    @staticmethod
    def _tproj_offset(UB: int) -> int:
        """start of the time projections inside a step row, 16-byte aligned (the epilogues read them as float4)"""
        return (UB + 4 + 3) // 4 * 4

    def _step_rows(self, ts, UB, fused, sched):
        """Per-step device rows [t x UB | cx, ce | time projections]: everything one step needs that depends on t only.
        The 22 time_emb_proj outputs are a function of the timestep (not of the image), so they are computed once per
        (timestep list, batch) and reused by every later request; each step then costs one small D2D copy."""
        key = (tuple(ts), UB, fused, type(sched).__name__, tuple(sorted(sched.config.items(), key=str)).__hash__())
        rows = self._rows_cache.get(key)
        if rows is not None:
            return rows
        dev = self.device
        TT = self.unet.temb_total
        off = self._tproj_offset(UB)
        rows = torch.zeros((len(ts), off + TT * UB), dtype=torch.float32, device=dev)
        tv = torch.empty((UB,), dtype=torch.float32, device=dev)
        for i, t in enumerate(ts):
            tv.fill_(float(t))
            rows[i, :UB] = float(t)
            if fused:
                co = sched.collapsed_coefficients(t)   # DDIM: (cx, ce); DDPM: (cx, ce, sigma)
                rows[i, UB], rows[i, UB + 1] = co[0], co[1]
                if len(co) > 2:                          # ancestral: sigma and the step index (uint32 bits) of the noise stream
                    rows[i, UB + 2] = co[2]
                    rows[i, UB + 3:UB + 4].view(torch.int32)[0] = i
            self.unet.time_projections(tv, rows[i, off:].view(UB, TT))
        self._rows_cache[key] = rows
        return rows

    def _step_graph(self, B, h, w, fused: bool):
        """Capture (once per shape) one denoising step: UNet + fused DDIM update, reading t / coefficients from the
        device `state` row that the loop refreshes with one small copy per step."""
        # the glyph-context length is a kernel argument (Nk, K/V strides) baked into the captured step
        key = (B, h, w, fused, self.unet.n_ctx, self.unet.ctx_batch)
        g = self._graphs.get(key)
        if g is not None and g[2] == self.unet.buffer_generation():
            return g
        A = self.unet.arena
        lat = A.get("pipe.latents", (B, 4, h, w))
        mask = A.get("pipe.mask", (B, 1, h, w))
        ml = A.get("pipe.masked", (B, 4, h, w))
        off = self._tproj_offset(B)
        state = A.get("pipe.state", (off + self.unet.temb_total * B,))
        step_io = None
        if fused == 1:
            step_io = (lat, lat, state[B:B + 2])
        elif fused == 2:  # DDPM: + the per-request noise seed (two 32-bit words, written once per call)
            step_io = (lat, lat, state[B:B + 4], A.get("pipe.seed", (2,), torch.int32))
        tproj = state[off:].view(B, self.unet.temb_total)

        def run():
            return self.unet._forward_impl(B, h, w, step_io=step_io, srcs=[lat, mask, ml], t=state[:B], tproj=tproj)

        run()  # warm-up: allocates static buffers
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            eps = run()
        g = (graph, eps, self.unet.buffer_generation())
        self._graphs[key] = g
        return g

    @torch.no_grad()
    def __call__(self, image: Optional[torch.Tensor] = None, mask_image: Optional[torch.Tensor] = None,
                 glyph_embeds: Optional[torch.Tensor] = None, text=None, masked_image: Optional[torch.Tensor] = None,
                 height: Optional[int] = None, width: Optional[int] = None, num_inference_steps: int = 50,
                 guidance_scale: float = 1.0, negative_glyph_embeds: Optional[torch.Tensor] = None, eta: float = 0.0,
                 generator: Optional[torch.Generator] = None, latents: Optional[torch.Tensor] = None,
                 posterior_noise: Optional[torch.Tensor] = None, sample_posterior: bool = True,
                 output_type: str = "pt", return_dict: bool = True, noise_seed: Optional[int] = None):
        """image / masked_image [B,3,H,W] in [-1,1]; mask_image [B,1,H,W] (1 = region to rewrite);
        glyph_embeds [B,577,1024] (or `text` = glyph images for the attached TrOCR encoder).  Returns decoded RGB in
        [-1,1] for output_type "pt" (what vae.decode returns at app.ipynb:819), [0,1] HWC numpy for "np", PIL for
        "pil", the final latents for "latent"."""
        if (glyph_embeds is None) == (text is None):
            raise ValueError("pass exactly one of glyph_embeds / text")
        if glyph_embeds is None:
            glyph_embeds = self.encode_glyph(text)
        if masked_image is None:
            if image is None or mask_image is None:
                raise ValueError("pass masked_image, or image and mask_image")
            # the reference zeroes the masked pixels of the uint8 image before Normalize(0.5,0.5): they become -1
            masked_image = torch.where(mask_image > 0.5, torch.full_like(image, -1.0), image)
        if mask_image is None:
            raise ValueError("mask_image is required")
        dev = self.device
        B, _, H, W = masked_image.shape
        if (height and height != H) or (width and width != W):
            raise ValueError("height/width must match the image tensors (resize is reference-side glue)")
        vsf = 2 ** (len(self.vae.config["block_out_channels"]) - 1)
        if H % (8 * vsf) or W % (8 * vsf):
            raise ValueError(f"height and width must be multiples of {8 * vsf}")
        h, w = H // vsf, W // vsf
        sf = float(self.vae.config["scaling_factor"])
        do_cfg = guidance_scale != 1.0 and negative_glyph_embeds is not None
        sched = self.scheduler
        # 1: DDIM update in conv_out's epilogue; 2: the ancestral DDPM update the reference runs (app.ipynb:545, :816) with
        # its Gaussian noise generated in the same epilogue -- unless the caller supplies a torch generator, whose stream
        # only torch can reproduce (then the step runs as its own kernel on torch.randn noise)
        fused = 0
        if self.fuse_scheduler_step and not sched.config["clip_sample"] and not do_cfg:
            if isinstance(sched, DDIMScheduler) and eta == 0.0:
                fused = 1
            elif isinstance(sched, DDPMScheduler) and (generator is None or noise_seed is not None):
                fused = 2
        UB = 2 * B if do_cfg else B

        # --- step-invariant work (once per request) -------------------------------------------------
        post = self.vae.encode(masked_image.to(dev)).latent_dist                      # app.ipynb:793
        if posterior_noise is not None:
            ml = post.sample(noise=posterior_noise, scale=sf)
        elif sample_posterior:
            ml = post.sample(generator=generator, scale=sf)
        else:
            ml = post.mode(scale=sf)
        mask_l = mask_image.to(dev, torch.float32)[:, :, ::vsf, ::vsf].contiguous()   # nearest, app.ipynb:787-790
        if latents is None:                                                            # app.ipynb:796-801
            gdev = generator.device if generator is not None else "cpu"
            latents = torch.randn((B, 4, h, w), generator=generator, device=gdev, dtype=torch.float32)
        latents = latents.to(dev, torch.float32) * float(sched.init_noise_sigma)
        ctx = glyph_embeds.to(dev, torch.float32)
        if do_cfg:
            ctx = torch.cat([negative_glyph_embeds.to(dev, torch.float32), ctx], 0)
        self.unet.prepare_context(ctx)
        self.unet._ctx_key = None

        A = self.unet.arena
        graph, eps, _ = self._step_graph(UB, h, w, fused)
        lat_buf = A.get("pipe.latents", (UB, 4, h, w))
        A.get("pipe.mask", (UB, 1, h, w)).copy_(mask_l.repeat(2, 1, 1, 1) if do_cfg else mask_l)
        A.get("pipe.masked", (UB, 4, h, w)).copy_(ml.repeat(2, 1, 1, 1) if do_cfg else ml)
        TT = self.unet.temb_total
        state = A.get("pipe.state", (self._tproj_offset(UB) + TT * UB,))
        sched.set_timesteps(num_inference_steps)                                       # app.ipynb:803
        ts = [int(t) for t in sched.timesteps]
        rows = self._step_rows(ts, UB, fused, sched)

        # --- the loop (app.ipynb:806-816) -------------------------------------------------------------
        if fused:
            lat_buf.copy_(latents)
            if fused == 2:
                if noise_seed is None:  # from torch's default generator: torch.manual_seed makes the run reproducible
                    noise_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
                self.last_noise_seed = int(noise_seed)
                words = [noise_seed & 0xFFFFFFFF, (noise_seed >> 32) & 0xFFFFFFFF]
                A.get("pipe.seed", (2,), torch.int32).copy_(
                    torch.tensor([w - (1 << 32) if w >= (1 << 31) else w for w in words], dtype=torch.int32))
            for i in range(len(ts)):
                state.copy_(rows[i])       # one small D2D copy: timestep + DDIM coefficients of this step
                graph.replay()             # UNet + scheduler update, latents advanced in place
            latents = lat_buf.clone()  # the arena buffer is overwritten by the next call: hand out a copy
        else:
            eps_g = torch.empty((B, 4, h, w), device=dev) if do_cfg else None
            for i, t in enumerate(ts):
                lat_buf.copy_(torch.cat([latents, latents], 0) if do_cfg else latents)
                state.copy_(rows[i])
                graph.replay()
                e = eps
                if do_cfg:  # eps = eps_u + g (eps_c - eps_u)
                    ops.axpbypcz(eps[:B], eps[B:], None, 1.0 - guidance_scale, guidance_scale, 0.0, eps_g)
                    e = eps_g
                if isinstance(sched, DDIMScheduler):
                    latents = sched.step(e, t, latents, eta=eta, generator=generator, return_dict=False)[0]
                else:
                    latents = sched.step(e, t, latents, generator=generator, return_dict=False)[0]

        if output_type == "latent":
            out = latents.clone()
            return DiffUTEPipelineOutput(out, out) if return_dict else (out,)
        img = self.vae.decode(latents, pre_scale=1.0 / sf).sample                      # app.ipynb:818-819
        if output_type in ("np", "pil"):
            arr = (img / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).cpu().numpy()       # app.ipynb:822-824
            if output_type == "pil":
                from PIL import Image
                arr = [Image.fromarray((a * 255).round().astype("uint8")) for a in arr]
            img = arr
        return DiffUTEPipelineOutput(img, latents) if return_dict else (img,)
