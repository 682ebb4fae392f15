#!/usr/bin/env python
"""bench.py — DiffUTE sampling throughput on B200 (BASELINE.json metric: images/sec at 512x512, 50 DDIM steps).

    python bench.py --gpus N --steps K --warmup W [--precision mixed|fp16x2|fp16] [--impl reference]

A "step" is one full pass of the hot path over one batch: VAE-encode the masked image, 50 x (UNet + DDIM update),
VAE-decode — one 512x512 image per GPU (BASELINE config 2; weak scaling over GPUs = config 3).  Rank 0 prints ONE
JSON line.  `value` times device-resident inputs; `e2e` times DiffUTEPipeline.__call__ with pinned HOST inputs and
a host read-back of the decoded image.  `--impl reference` times the CPU restatement of the reference path
(oracle/: the reference's diffusers is not installable offline) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PX = 512
NSTEPS = 50
# algorithmic work (FLOP = 2*MAC over conv / linear / attention matmuls)
UNET_GFLOP = 853.04           # one UNet forward @64x64 latent, incl. 29.50 of glyph K/V projection
UNET_CTX_GFLOP = 29.50        # step-invariant, executed once per image by this engine
VAE_ENC_GFLOP = 1116.66
VAE_DEC_GFLOP = 2514.52
IMAGE_GFLOP = NSTEPS * (UNET_GFLOP - UNET_CTX_GFLOP) + UNET_CTX_GFLOP + VAE_ENC_GFLOP + VAE_DEC_GFLOP


def _cpu_threads() -> int:
    """Threads for the CPU arm: the oracle's torch/oneDNN kernels stop scaling (and regress badly) far below the 128
    hardware threads of the GPU box's host — 64 s per UNet step with 128 threads vs ~3 s with 8-32 — so cap at 32."""
    return max(1, min(os.cpu_count() or 1, int(os.environ.get("DFU_CPU_THREADS", "32"))))


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                    hbm=float(d.get("hbm_gbs", 6650.0)), src="measured (MEASURED_PEAKS.json, sustained)")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# =================================================================================================
# CPU arm: the restated reference path on the host cores
# =================================================================================================
class CpuArm:
    """The restated reference path (oracle/) on the host cores: BASELINE config 2 cut into bounded pieces.

    One CPU "step" is ONE UNet forward at the 64x64 latent (1/50 of an image's loop); the VAE encode + decode at
    512x512 is timed once.  images/s is then 1 / (50 x median(UNet step) + VAE) — an extrapolation, reported as such;
    `ms_per_step` of the reference line is what one timed step really took.  Both the in-process `cpu_baseline` leg and
    `--impl reference` go through this class (same threads, same warm-up, same median) so that they agree."""

    def __init__(self, threads=None):
        import torch
        from diffute_b200 import arch, synthetic
        from oracle import UNetOracle, VAEOracle
        self.torch = torch
        self.threads = threads or _cpu_threads()
        torch.set_num_threads(self.threads)
        self.u, self.v = UNetOracle(), VAEOracle()
        self.u.load_state_dict(synthetic.make_state_dict(arch.unet_param_shapes()))
        self.v.load_state_dict(synthetic.make_state_dict(arch.vae_param_shapes()))
        self.inp = synthetic.make_inputs(1, PX, PX)
        self.x = torch.cat([self.inp["latents"], self.inp["mask"][:, :, ::8, ::8], self.inp["latents"]], 1)

    def unet_step(self) -> float:
        t0 = time.perf_counter()
        with self.torch.no_grad():
            self.u(self.x, 981, self.inp["glyph_embeds"])
        return time.perf_counter() - t0

    def vae(self):
        with self.torch.no_grad():
            t0 = time.perf_counter()
            z = self.v.encode(self.inp["masked_image"]).latent_dist.mode()
            t1 = time.perf_counter()
            self.v.decode(z)
            t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    @staticmethod
    def images_per_s(unet_step_s, vae_s):
        return 1.0 / (NSTEPS * unet_step_s + vae_s)


def cpu_reference_sample(threads=None, unet_reps=3, warmup=1):
    """Bounded sample for the `cpu_baseline` key: `warmup` + `unet_reps` UNet steps and one VAE encode + decode."""
    arm = CpuArm(threads)
    for _ in range(warmup):
        arm.unet_step()
    t_u = [arm.unet_step() for _ in range(unet_reps)]
    t_enc, t_dec = arm.vae()
    tu = statistics.median(t_u)
    return arm.images_per_s(tu, t_enc + t_dec), dict(unet_step_s=tu, vae_encode_s=t_enc, vae_decode_s=t_dec,
                                                       image_s_extrapolated=NSTEPS * tu + t_enc + t_dec,
                                                       cores=arm.threads, unet_reps=unet_reps)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    arm = CpuArm()
    t_enc, t_dec = arm.vae()               # once, outside the timed steps (it is 1/50th as frequent as a UNet step)
    for _ in range(max(args.warmup, 1)):
        arm.unet_step()
    ts = [arm.unet_step() for _ in range(args.steps)]
    tu = statistics.median(ts)
    val = arm.images_per_s(tu, t_enc + t_dec)
    sample = (f"{args.steps} timed steps, each ONE UNet forward (64x64 latent, B=1) on the fp32 CPU oracle (median "
              f"{tu:.3f} s); VAE encode {t_enc:.2f} s + decode {t_dec:.2f} s at 512x512 measured once; images/s = "
              f"1 / ({NSTEPS} x median UNet step + VAE), i.e. extrapolated from 1/{NSTEPS} of the loop per step")
    line = {"impl": "reference", "metric": "images/sec at 512x512, 50 DDIM steps", "value": val, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1),
            "ms_per_step": sum(ts) / len(ts) * 1e3,
            "ms_per_step_is": "one timed CPU step = one UNet forward (1/50 of an image's loop), as measured",
            "image_ms_extrapolated": (NSTEPS * tu + t_enc + t_dec) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "512x512 glyph-conditioned inpaint, 50 DDIM steps, batch 1 (BASELINE config 2)",
                       "arm": "CPU restatement of the reference's diffusers path (oracle/), torch fp32"},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": arm.threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# =================================================================================================
# GPU arm
# =================================================================================================
class _AblatingLib:
    """Timing ablation (this file only): a proxy over the CDLL whose ablated entry points are no-ops, installed by
    swapping `ops.lib` while a step is re-captured.  Results of an ablated step are garbage; only its duration is used."""
    CLASSES = {"dfu_gemm": "gemm", "dfu_attention": "attention", "dfu_groupnorm": "groupnorm",
               "dfu_layernorm": "layernorm"}

    def __init__(self, real, ablate):
        self._real, self._ablate = real, set(ablate)

    def __getattr__(self, name):
        fn = getattr(self._real, name)
        if self.CLASSES.get(name) in self._ablate:
            return lambda *a: 0
        return fn


def _event_ms(fn, n):
    import torch
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def _median_call_ms(fn, warmup=2, n=3):
    """Seconds-long calls (a whole batch of images): `warmup` untimed calls (the first replays of a freshly captured
    graph and first-touch allocations are slower), then the median of `n` individually timed calls."""
    for _ in range(warmup):
        fn()
    return statistics.median(_event_ms(fn, 1) for _ in range(n))


def in_graph_class_ms(pipe, B, h, w, classes=("gemm", "attention", "groupnorm", "layernorm"), reps=20):
    """In-graph cost of each kernel class of one UNet step by ablation: the step is re-captured with that class's
    launches removed (same buffers, same order, PDL edges intact) and replayed; cost = full - ablated (CUDA events on the
    replay stream, no per-launch event overhead).  Returns (full_ms, {class: ms})."""
    import torch
    from diffute_b200 import ops
    A = pipe.unet.arena
    srcs = [A.get("pipe.latents", (B, 4, h, w)), A.get("pipe.mask", (B, 1, h, w)), A.get("pipe.masked", (B, 4, h, w))]
    t = A.get("pipe.state", (B + 2,))[:B]
    real_lib = ops.lib

    def graph_ms(ablate):
        if ablate:
            ops.lib = lambda: _AblatingLib(real_lib(), ablate)
        try:
            pipe.unet._forward_impl(B, h, w, srcs=srcs, t=t)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                pipe.unet._forward_impl(B, h, w, srcs=srcs, t=t)
        finally:
            ops.lib = real_lib
        gr.replay()
        return _event_ms(gr.replay, reps)

    full = graph_ms(())
    return full, {c: max(full - graph_ms((c,)), 0.0) for c in classes}


def extra_configs(pipe, dev, peaks):
    """BASELINE configs 3 (1-GPU leg: batch 8), 4 (VAE encode+decode, batch 32) and 5 (768x768, batch 4, CFG x2), each
    with its own roofline sub-record.  Inputs are resident in HBM; CUDA events on the launch stream; warm-up calls, then
    the median of three individually timed calls (each call is 0.3-2 s of GPU work — far beyond L2, no flush needed)."""
    import torch
    from diffute_b200 import synthetic
    out = {}
    tf = peaks["tflops"]

    def dev_inputs(B, px, seed=0):
        inp = synthetic.make_inputs(B, px, px, seed=seed)
        return {k: v.to(dev) for k, v in inp.items()}

    # ---- config 3, single-GPU leg: 512x512, 50 steps, batch 8 ------------------------------------------------
    B = 8
    d = dev_inputs(B, PX)
    call = lambda: pipe(masked_image=d["masked_image"], mask_image=d["mask"], glyph_embeds=d["glyph_embeds"],
                        latents=d["latents"], posterior_noise=d["posterior_noise"], num_inference_steps=NSTEPS).images
    ms = _median_call_ms(call)
    graph, _, _ = pipe._step_graph(B, PX // 8, PX // 8, True)
    graph.replay()
    step_ms = _event_ms(graph.replay, 10)
    step_tf = (UNET_GFLOP - UNET_CTX_GFLOP) * B / 1e3 / (step_ms / 1e3)
    _, cls8 = in_graph_class_ms(pipe, B, PX // 8, PX // 8, reps=5)
    gemm_gf = (UNET_GFLOP - UNET_CTX_GFLOP - 149.15) * B          # conv + linear FLOP of a step (attention cores excluded)
    gemm_tf = gemm_gf / 1e3 / (cls8["gemm"] / 1e3) if cls8["gemm"] > 0 else 0.0
    attn_tf = 149.15 * B / 1e3 / (cls8["attention"] / 1e3) if cls8["attention"] > 0 else 0.0
    out["config3_b8_1gpu"] = {
        "workload": "512x512, 50 DDIM steps, batch 8 on ONE GPU (BASELINE config 3, single-GPU leg)",
        "images_per_s": B / (ms / 1e3), "ms_per_batch": ms, "unet_step_ms": step_ms,
        "roofline": {"bound": "tensor", "achieved": step_tf, "peak": tf, "unit": "TFLOP/s", "frac": step_tf / tf,
                     "what": "whole UNet step at B=8 (all kernels): algorithmic FLOP / graph-replay time",
                     "per_layer_roofline_ms": 4.411, "frac_of_per_layer_roofline": 4.411 / step_ms},
        "in_graph_ms_per_unet_step": {k: round(v, 3) for k, v in cls8.items()},
        "roofline_gemm": {"kernel": "gemm_tc_kernel / gemm2_kernel at batch 8", "bound": "tensor", "achieved": gemm_tf,
                          "peak": tf, "unit": "TFLOP/s", "frac": gemm_tf / tf,
                          "what": "conv + linear FLOP of one step / in-graph duration of the contraction kernels"},
        "roofline_attention": {"kernel": "attn_fwd_kernel at batch 8", "bound": "tensor", "achieved": attn_tf, "peak": tf,
                               "unit": "TFLOP/s", "frac": attn_tf / tf}}
    del d

    # ---- config 4: AutoencoderKL encode + decode, 512x512, batch 32 -------------------------------------------
    B = 32
    g = torch.Generator().manual_seed(1)
    x = (torch.rand((B, 3, PX, PX), generator=g) * 2 - 1).to(dev)
    vae = pipe.vae
    enc = lambda: vae.encode(x).latent_dist.mode()
    z = enc()
    dec = lambda: vae.decode(z).sample
    dec()
    rt = lambda: vae(x)["sample"]
    t_enc, t_dec, t_rt = _event_ms(enc, 2), _event_ms(dec, 2), _event_ms(rt, 2)
    enc_tf = VAE_ENC_GFLOP * B / 1e3 / (t_enc / 1e3)
    dec_tf = VAE_DEC_GFLOP * B / 1e3 / (t_dec / 1e3)
    out["config4_vae_b32"] = {
        "workload": "AutoencoderKL encode + decode, 512x512, batch 32, one GPU (BASELINE config 4)",
        "encode_images_per_s": B / (t_enc / 1e3), "decode_images_per_s": B / (t_dec / 1e3),
        "roundtrip_images_per_s": B / (t_rt / 1e3), "encode_ms": t_enc, "decode_ms": t_dec, "roundtrip_ms": t_rt,
        "precision": {"encoder": "fp16" if vae.enc_prec == 1 else "fp16x2 (3 tensor passes)",
                      "decoder": "fp16" if vae.dec_prec == 1 else "fp16x2 (3 tensor passes)"},
        "roofline": {"bound": "tensor", "unit": "TFLOP/s", "peak": tf,
                     "encode": {"achieved": enc_tf, "frac": enc_tf / tf, "per_layer_roofline_ms": 22.97},
                     "decode": {"achieved": dec_tf, "frac": dec_tf / tf, "per_layer_roofline_ms": 51.68},
                     "what": "algorithmic FLOP (each contraction counted once, whatever the pass count) / time"}}
    del x, z

    # ---- config 5: 768x768, batch 4, classifier-free guidance x2 (UNet batch 8 at 96x96 latents) ---------------
    B, px = 4, 768
    d = dev_inputs(B, px)
    neg = torch.randn((B, 577, 1024), generator=torch.Generator().manual_seed(11)).to(dev)
    call = lambda: pipe(masked_image=d["masked_image"], mask_image=d["mask"], glyph_embeds=d["glyph_embeds"],
                        negative_glyph_embeds=neg, guidance_scale=2.0, latents=d["latents"],
                        posterior_noise=d["posterior_noise"], num_inference_steps=NSTEPS).images
    ms = _median_call_ms(call, warmup=1, n=3)
    graph, _, _ = pipe._step_graph(2 * B, px // 8, px // 8, False)
    graph.replay()
    step_ms = _event_ms(graph.replay, 5)
    step_tf = (2226.91 - UNET_CTX_GFLOP) * 2 * B / 1e3 / (step_ms / 1e3)
    img_gflop = 2 * NSTEPS * (2226.91 - UNET_CTX_GFLOP) + 2 * UNET_CTX_GFLOP + 2609.12 + 5754.30
    out["config5_768_b4_cfg"] = {
        "workload": "768x768, 50 DDIM steps, batch 4, classifier-free guidance x2 (BASELINE config 5)",
        "images_per_s": B / (ms / 1e3), "ms_per_batch": ms, "unet_step_ms_batch8_96x96": step_ms,
        "image_tflops": img_gflop * B / 1e3 / (ms / 1e3),
        "roofline": {"bound": "tensor", "achieved": step_tf, "peak": tf, "unit": "TFLOP/s", "frac": step_tf / tf,
                     "what": "whole UNet step at UNet-batch 8, 96x96 latents: algorithmic FLOP / graph-replay time",
                     "per_layer_roofline_ms": 11.422, "frac_of_per_layer_roofline": 11.422 / step_ms}}
    del d

    # ---- the reference's own serving call: text_editing(photo, box) with ITS sampler (ancestral DDPM), uint8 in / out --
    import numpy as np
    from diffute_b200 import glue
    from diffute_b200.schedulers import DDPMScheduler
    photo = np.random.default_rng(5).integers(0, 256, (1080, 1440, 3), dtype=np.uint8)   # host memory, like cv2.imread
    box = (500, 400, 860, 470)
    emb = dev_inputs(1, PX)["glyph_embeds"]
    ddim = pipe.scheduler
    try:
        pipe.scheduler = DDPMScheduler(**{k: v for k, v in ddim.config.items() if k in DDPMScheduler._defaults})
        call = lambda: glue.text_editing(pipe, None, photo, NSTEPS, *box, glyph_embeds=emb, noise_seed=1)
        ms = _median_call_ms(call, warmup=2, n=3)
        reqs = [dict(instance_image=photo, bbox=box, glyph_embeds=emb) for _ in range(8)]
        ms8 = _median_call_ms(lambda: glue.text_editing_batch(pipe, reqs, NSTEPS, max_batch=8, noise_seed=1), warmup=2, n=3)
        pre = glue.preprocess(photo, box, device=dev)
        t_pre = _event_ms(lambda: glue.preprocess(photo, box, device=dev), 5)
        dec = torch.zeros((1, 3, PX, PX), device=dev)
        t_post = _event_ms(lambda: glue.composite(dec, pre).cpu(), 5)
    finally:
        pipe.scheduler = ddim
    out["text_editing_ddpm"] = {
        "workload": "reference serving call text_editing (app.ipynb:653-856): 1440x1080 uint8 photograph on the host -> crop "
                    "window / mask / resize / normalise on the GPU -> 50 ancestral DDPM steps (the reference's sampler, noise "
                    "generated in conv_out's epilogue) -> decode -> resize + paste on the GPU -> uint8 photograph on the host",
        "images_per_s": 1.0 / (ms / 1e3), "ms_per_image": ms,
        "queue_of_8_images_per_s": 8.0 / (ms8 / 1e3), "queue_of_8_ms": ms8,
        "queue_note": "text_editing_batch: eight pending requests served as one UNet batch on this GPU",
        "preprocess_ms_incl_h2d": t_pre, "composite_ms_incl_d2h": t_post,
        "h2d_bytes": int(photo.nbytes), "d2h_bytes": int(photo.nbytes)}
    return out


def run_gpu(args):
    # The CPU leg runs first, in a fresh process state (no CUDA context, no process group) — the same conditions as
    # `--impl reference` — and only at N=1: under torchrun the other ranks would spin in a barrier while rank 0 computes.
    cpu = None
    if int(os.environ.get("WORLD_SIZE", "1")) == 1 and not args.no_cpu_baseline:
        v_cpu, det = cpu_reference_sample()
        cpu = {"value": v_cpu, "unit": "images/s", "cores": det["cores"], "kind": "port",
               "sample": (f"{det['unet_reps']} UNet steps after 1 warm-up (median {det['unet_step_s']:.2f} s) + 1 VAE "
                          f"encode ({det['vae_encode_s']:.1f} s) + 1 VAE decode ({det['vae_decode_s']:.1f} s) at 512x512 "
                          f"on the fp32 CPU oracle, extrapolated to 50 steps")}
    import torch
    from diffute_b200 import arch, dist as ddist, ops, synthetic
    from diffute_b200.pipeline import DiffUTEPipeline

    rank, world, local = ddist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback; use --impl reference)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    up, vp = {"mixed": ("fp16", "fp16x2"), "fp16x2": ("fp16x2", "fp16x2"), "fp16": ("fp16", "fp16")}[args.precision]
    vep = "fp16" if args.precision == "mixed" else None  # mixed: the VAE encoder runs one fp16 pass

    # one NCCL broadcast of the fp32 weight arena (rank 0 -> all), outside the timed region
    ushapes, vshapes = arch.unet_param_shapes(), arch.vae_param_shapes()
    usd = synthetic.make_state_dict(ushapes) if rank == 0 else None
    vsd = synthetic.make_state_dict(vshapes) if rank == 0 else None
    usd = ddist.broadcast_state_dict(usd, ushapes, dev)
    vsd = ddist.broadcast_state_dict(vsd, vshapes, dev)
    pipe = DiffUTEPipeline.from_synthetic(up, vp, state_dicts=(usd, vsd), vae_encoder_precision=vep)
    del usd, vsd

    B = args.batch_per_gpu
    inp = synthetic.make_inputs(B, PX, PX, seed=rank)  # each rank works on its own images
    host = {k: inp[k].pin_memory() for k in ("masked_image", "mask", "glyph_embeds", "latents", "posterior_noise")}
    devin = {k: v.to(dev) for k, v in host.items()}

    def one_image_resident():
        return pipe(masked_image=devin["masked_image"], mask_image=devin["mask"], glyph_embeds=devin["glyph_embeds"],
                    latents=devin["latents"], posterior_noise=devin["posterior_noise"],
                    num_inference_steps=NSTEPS).images

    out_host = torch.empty((B, 3, PX, PX), dtype=torch.float32).pin_memory()

    def one_image_e2e():
        d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        img = pipe(masked_image=d["masked_image"], mask_image=d["mask"], glyph_embeds=d["glyph_embeds"],
                   latents=d["latents"], posterior_noise=d["posterior_noise"], num_inference_steps=NSTEPS).images
        out_host.copy_(img, non_blocking=True)
        return img

    def timed(fn, n):
        ddist.barrier(dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ddist.barrier(dev)
        return ddist.max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)

    for _ in range(max(args.warmup, 3)):
        one_image_resident()
    torch.cuda.synchronize()
    l0 = ops.STATS["launches"]
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    t_res = timed(one_image_resident, args.steps)
    eager_launches = ops.STATS["launches"] - l0
    clk = clocks.stop() if rank == 0 else None
    one_image_e2e()
    t_e2e = timed(one_image_e2e, args.steps)

    # ---- UNet step time and per-kernel roofline (CUDA events on the launch stream, eager replays) ----
    h = w = PX // 8
    graph, _, _ = pipe._step_graph(B, h, w, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    unet_ms = e0.elapsed_time(e1) / 20
    A = pipe.unet.arena
    lat = A.get("pipe.latents", (B, 4, h, w))
    srcs = [lat, A.get("pipe.mask", (B, 1, h, w)), A.get("pipe.masked", (B, 4, h, w))]
    state = A.get("pipe.state", (B + 2,))
    lat.copy_(devin["latents"])
    ops.PROFILE = {}
    l1 = ops.STATS["launches"]
    reps = 3
    for _ in range(reps):
        # Give the GPU a ~40 ms head start of busy-waiting so the ~500 launches (+ an event pair each) of the step are
        # queued before it reaches them: the event intervals are then GPU execution time, not host launch latency.
        torch.cuda._sleep(int(8e7))
        pipe.unet._forward_impl(B, h, w, srcs=srcs, t=state[:B])
    torch.cuda.synchronize()
    launches_per_unet_step = (ops.STATS["launches"] - l1) // reps
    prof, ops.PROFILE = ops.PROFILE, None
    per_kernel = {}
    for name, evs in prof.items():
        ms = sum(a.elapsed_time(b) for a, b, _, _ in evs) / reps
        per_kernel[name] = dict(ms_per_step=ms, launches=len(evs) // reps, gflop=sum(f for *_, f, _ in evs) / reps / 1e9,
                                gbytes=sum(b for *_, b in evs) / reps / 1e9)
    peaks = _peaks()
    passes = 3 if up == "fp16x2" else 1
    gk = [per_kernel.get("gemm_conv", {}), per_kernel.get("gemm_linear", {})]
    g_ms_events = sum(k.get("ms_per_step", 0.0) for k in gk)
    g_gf = sum(k.get("gflop", 0.0) for k in gk)  # algorithmic: each contraction counted once, whatever the passes
    g_n = sum(k.get("launches", 0) for k in gk)

    # In-graph cost of each kernel class by ablation: replay the captured step with that class's launches removed
    # (same buffers, same order, PDL edges intact) and take the difference — CUDA events on the replay stream, no
    # per-launch event overhead.  This is the duration used for the roofline.
    full_ms, in_graph = in_graph_class_ms(pipe, B, h, w)
    g_ms = in_graph["gemm"] if in_graph["gemm"] > 0 else g_ms_events
    achieved = (g_gf / 1e3) / (g_ms / 1e3) if g_ms > 0 else 0.0  # TFLOP/s
    traffic = None
    tp = os.path.join(ROOT, "profiles", "gemm_traffic_r02.json")  # ncu DRAM pass over one eager UNet step (B=1)
    if os.path.exists(tp) and B == 1:
        try:
            with open(tp) as f:
                traffic = json.load(f)["dram_bytes_per_launch"]
        except (ValueError, KeyError):
            traffic = None
    roofline = {"kernel": "gemm_tc_kernel (tcgen05 implicit-GEMM conv + linear)", "bound": "tensor",
                "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                "traffic": traffic, "traffic_note": "avg DRAM bytes per gemm launch over one UNet step (ncu, cold L2; "
                "profiles/gemm_dram_r02.csv): 2.19 GB per step = 1.73 GB of fp16 weights streamed once + operand reads; "
                "writes stay in L2",
                "peak_source": peaks["src"], "launches_per_unet_step": g_n,
                "avg_launch_us": (g_ms * 1e3 / g_n) if g_n else None,
                "algorithmic_gflop_per_unet_step": g_gf, "tensor_passes": passes,
                "unet_step_ms": unet_ms, "unet_step_frac_of_flop_roofline":
                    ((UNET_GFLOP - UNET_CTX_GFLOP) * B / 1e3) / (unet_ms / 1e3) / peaks["tflops"],
                "in_graph_ms_per_unet_step": {k: round(v, 4) for k, v in in_graph.items()},
                "in_graph_method": "captured step replayed with one kernel class removed; cost = full - ablated",
                "eager_event_ms_per_unet_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(per_kernel.items())}}
    # memory-bound kernels: algorithmic bytes (read fp32 once, write the fp16 operand planes) / in-graph duration
    roofline_norm = {"bound": "hbm", "peak": peaks["hbm"], "unit": "GB/s", "peak_source": peaks["src"],
                     "note": "at B=1 these tensors are L2-resident (<= 5 MB each): the figure is latency-bound, not "
                             "bandwidth-bound; per-kernel DRAM / L2 bytes from ncu: profiles/norm_kernels_summary_r02.txt"}
    for cls in ("groupnorm", "layernorm"):
        k = per_kernel.get(cls)
        if k and in_graph[cls] > 0:
            gbs = k["gbytes"] / (in_graph[cls] / 1e3)
            roofline_norm[cls] = {"achieved": gbs, "frac": gbs / peaks["hbm"], "launches": k["launches"],
                                  "algorithmic_gbytes_per_unet_step": k["gbytes"], "in_graph_ms": in_graph[cls]}

    total_images = B * world * args.steps
    value = total_images / t_res
    e2e_val = total_images / t_e2e
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = out_host.numel() * out_host.element_size()
    # launches inside the timed region: graph replays execute `launches_per_unet_step` kernels each without passing
    # through the Python wrappers, which only counted the eager ones (VAE, context projections, scheduler state)
    gpu_launches = eager_launches + args.steps * NSTEPS * launches_per_unet_step

    configs = None
    if world == 1 and not args.no_extra_configs:
        configs = extra_configs(pipe, dev, peaks)
    if rank == 0:
        line = {"metric": "images/sec at 512x512, 50 DDIM steps", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_res / args.steps * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"mixed": "f16 operands / f32 accumulate (UNet and VAE encoder 1 pass; VAE decoder 3-pass hi/lo split)",
                          "fp16x2": "f16 hi/lo split operands, 3 tensor passes, f32 accumulate (~fp32)",
                          "fp16": "f16 operands / f32 accumulate"}[args.precision],
                "data": "synthetic",
                "config": {"workload": "512x512 glyph-conditioned inpaint, 50 DDIM steps (BASELINE config 2), "
                                       f"{B} image(s) per GPU, batch sharded over GPUs with no per-step collective",
                           "global_batch": B * world, "precision": args.precision,
                           "l2": "working set (1.9-3.8 GB of packed weights per step) exceeds the 126 MB L2",
                           "parity": "decoded RGB vs fp32 CPU oracle at this exact workload and precision mode, bar 1e-3: "
                                     "tests/test_pipeline_gpu.py::test_benched_mode_parity_at_baseline_size",
                           "parallelism": f"dp{world}"},
                "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(gpu_launches), "clocks": clk, "roofline": roofline, "roofline_norm": roofline_norm,
                "cpu_baseline": cpu,
                "configs": configs, "image_gflop": IMAGE_GFLOP,
                "image_frac_of_flop_roofline": (IMAGE_GFLOP / 1e3) * B / (t_res / args.steps) / peaks["tflops"]}
        print(json.dumps(line))
    import torch.distributed as tdist
    if tdist.is_initialized():
        tdist.barrier()
        tdist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DFU_PRECISION", "mixed"), choices=["mixed", "fp16x2", "fp16"])
    ap.add_argument("--batch-per-gpu", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="skip BASELINE configs 3/4/5 (batch 8, VAE batch 32, 768x768 CFG); they only run at N=1")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
