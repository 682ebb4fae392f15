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
def cpu_reference_sample(threads=None, unet_reps=1):
    """Bounded sample of BASELINE config 2 on the CPU oracle: `unet_reps` UNet steps (64x64 latent) + one VAE encode +
    one VAE decode at 512x512, extrapolated to 50 steps.  Returns (images_per_s, detail dict)."""
    import torch
    from diffute_b200 import arch, synthetic
    from oracle import UNetOracle, VAEOracle
    threads = threads or _cpu_threads()
    torch.set_num_threads(threads)
    u, v = UNetOracle(), VAEOracle()
    u.load_state_dict(synthetic.make_state_dict(arch.unet_param_shapes()))
    v.load_state_dict(synthetic.make_state_dict(arch.vae_param_shapes()))
    inp = synthetic.make_inputs(1, PX, PX)
    x = torch.cat([inp["latents"], inp["mask"][:, :, ::8, ::8], inp["latents"]], 1)
    t_u = []
    for _ in range(unet_reps):
        t0 = time.perf_counter()
        u(x, 981, inp["glyph_embeds"])
        t_u.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    z = v.encode(inp["masked_image"]).latent_dist.mode()
    t_enc = time.perf_counter() - t0
    t0 = time.perf_counter()
    v.decode(z)
    t_dec = time.perf_counter() - t0
    tu = statistics.median(t_u)
    total = NSTEPS * tu + t_enc + t_dec
    return 1.0 / total, dict(unet_step_s=tu, vae_encode_s=t_enc, vae_decode_s=t_dec, image_s_extrapolated=total,
                             cores=threads)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    threads = _cpu_threads()
    from diffute_b200 import arch, synthetic
    from oracle import UNetOracle, VAEOracle
    torch.set_num_threads(threads)
    u, v = UNetOracle(), VAEOracle()
    u.load_state_dict(synthetic.make_state_dict(arch.unet_param_shapes()))
    v.load_state_dict(synthetic.make_state_dict(arch.vae_param_shapes()))
    inp = synthetic.make_inputs(1, PX, PX)
    x = torch.cat([inp["latents"], inp["mask"][:, :, ::8, ::8], inp["latents"]], 1)
    # VAE encode + decode measured once (warm-up leg); every timed step = one UNet forward, extrapolated
    t0 = time.perf_counter()
    z = v.encode(inp["masked_image"]).latent_dist.mode()
    v.decode(z)
    t_vae = time.perf_counter() - t0
    for _ in range(max(args.warmup - 1, 0)):
        u(x, 981, inp["glyph_embeds"])
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        u(x, 981, inp["glyph_embeds"])
        ts.append(time.perf_counter() - t0)
    per_image = NSTEPS * (sum(ts) / len(ts)) + t_vae
    val = 1.0 / per_image
    sample = (f"per step: 1 UNet forward (64x64 latent, B=1) on the fp32 CPU oracle, x{NSTEPS} + one VAE encode+decode "
              f"at 512x512 measured once ({t_vae:.1f} s); extrapolated to one 50-step image")
    line = {"impl": "reference", "metric": "images/sec at 512x512, 50 DDIM steps", "value": val, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_image * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "512x512 glyph-conditioned inpaint, 50 DDIM steps, batch 1 (BASELINE config 2)",
                       "arm": "CPU restatement of the reference's diffusers path (oracle/), torch fp32"},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# =================================================================================================
# GPU arm
# =================================================================================================
def run_gpu(args):
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
    def graph_ms(ablate):
        ops.ABLATE = set(ablate)
        try:
            pipe.unet._forward_impl(B, h, w, srcs=srcs, t=state[:B])
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                pipe.unet._forward_impl(B, h, w, srcs=srcs, t=state[:B])
        finally:
            ops.ABLATE = set()
        gr.replay()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(20):
            gr.replay()
        a1.record()
        torch.cuda.synchronize()
        return a0.elapsed_time(a1) / 20

    full_ms = graph_ms(())
    in_graph = {c: max(full_ms - graph_ms((c,)), 0.0) for c in ("gemm", "attention", "groupnorm", "layernorm")}
    g_ms = in_graph["gemm"] if in_graph["gemm"] > 0 else g_ms_events
    achieved = (g_gf / 1e3) / (g_ms / 1e3) if g_ms > 0 else 0.0  # TFLOP/s
    traffic = None
    tp = os.path.join(ROOT, "profiles", "gemm_traffic_r01.json")
    if os.path.exists(tp) and B == 1:
        with open(tp) as f:
            tj = json.load(f)
        traffic = tj["dram_bytes_total"] / tj["gemm_launches_per_unet_step"]  # bytes per launch (ncu capture)
    roofline = {"kernel": "gemm_tc_kernel (tcgen05 implicit-GEMM conv + linear)", "bound": "tensor",
                "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                "traffic": traffic, "traffic_note": "avg DRAM bytes per gemm launch over one UNet step (ncu, cold L2): "
                "2.15 GB per step = 1.73 GB of fp16 weights streamed once + operand reads; writes stay in L2",
                "peak_source": peaks["src"], "launches_per_unet_step": g_n,
                "avg_launch_us": (g_ms * 1e3 / g_n) if g_n else None,
                "algorithmic_gflop_per_unet_step": g_gf, "tensor_passes": passes,
                "unet_step_ms": unet_ms, "unet_step_frac_of_flop_roofline":
                    ((UNET_GFLOP - UNET_CTX_GFLOP) * B / 1e3) / (unet_ms / 1e3) / peaks["tflops"],
                "in_graph_ms_per_unet_step": {k: round(v, 4) for k, v in in_graph.items()},
                "in_graph_method": "captured step replayed with one kernel class removed; cost = full - ablated",
                "eager_event_ms_per_unet_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(per_kernel.items())}}
    gn = per_kernel.get("groupnorm")
    if gn and in_graph["groupnorm"] > 0:
        roofline["groupnorm_hbm_gbs"] = gn["gbytes"] / (in_graph["groupnorm"] / 1e3)
        roofline["groupnorm_frac_of_hbm"] = roofline["groupnorm_hbm_gbs"] / peaks["hbm"]

    total_images = B * world * args.steps
    value = total_images / t_res
    e2e_val = total_images / t_e2e
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = out_host.numel() * out_host.element_size()
    # launches inside the timed region: graph replays execute `launches_per_unet_step` kernels each without passing
    # through the Python wrappers, which only counted the eager ones (VAE, context projections, scheduler state)
    gpu_launches = eager_launches + args.steps * NSTEPS * launches_per_unet_step

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        v_cpu, det = cpu_reference_sample(unet_reps=2)
        cpu = {"value": v_cpu, "unit": "images/s", "cores": det["cores"], "kind": "port",
               "sample": (f"2 UNet steps (median {det['unet_step_s']:.2f} s) + 1 VAE encode ({det['vae_encode_s']:.1f} s) "
                          f"+ 1 VAE decode ({det['vae_decode_s']:.1f} s) at 512x512 on the fp32 CPU oracle, "
                          f"extrapolated to 50 steps")}
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
                           "parity": "decoded RGB vs fp32 CPU oracle at this exact workload: max rel err 3.0e-4 (mixed), "
                                     "1.4e-5 (fp16x2); bar 1e-3 (profiles/parity_512_r01.json, tests/test_pipeline_gpu.py)",
                           "parallelism": f"dp{world}"},
                "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(gpu_launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
                "image_gflop": IMAGE_GFLOP,
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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
