"""Full-size parity on the GPU box: 512x512, 50 DDIM steps, decoded RGB vs the fp32 CPU oracle (BASELINE bar 1e-3).
Writes gpurun_out/parity_512.json.  Usage: [PARITY_SEED=n] python scripts/parity_512.py [px] [steps] [modes...]
PARITY_SEED selects other synthetic inputs (latents, image, mask stay a text-line box, glyph embedding) and, with
PARITY_WSEED, other synthetic weights: the 1e-3 bar is not a property of one draw."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import arch, synthetic
from diffute_b200.pipeline import DiffUTEPipeline
from oracle import DDIMOracle, DDPMOracle, UNetOracle, VAEOracle, philox, sample_loop

px = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
modes = sys.argv[3:] or ["mixed", "fp16x2"]
torch.set_num_threads(min(os.cpu_count(), int(os.environ.get('DFU_CPU_THREADS', '32'))))
seed = int(os.environ.get("PARITY_SEED", "0"))
wseed = int(os.environ.get("PARITY_WSEED", "1234"))
usd = synthetic.make_state_dict(arch.unet_param_shapes(), wseed)
vsd = synthetic.make_state_dict(arch.vae_param_shapes(), wseed)
inp = synthetic.make_inputs(1, px, px, seed=seed)
uo, vo = UNetOracle(), VAEOracle()
uo.load_state_dict(usd); vo.load_state_dict(vsd)
SCHED = os.environ.get("PARITY_SCHED", "ddim")   # ddpm: the reference's ancestral sampler, noise from the Philox stream
NOISE_SEED = 20261017
t0 = time.time()
if SCHED == "ddim":
    ref = sample_loop(uo, vo, DDIMOracle(), inp["masked_image"], inp["mask"], inp["glyph_embeds"], inp["latents"], steps,
                      posterior_noise=inp["posterior_noise"])
else:
    import torch.nn.functional as F
    o = DDPMOracle()
    o.set_timesteps(steps)
    mask_l = F.interpolate(inp["mask"], size=(px // 8, px // 8))
    ml = vo.encode(inp["masked_image"]).latent_dist.sample(noise=inp["posterior_noise"]) * 0.18215
    lat = inp["latents"].clone()
    for i, t in enumerate(o.timesteps):
        eps = uo(torch.cat([lat, mask_l, ml], 1), t, inp["glyph_embeds"]).sample
        z = torch.from_numpy(philox.normal(NOISE_SEED, i, lat.numel())).view_as(lat)
        lat = o.step(eps, t, lat, noise=z).prev_sample
    ref = vo.decode(lat / 0.18215).sample
t_cpu = time.time() - t0
res = {"px": px, "steps": steps, "scheduler": SCHED, "input_seed": seed, "weight_seed": wseed, "cpu_oracle_seconds": t_cpu, "cpu_threads": torch.get_num_threads(), "modes": {}}
for m in modes:
    up, vp = {"mixed": ("fp16", "fp16x2"), "fp16x2": ("fp16x2", "fp16x2"), "fp16": ("fp16", "fp16")}[m]
    pipe = DiffUTEPipeline.from_synthetic(up, vp, state_dicts=(usd, vsd),
                                          vae_encoder_precision="fp16" if m == "mixed" else None)
    kw = {}
    if SCHED == "ddpm":
        from diffute_b200.schedulers import DDPMScheduler
        pipe.scheduler = DDPMScheduler()
        kw["noise_seed"] = NOISE_SEED
    out = pipe(masked_image=inp["masked_image"], mask_image=inp["mask"], glyph_embeds=inp["glyph_embeds"],
               latents=inp["latents"], posterior_noise=inp["posterior_noise"], num_inference_steps=steps, **kw).images.cpu()
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    res["modes"][m] = {"rgb_max_rel_err": err, "passes_1e-3": err <= 1e-3}
    print(m, err, flush=True)
    del pipe
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
tag = ("" if (seed, wseed) == (0, 1234) else f"_s{seed}_w{wseed}") + ("" if SCHED == "ddim" else "_" + SCHED)
json.dump(res, open(f"gpurun_out/parity_512{tag}.json", "w"), indent=1)
print(json.dumps(res))
