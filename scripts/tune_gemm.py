"""Measure the best (block_n, splits, stages) per GEMM shape on the B200 and write diffute_b200/tuning_b200.json.

Runs the real orchestration eagerly (UNet step at the bench shape, VAE encode + decode) with ops.TUNER hooked: the
first time a shape is seen, every candidate tiling is timed with CUDA events on the very descriptor the engine built
(L2 flushed before each launch so weights stream from HBM as they do inside a 1.7 GB step), and the fastest wins.
Usage: python scripts/tune_gemm.py [batch] [px]"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import ops, synthetic, _lib
from diffute_b200.pipeline import DiffUTEPipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
px = int(sys.argv[2]) if len(sys.argv) > 2 else 512
modes = [("fp16", "fp16x2"), ("fp16x2", "fp16x2")]
L = _lib.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
log = {}
ops.TUNE_TABLE.clear()


NL = 6  # launches per timed window (CUDA events resolve ~2 us: one launch per window cannot rank candidates)


def _window(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ok = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 if ok else None


def _flush_only():
    for _ in range(NL):
        flush.zero_()
    return True


T_FLUSH = None


def time_cfg(d, reps=3):
    """median over `reps` windows of NL x (L2 flush + launch), minus the flush-only window, per launch"""
    global T_FLUSH
    if T_FLUSH is None:
        _window(_flush_only)
        T_FLUSH = sorted(_window(_flush_only) for _ in range(5))[2]
    st = torch.cuda.current_stream().cuda_stream

    def run():
        for _ in range(NL):
            flush.zero_()
            if L.dfu_gemm(C.byref(d), st) != 0:
                return False
        return True

    if _window(run) is None:
        return None
    ts = sorted(_window(run) for _ in range(reps))
    return max((ts[len(ts) // 2] - T_FLUSH) / NL, 0.01)


def tuner(d, ws, key):
    kb = sum(d.g[i].ntaps * (d.g[i].k_per_tap // 64) for i in range(d.ngroups)) * d.npass
    cands = []
    for bn in (256, 160, 128, 96, 64, 32):
        if d.n % bn:
            continue
        for sp in (1, 2, 3, 4, 6, 8, 12, 16, 24):
            if sp > 1 and kb // sp < 2:
                continue
            tiles = -(-d.m // 128) * (d.n // bn)
            if tiles * sp > 1200 or (sp > 1 and tiles * sp > 296):
                continue
            for st in ((3, 4, 6) if bn <= 160 else (3, 4)):
                cands.append((bn, sp, st))
    best, best_t, res = None, 1e30, []
    need_max = 32 * d.m * d.n * 4
    buf = ws.ensure(need_max)
    d.workspace, d.workspace_bytes = buf.data_ptr(), buf.numel()
    for bn, sp, st in cands:
        d.block_n, d.splits, d.stages = bn, sp, st
        t = time_cfg(d)
        if t is None:
            continue
        res.append((round(t, 2), bn, sp, st))
        if t < best_t:
            best, best_t = (bn, sp, st), t
    d.block_n = d.splits = d.stages = 0
    out = (C.c_int32 * 6)()
    L.dfu_gemm_plan(C.byref(d), out)
    d.block_n, d.splits, d.stages = out[0], out[1], out[2]
    t_model = time_cfg(d)
    d.block_n = d.splits = d.stages = 0
    res.sort()
    log[key] = dict(best=best, best_us=best_t, model=(out[0], out[1], out[2]), model_us=t_model, top=res[:6])
    print(f"{key:28s} best {best} {best_t:7.1f} us | cost-model {tuple(out[:3])} {t_model:7.1f} us", flush=True)
    return best


ops.TUNER = tuner
t0 = time.time()
for up, vp in modes:
    pipe = DiffUTEPipeline.from_synthetic(up, vp)
    pipe.unet.use_cuda_graph = False
    inp = synthetic.make_inputs(B, px, px)
    dev = pipe.device
    h = w = px // 8
    A = pipe.unet.arena
    pipe.unet.prepare_context(inp["glyph_embeds"].to(dev))
    lat = A.get("pipe.latents", (B, 4, h, w)); lat.copy_(inp["latents"])
    mask = A.get("pipe.mask", (B, 1, h, w)); mask.copy_(inp["mask"][:, :, ::8, ::8])
    ml = A.get("pipe.masked", (B, 4, h, w)); ml.copy_(inp["latents"] * 0.3)
    state = A.get("pipe.state", (B + 2,)); state.fill_(981.0)
    pipe.unet._forward_impl(B, h, w, srcs=[lat, mask, ml], t=state[:B])
    if (up, vp) == modes[0]:
        z = pipe.vae.encode(inp["masked_image"].to(dev)).latent_dist.mode()
        pipe.vae.decode(z)
    torch.cuda.synchronize()
    del pipe
    torch.cuda.empty_cache()
table = {k: list(v) for k, v in ops.TUNE_TABLE.items() if v is not None}
meta = dict(device=torch.cuda.get_device_name(0), batch=B, px=px, seconds=time.time() - t0,
            key="conv:m:n:k_blocks(64, all passes):epilogue", value="[block_n, splits, stages]")
os.makedirs("gpurun_out", exist_ok=True)
for path in (ops.TUNE_PATH, "gpurun_out/tuning_b200.json"):   # only gpurun_out/ travels back from the GPU box
    json.dump(dict(meta=meta, table=table), open(path, "w"), indent=0)
json.dump(dict(meta=meta, log=log), open("gpurun_out/tune_gemm_log.json", "w"), indent=0)
print(f"wrote {len(table)} entries in {time.time() - t0:.0f} s")
