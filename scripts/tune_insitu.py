"""In-situ GEMM tiling tuner: every candidate (block_n, splits, stages) of every contraction shape of the UNet step is
timed INSIDE the captured step — graph replay, PDL overlap, weights streaming from HBM, activations in L2 — from the
in-kernel timeline records of the diagnostic library (diffute_b200/trace.py).  The cost of a launch is its contribution
to the step's critical path: (its last CTA's exit [+ the split-K reduce launch that follows]) - (previous kernel's
last exit).  Round r gives every shape its r-th candidate; a second phase re-times the best few per shape with all
other shapes at their winners.
  DFU_TRACE=1 python scripts/tune_insitu.py [batch] [px] [unet|vae]  ->  gpurun_out/tuning_b200.json (+ log)
`vae` tunes the shapes of one VAE encode + decode (captured as a graph the same way) instead of the UNet step."""
import ctypes as C, json, os, sys, time
os.environ["DFU_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import ops, synthetic, trace, _lib
from diffute_b200.pipeline import DiffUTEPipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
px = int(sys.argv[2]) if len(sys.argv) > 2 else 512
WHAT = sys.argv[3] if len(sys.argv) > 3 else "unet"
MAXR = int(os.environ.get("TUNE_ROUNDS", "120"))
pipe = DiffUTEPipeline.from_synthetic("fp16", "fp16x2", vae_encoder_precision="fp16")
inp = synthetic.make_inputs(B, px, px)
dev = pipe.device
h = w = px // 8
A = pipe.unet.arena
pipe.unet.prepare_context(inp["glyph_embeds"].to(dev))
lat = A.get("pipe.latents", (B, 4, h, w)); lat.copy_(inp["latents"])
mask = A.get("pipe.mask", (B, 1, h, w)); mask.copy_(inp["mask"][:, :, ::8, ::8])
ml = A.get("pipe.masked", (B, 4, h, w)); ml.copy_(inp["latents"] * 0.3)
state = A.get("pipe.state", (B + 2,)); state.fill_(981.0)
tproj = A.get("t.proj", (B, pipe.unet.temb_total))
pipe.unet.time_projections(state[:B], tproj)
pipe.unet.ws.ensure(512 << 20)
L = _lib.lib()

ORDER = []      # gemm keys in launch order (one captured step)
SHAPES = {}     # key -> descriptor facts
_lg = ops.launch_gemm


def launch_gemm(d, ws=None):
    key = ops.gemm_key(d)
    ORDER.append(key)
    if key not in SHAPES:
        kb = sum(d.g[i].ntaps * (d.g[i].k_per_tap // 64) for i in range(d.ngroups)) * d.npass
        SHAPES[key] = {"m": d.m, "n": d.n, "kb": kb, "epi": d.epi, "conv": d.conv, "B": d.B, "H": d.H, "W": d.W}
    _lg(d, ws)


ops.launch_gemm = launch_gemm


x_img = inp["masked_image"].to(dev)
lat_in = inp["latents"].to(dev)
pipe.vae.ws.ensure(512 << 20)


def step():
    if WHAT == "vae":
        pipe.vae.encode(x_img)
        return pipe.vae.decode(lat_in, pre_scale=1 / 0.18215)
    return pipe.unet._forward_impl(B, h, w, srcs=[lat, mask, ml], t=state[:B], tproj=tproj)


def tiles_m(s):
    if not s["conv"]:
        return (s["m"] + 127) // 128
    W, H, Bn = s["W"], s["H"], s["B"]
    bw = min(W, 128); bh = 1; bn = 1
    if W < 128:
        bh = min(128 // W, H)
        if bh == H:
            bn = max(1, min(128 // (H * W), Bn))
    return ((W + bw - 1) // bw) * ((H + bh - 1) // bh) * ((Bn + bn - 1) // bn)


def candidates(s):
    out = []
    for bn in (256, 192, 160, 128, 96, 80, 64, 32):
        if s["n"] % bn:
            continue
        if s["epi"] == 2 and bn % 32:
            continue
        tiles = tiles_m(s) * (s["n"] // bn)
        stage_b = 16384 + bn * 128
        for sp in (1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 20):
            if sp > 1 and s["kb"] // sp < 2:
                continue
            ctas = tiles * sp
            if ctas > 2 * 148 and sp > 1:
                continue
            if sp > 1 and ctas < 64:
                pass
            if sp * s["m"] * s["n"] * 4 > (400 << 20):
                continue
            kb_cta = -(-s["kb"] // sp)
            st_half = max(2, min(kb_cta, (110 * 1024) // stage_b, 12))
            st_full = max(2, min(kb_cta, (222 * 1024) // stage_b, 12))
            for st in sorted({st_half, st_full}):
                out.append((bn, sp, st))
    # prune: drop configurations that leave most of the chip idle unless nothing else exists
    good = [c for c in out if tiles_m(s) * (s["n"] // c[0]) * c[1] >= 48]
    good = good or out
    # persistent CTA-pair kernel (cta_group::2): 256 x bn tiles, no split-K; stages 0 = as deep as shared memory allows
    for bn in (256, 192, 160, 128, 96, 80, 64):
        if s["n"] % bn or (s["epi"] == 2 and bn % 32):
            continue
        if ((tiles_m(s) + 1) // 2) * (s["n"] // bn) >= 24 and s["kb"] >= 3:
            good.append((bn, 1, 0, 2))
    return good


def measure(nrep=2):
    """capture the step with the current TUNE_TABLE, replay, return per-launch-index cost list for gemm launches"""
    ORDER.clear()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    order = list(ORDER)
    g.replay()
    costs = None
    for _ in range(nrep):
        trace.reset()
        g.replay()
        recs = trace.collect_ends()
        gi = 0
        prev_end = None
        cur = []
        pending = None  # index in cur of a gemm whose split-K reduce launch follows
        for d in recs:
            end = d.get("end_last", d.get("end_clk_last", d["start_last"]))
            dt = 0.0 if prev_end is None else max(0.0, end - prev_end)
            if d["kernel"] == "gemm":
                cur.append(dt)
                pending = len(cur) - 1
                gi += 1
            elif d["kernel"] == "splitk_reduce" and pending is not None:
                cur[pending] += dt
                pending = None
            else:
                pending = None
            prev_end = end if prev_end is None else max(prev_end, end)
        if gi != len(order):
            raise RuntimeError(f"traced {gi} gemm launches, host logged {len(order)}")
        costs = cur if costs is None else [min(a, b) for a, b in zip(costs, cur)]
    span = recs[-1].get("end_last", recs[-1]["start_last"]) - recs[0]["start_first"]
    del g
    return order, costs, span


# discover shapes with the shipped table
for _ in range(2):
    step()
torch.cuda.synchronize()
trace.enable(1 << 20 if (WHAT == "vae" or B > 1) else 1 << 18)
base_order, base_costs, base_span = measure()
print(f"baseline span {base_span:.1f} us, {len(base_order)} gemm launches, {len(SHAPES)} shapes", flush=True)
base_table = dict(ops.TUNE_TABLE)
base_by_key = {}
for k, c in zip(base_order, base_costs):
    base_by_key[k] = base_by_key.get(k, 0.0) + c
CANDS = {k: candidates(s) for k, s in SHAPES.items()}
nr = min(MAXR, max(len(v) for v in CANDS.values()))
print("candidates per shape:", {k: len(v) for k, v in CANDS.items()}, "rounds", nr, flush=True)
RESULT = {k: {} for k in SHAPES}
t0 = time.time()
for r in range(nr):
    for k, cs in CANDS.items():
        ops.TUNE_TABLE[k] = cs[r % len(cs)]
    try:
        order, costs, span = measure(1)
    except Exception as e:  # a bad candidate must not end the run
        print("round", r, "failed:", repr(e)[:200], flush=True)
        continue
    for k, c in zip(order, costs):
        cfg = tuple(ops.TUNE_TABLE[k])
        RESULT[k].setdefault(cfg, []).append(c)
print(f"phase 1: {nr} rounds in {time.time() - t0:.1f} s", flush=True)


def per_key_cost(samples, count):
    """samples = per-launch costs collected over rounds; a shape launched `count` times per step contributes `count`
    samples per round: sum per round = mean * count"""
    return sum(samples) / len(samples) * count


COUNT = {}
for k in base_order:
    COUNT[k] = COUNT.get(k, 0) + 1
best = {}
top = {}
for k in SHAPES:
    ranked = sorted(((per_key_cost(v, COUNT[k]), cfg) for cfg, v in RESULT[k].items()))
    top[k] = [cfg for _, cfg in ranked[:4]]
    best[k] = ranked[0][1] if ranked else base_table.get(k)
# phase 2: re-time the top few of each shape with everything else at its winner (also the shipped entry)
FINAL = {k: {} for k in SHAPES}
for k in SHAPES:
    if base_table.get(k) is not None and tuple(base_table[k]) not in top[k]:
        top[k].append(tuple(base_table[k]))
for r in range(max(len(v) for v in top.values())):
    for rep in range(2):
        for k in SHAPES:
            ops.TUNE_TABLE[k] = top[k][r] if r < len(top[k]) else best[k]
        try:
            order, costs, span = measure(2)
        except Exception as e:
            print("phase-2 round", r, "failed:", repr(e)[:200], flush=True)
            continue
        for k, c in zip(order, costs):
            cfg = tuple(ops.TUNE_TABLE[k])
            if r < len(top[k]) and cfg == top[k][r]:
                FINAL[k].setdefault(cfg, []).append(c)
table = {}
for k in SHAPES:
    ranked = sorted(((per_key_cost(v, COUNT[k]), cfg) for cfg, v in FINAL[k].items()))
    if ranked:
        table[k] = list(ranked[0][1])
        print(f"{k:28s} x{COUNT[k]:2d} best {ranked[0][1]} {ranked[0][0]:7.1f} us | shipped {base_table.get(k)} "
              f"{base_by_key.get(k, 0):7.1f} us | runner-up {ranked[1][1] if len(ranked) > 1 else None}", flush=True)
for k, v in table.items():
    ops.TUNE_TABLE[k] = tuple(v)
order, costs, span = measure(2)
print(f"tuned span {span:.1f} us (baseline {base_span:.1f}); gemm critical-path sum {sum(costs):.1f} (baseline {sum(base_costs):.1f})")
os.makedirs("gpurun_out", exist_ok=True)
merged = {k: list(v) for k, v in base_table.items()}
merged.update(table)
OUT = os.environ.get("TUNE_OUT", WHAT)
with open(f"gpurun_out/tuning_b200_{OUT}.json", "w") as f:
    json.dump({"meta": {"device": torch.cuda.get_device_name(0), "batch": B, "px": px, "method": "in-situ graph replay, in-kernel timestamps (scripts/tune_insitu.py)",
                        "key": "conv:m:n:k_blocks(64, all passes):epilogue", "value": "[block_n, splits, stages(, kernel: 1 tile-per-CTA, 2 persistent CTA pairs)]",
                        "span_us": span, "baseline_span_us": base_span}, "table": merged}, f, indent=0)
with open(f"gpurun_out/tune_insitu_log_{OUT}.json", "w") as f:
    json.dump({k: {str(cfg): v for cfg, v in RESULT[k].items()} for k in RESULT}, f)
