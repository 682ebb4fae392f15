"""Timeline of ONE graph-replayed UNet denoising step from in-kernel records (diagnostic library, DFU_TRACE=1):
  DFU_TRACE=1 python scripts/trace_step.py [mixed|fp16x2|fp16] [batch] > gpurun_out/trace_step.txt
Per launch: wait = griddepcontrol.wait released (first CTA), gap = that minus the previous kernel's last exit,
body = last exit minus first release.  Writes gpurun_out/trace_step.json as well."""
import ctypes as C, json, os, sys
os.environ["DFU_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffute_b200 import ops, synthetic, trace, _lib
from diffute_b200.pipeline import DiffUTEPipeline

mode = sys.argv[1] if len(sys.argv) > 1 else "mixed"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
up, vp = {"mixed": ("fp16", "fp16x2"), "fp16x2": ("fp16x2", "fp16x2"), "fp16": ("fp16", "fp16")}[mode]
pipe = DiffUTEPipeline.from_synthetic(up, vp, vae_encoder_precision="fp16" if mode == "mixed" else None)
inp = synthetic.make_inputs(B, 512, 512)
dev = pipe.device
h = w = 64
A = pipe.unet.arena
pipe.unet.prepare_context(inp["glyph_embeds"].to(dev))
lat = A.get("pipe.latents", (B, 4, h, w)); lat.copy_(inp["latents"])
mask = A.get("pipe.mask", (B, 1, h, w)); mask.copy_(inp["mask"][:, :, ::8, ::8])
ml = A.get("pipe.masked", (B, 4, h, w)); ml.copy_(inp["latents"] * 0.3)
state = A.get("pipe.state", (B + 2,)); state.fill_(981.0)
tproj = A.get("t.proj", (B, pipe.unet.temb_total))
pipe.unet.time_projections(state[:B], tproj)

# host-side log of what each launch is, in order, per kernel class
LOG = {"gemm": [], "attn": [], "gn": [], "layernorm": []}
_lg = ops.launch_gemm


def launch_gemm(d, ws=None):
    _lg(d, ws)
    plan = (C.c_int32 * 8)()
    _lib.lib().dfu_gemm_plan(C.byref(d), plan)
    k = sum(d.g[i].ntaps * d.g[i].k_per_tap for i in range(d.ngroups))
    LOG["gemm"].append({"conv": d.conv, "m": d.m, "n": d.n, "k": k, "epi": d.epi, "block_n": plan[0], "splits": plan[1],
                        "stages": plan[2], "tiles": plan[3] * plan[4], "kb": plan[5],
                        "gflop": 2e-9 * d.m * d.n * k, "wbytes": 2.0 * d.n * k})


ops.launch_gemm = launch_gemm
_at = ops.attention


def attention(q16, q_col0, k16, k_col0, v16, v_col0, B, heads, Nq, Nk, scale, out16, **kw):
    _at(q16, q_col0, k16, k_col0, v16, v_col0, B, heads, Nq, Nk, scale, out16, **kw)
    LOG["attn"].append({"heads": heads, "Nq": Nq, "Nk": Nk, "gflop": 4e-9 * B * heads * Nq * Nk * 64})


ops.attention = attention


def step():
    return pipe.unet._forward_impl(B, h, w, srcs=[lat, mask, ml], t=state[:B], tproj=tproj)


for _ in range(2):
    step()
torch.cuda.synchronize()
for k in LOG:
    LOG[k].clear()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
trace.enable(1 << 18)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
trace.reset()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1)
L = trace.collect()
gi = ai = 0
prev_end = None
rows = []
for d in L:
    info = ""
    if d["kernel"] == "gemm" and gi < len(LOG["gemm"]):
        x = LOG["gemm"][gi]; gi += 1
        d["desc"] = x
        info = (f"{'conv' if x['conv'] else 'lin '} m{x['m']:5d} n{x['n']:5d} k{x['k']:6d} bn{x['block_n']:3d} s{x['splits']:2d} "
                f"st{x['stages']} {x['gflop']:6.2f}GF {x['wbytes'] / 1e6:6.1f}MB")
    elif d["kernel"] == "attn" and ai < len(LOG["attn"]):
        x = LOG["attn"][ai]; ai += 1
        d["desc"] = x
        info = f"h{x['heads']} Nq{x['Nq']} Nk{x['Nk']} {x['gflop']:.2f}GF kvs{d['extra']}"
    end = d.get("end_last", d.get("end_clk_last", d["start_last"]))
    wait = d.get("wait_first", d["start_first"])
    d["gap"] = (wait - prev_end) if prev_end is not None else 0.0
    d["body"] = end - wait
    d["end"] = end
    d["info"] = info
    prev_end = end if prev_end is None else max(prev_end, end)
    rows.append(d)
print(f"step {step_ms * 1e3:.1f} us by CUDA events; {len(L)} launches traced; timeline span "
      f"{rows[-1]['end'] - rows[0]['start_first']:.1f} us")
print(f"{'#':>4} {'kernel':13} {'ctas':>5} {'start':>8} {'wait':>8} {'gap':>6} {'body':>7} | setup->wait first-full mma-done epi-done (us after wait, median CTA) | info")
for i, d in enumerate(rows):
    w0 = d.get("wait_med", d["start_first"])
    ph = ""
    if d["kernel"] == "gemm":
        ph = (f"pro {d.get('setup_med', 0) - d['start_first']:5.1f} ff {d.get('p7_med', w0) - w0:5.1f} mma {d.get('p8_med', w0) - w0:5.1f} "
              f"epi {d.get('p9_med', w0) - w0:5.1f} end {d.get('end_clk_med', w0) - w0:5.1f}"
              f" [ld {d.get('x12_med', w0) - w0:4.1f} stg {d.get('x14_med', w0) - w0:4.1f} st1 {d.get('x15_med', w0) - w0:4.1f}"
              f" c1 {d.get('x13_med', w0) - w0:4.1f}]")
    elif d["kernel"] == "attn":
        ph = f"pro {d.get('setup_med', 0) - d['start_first']:5.1f} main {d.get('p8_med', w0) - w0:5.1f} end {d.get('end_clk_med', w0) - w0:5.1f}"
    print(f"{i:4d} {d['kernel']:13} {d['nctas']:5d} {d['start_first']:8.1f} {d.get('wait_first', 0):8.1f} {d['gap']:6.1f} {d['body']:7.1f} | {ph} | {d['info']}")
agg = {}
for d in rows:
    a = agg.setdefault(d["kernel"], [0, 0.0, 0.0])
    a[0] += 1; a[1] += d["gap"]; a[2] += d["body"]
print("\nper kernel class: launches, sum gap us, sum body us")
for k, a in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
    print(f"  {k:14} {a[0]:4d} {a[1]:8.1f} {a[2]:8.1f}")
os.makedirs("gpurun_out", exist_ok=True)
if os.environ.get("TRACE_DUMP"):
    # raw per-CTA records of the n-th launch of one kernel class: TRACE_DUMP=attn:0
    kname, nth = os.environ["TRACE_DUMP"].split(":")
    gids = [d["grid_id"] for d in rows if d["kernel"] == kname]
    gid = gids[int(nth)]
    raw = trace._buf[8:8 + int(trace._buf[0].item()) * trace.REC].view(-1, trace.REC).cpu().numpy()
    sel = raw[raw[:, 0] == gid]
    t0 = sel[:, 3].min()
    out = [{"bid": int(r[1] >> 32), "smid": int(r[2] & 0xFFFFFFFF), "start": (int(r[3]) - int(t0)) / 1e3,
            "marks": [(int(r[i]) - int(r[4])) / 1965.0 if r[i] else None for i in range(5, 11)],
            "end": (int(r[11]) - int(t0)) / 1e3} for r in sel]
    with open("gpurun_out/trace_dump.json", "w") as f:
        json.dump(out, f)
with open("gpurun_out/trace_step.json", "w") as f:
    json.dump({"step_us": step_ms * 1e3, "launches": rows}, f)
