"""The diagnostic -DDFU_TRACE library: every CTA leaves a timeline record, launches come back in order with sane phase
stamps, and the product library refuses tracing (it contains no tracing code).  Runs the traced part in a subprocess
because the library variant is chosen once per process (DFU_TRACE=1)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r"""
import json, sys, torch
sys.path.insert(0, %r)
from diffute_b200 import ops, trace
trace.enable(1 << 14)
x = torch.randn(1024, 640, device="cuda")
g = torch.ones(640, device="cuda"); b = torch.zeros(640, device="cuda")
ln = torch.empty(1, 1024, 640, dtype=torch.float16, device="cuda")
w16 = ops.pack_linear_weight(torch.randn(640, 640, device="cuda") * 0.05, 1)
out = torch.empty(1024, 640, device="cuda")
for _ in range(2):
    trace.reset()
    ops.layernorm(x, g, b, 1e-5, ln)
    ops.linear(ln, w16, 640, 1, out_f32=out, residual=x)
    recs = trace.collect()
ref = torch.nn.functional.layer_norm(x, (640,)).half().float() @ w16.float().t() + x
err = ((out - ref).abs().max() / ref.abs().max()).item()
print(json.dumps({"recs": [{k: r[k] for k in ("kernel", "ctas", "nctas", "start_first", "wait_first", "end_last") if k in r}
                           | {"p8": r.get("p8_med"), "p7": r.get("p7_med")} for r in recs], "err": err}))
""" % ROOT


@pytest.mark.gpu
def test_trace_library_records_every_cta():
    env = dict(os.environ, DFU_TRACE="1")
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["err"] < 2e-3                      # the traced build computes the same thing
    kinds = [x["kernel"] for x in d["recs"]]
    assert kinds[0] == "layernorm" and "gemm" in kinds
    for x in d["recs"]:
        assert x["ctas"] == x["nctas"] > 0                          # one record per CTA
        assert x["start_first"] <= x["wait_first"] <= x["end_last"]  # entry <= PDL wait released <= last exit
    ln, gm = d["recs"][0], d["recs"][kinds.index("gemm")]
    assert gm["wait_first"] >= ln["end_last"] - 0.5   # the GEMM's producer waits for the LayerNorm (PDL), us
    assert gm["p7"] <= gm["p8"] <= gm["end_last"]     # first full stage <= accumulator complete <= exit


@pytest.mark.gpu
def test_product_library_has_no_tracing():
    import torch
    from diffute_b200 import _lib
    if os.environ.get("DFU_TRACE") == "1":
        pytest.skip("diagnostic library selected for this process")
    L = _lib.lib()
    buf = torch.zeros(64, dtype=torch.int64, device="cuda")
    for n in ("gemm", "gemm2", "attn", "norm", "misc"):
        assert getattr(L, f"dfu_trace_set_{n}")(buf.data_ptr()) == -1
