"""In-kernel timeline tracing (diagnostic build only: DFU_TRACE=1 loads libdiffute_b200_trace.so).

Every CTA of every traced kernel appends one record (csrc/common.cuh): grid id, kernel tag, SM id, globaltimer at
entry / exit and clock64 phase stamps.  `collect()` groups the records per launch and converts the per-SM clock64
stamps to the globaltimer axis, so that a whole captured UNet step can be laid out on one timeline: when each kernel's
first CTA started, when griddepcontrol.wait released it, when its MMAs / epilogue finished and when its last CTA left.
"""
from __future__ import annotations

import collections
import ctypes as C
from typing import Dict, List

import torch

from . import _lib

TAGS = {1: "gemm", 2: "splitk_reduce", 3: "attn", 4: "attn_merge", 5: "gn_stats", 6: "gn_finalize", 7: "gn_apply",
        8: "gn_fused", 9: "layernorm", 10: "cast", 11: "temb", 12: "gemv", 13: "conv_in", 14: "conv_out", 15: "misc"}
REC = 16
_buf = None


def enable(capacity: int = 1 << 20) -> torch.Tensor:
    """Allocate the record buffer and hand it to every translation unit of the trace library."""
    global _buf
    L = _lib.lib()
    _buf = torch.zeros(8 + capacity * REC, dtype=torch.int64, device="cuda")
    _buf[1] = capacity
    torch.cuda.synchronize()
    for n in ("gemm", "gemm2", "attn", "norm", "misc"):
        rc = getattr(L, f"dfu_trace_set_{n}")(C.c_void_p(_buf.data_ptr()))
        if rc != 0:
            raise _lib.DfuError("tracing needs the diagnostic library: run with DFU_TRACE=1")
    return _buf


def reset():
    _buf[0] = 0
    torch.cuda.synchronize()


def disable():
    L = _lib.lib()
    for n in ("gemm", "gemm2", "attn", "norm", "misc"):
        getattr(L, f"dfu_trace_set_{n}")(None)


def collect(sm_mhz: float = 1965.0) -> List[Dict]:
    """-> one dict per launch, in start order, times in microseconds relative to the first record."""
    torch.cuda.synchronize()
    n = int(_buf[0].item())
    cap = int(_buf[1].item())
    n = min(n, cap)
    r = _buf[8:8 + n * REC].view(n, REC).cpu().numpy().astype("float64")
    raw = _buf[8:8 + n * REC].view(n, REC).cpu().numpy()
    if n == 0:
        return []
    t0 = raw[:, 3].min()
    launches = collections.OrderedDict()
    for i in range(n):
        gid = int(raw[i, 0])
        launches.setdefault(gid, []).append(i)
    out = []
    cyc = 1.0 / sm_mhz  # us per SM cycle
    for gid, idx in launches.items():
        rr = raw[idx]
        tag = int(rr[0, 1]) & 0xFF
        extra = (int(rr[0, 1]) & 0xFFFFFFFF) >> 8
        gt_s = (rr[:, 3] - t0) / 1e3
        gt_e = (rr[:, 11] - t0) / 1e3
        clk0 = rr[:, 4].astype("float64")

        def ph(slot):  # phase stamp on the globaltimer axis (entry globaltimer + clock64 delta)
            v = rr[:, slot].astype("float64")
            ok = v > 0
            return (gt_s + (v - clk0) * cyc), ok

        d = {"grid_id": gid, "kernel": TAGS.get(tag, str(tag)), "extra": extra, "ctas": len(idx),
             "nctas": int(rr[0, 2] >> 32), "start_first": float(gt_s.min()), "start_last": float(gt_s.max())}
        for name, slot in (("setup", 5), ("wait", 6), ("p7", 7), ("p8", 8), ("p9", 9), ("end_clk", 10), ("x12", 12),
                           ("x13", 13), ("x14", 14), ("x15", 15)):
            v, ok = ph(slot)
            if ok.any():
                d[name + "_first"] = float(v[ok].min())
                d[name + "_last"] = float(v[ok].max())
                d[name + "_med"] = float(sorted(v[ok])[int(ok.sum()) // 2])
        has_end = rr[:, 11] > 0
        if has_end.any():
            d["end_first"] = float(gt_e[has_end].min())
            d["end_last"] = float(gt_e[has_end].max())
        out.append(d)
    out.sort(key=lambda d: d["start_first"])
    return out


def collect_ends() -> List[Dict]:
    """Vectorised subset of collect(): per launch {kernel, start_first, end_last}, start order (the tuner's hot loop)."""
    import numpy as np
    torch.cuda.synchronize()
    n = min(int(_buf[0].item()), int(_buf[1].item()))
    if n == 0:
        return []
    raw = _buf[8:8 + n * REC].view(n, REC).cpu().numpy()
    gid = raw[:, 0]
    uniq, inv = np.unique(gid, return_inverse=True)
    t0 = raw[:, 3].min()
    start = np.full(len(uniq), np.inf)
    np.minimum.at(start, inv, (raw[:, 3] - t0) / 1e3)
    endv = np.where(raw[:, 11] > 0, (raw[:, 11] - t0) / 1e3, (raw[:, 3] - t0) / 1e3)
    end = np.full(len(uniq), -np.inf)
    np.maximum.at(end, inv, endv)
    tag = np.zeros(len(uniq), dtype=np.int64)
    tag[inv] = raw[:, 1] & 0xFF
    order = np.argsort(start, kind="stable")
    return [{"kernel": TAGS.get(int(tag[i]), str(int(tag[i]))), "start_first": float(start[i]),
             "start_last": float(start[i]), "end_last": float(end[i])} for i in order]
