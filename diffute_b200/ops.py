"""Thin Python wrappers over the C-ABI: build descriptors from torch tensors and launch on torch's current stream.

torch is used here only for device memory and streams.  Every function launches hand-written kernels from
libdiffute_b200.so and raises DfuError on failure; nothing falls back to torch arithmetic.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import EPI_F16, EPI_F32, EPI_GEGLU, Gemm, check, lib

PREC_FP16 = 1     # one tensor-core pass, fp16 operands (RN), fp32 accumulate
PREC_FP16X2 = 2   # three passes over (hi, lo) fp16 operand planes: ~fp32-accurate contraction


def planes_of(prec: int) -> int:
    return 2 if prec == PREC_FP16X2 else 1


def npass_of(prec: int) -> int:
    return 3 if prec == PREC_FP16X2 else 1


# ---------------------------------------------------------------------------------------------
# launch accounting (bench.py reports `gpu_launches`) and optional per-kernel CUDA-event profiling
# ---------------------------------------------------------------------------------------------
STATS = {"launches": 0}
PROFILE = None  # when a dict: kernel name -> list of (start_event, end_event, algorithmic_flops, algorithmic_bytes)


class _Prof:
    def __init__(self, name, launches=1, flops=0.0, nbytes=0.0):
        self.name, self.launches, self.flops, self.nbytes = name, launches, flops, nbytes

    def __enter__(self):
        STATS["launches"] += self.launches
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.setdefault(self.name, []).append((self.e0, self.e1, self.flops, self.nbytes))
        return False


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ---------------------------------------------------------------------------------------------
# one-time weight packing (load time, not on the sampling path): a job list executed by ONE dfu_pack_weights launch
# ---------------------------------------------------------------------------------------------
class Packer:
    """Collects pack jobs (diffusers state-dict tensor -> kernel layout) and runs them in one launch.

    Layouts: conv [O, I, kh, kw] -> fp16 [planes*O, kh*kw*I] with k = tap*I + i (tap-major, channels innermost);
    linear [N, K] -> fp16 [planes*N, K]; plane 0 = RN(x), plane 1 = RN(x - plane0); GEGLU projections interleave
    value / gate rows in blocks of 16; `into=` stacks several tensors into one matrix (q|k|v, the 22 time_emb_proj)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.jobs, self.prefix, self.total, self._keep = [], [], 0, []

    def _src(self, t: torch.Tensor) -> torch.Tensor:
        t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()  # a copy engine transfer, not a kernel
        self._keep.append(t)
        return t

    def _job(self, src, dst, rows, cin, taps, *, mode=0, row0=0, ld=None, geglu=False, planes=0, plane_stride=0,
             src2=None):
        j = _lib.PackJob()
        j.src, j.src2, j.dst = src.data_ptr(), (None if src2 is None else src2.data_ptr()), dst.data_ptr()
        j.plane_stride, j.rows, j.cin, j.taps, j.mode = plane_stride, rows, cin, taps, mode
        j.dst_row0, j.dst_ld, j.geglu, j.planes = row0, (ld if ld is not None else cin * taps), int(geglu), planes
        self.jobs.append(j)
        self.prefix.append(self.total)
        self.total += rows * cin * taps

    def weight16(self, w: torch.Tensor, planes: int, geglu: bool = False, into: Optional[torch.Tensor] = None,
                 row0: int = 0, total_rows: Optional[int] = None) -> torch.Tensor:
        """conv [O,I,kh,kw] or linear [N,K] -> fp16 [planes*rows, K] (K-major, tap-major).  `into` / `row0` /
        `total_rows` stack several sources into one packed matrix of `total_rows` rows per plane."""
        rows, cin = w.shape[0], w.shape[1]
        taps = w.numel() // (rows * cin)
        K = cin * taps
        tr = total_rows if total_rows is not None else rows
        dst = into if into is not None else torch.empty((planes * tr, K), dtype=torch.float16, device=self.device)
        self._job(self._src(w), dst, rows, cin, taps, row0=row0, ld=K, geglu=geglu, planes=planes, plane_stride=tr * K)
        return dst

    def f32(self, t: torch.Tensor, add: Optional[torch.Tensor] = None, geglu: bool = False,
            into: Optional[torch.Tensor] = None, row0: int = 0) -> torch.Tensor:
        """fp32 copy of a vector / matrix (optionally + `add`, GEGLU row interleave, stacked `into` a larger matrix)."""
        rows = t.shape[0]
        cin = t.numel() // rows
        dst = into if into is not None else torch.empty(tuple(t.shape), dtype=torch.float32, device=self.device)
        self._job(self._src(t), dst, rows, cin, 1, row0=row0, ld=cin, geglu=geglu, planes=0,
                  src2=None if add is None else self._src(add))
        return dst

    def small_in(self, w: torch.Tensor) -> torch.Tensor:
        """[Cout, Cin, k, k] -> fp32 [Cin*k*k, Cout] (consecutive output channels contiguous; dfu_conv_small_in)."""
        O, I = w.shape[0], w.shape[1]
        taps = w.numel() // (O * I)
        dst = torch.empty((I * taps, O), dtype=torch.float32, device=self.device)
        self._job(self._src(w), dst, O, I, taps, mode=1, ld=O)
        return dst

    def small_out(self, w: torch.Tensor) -> torch.Tensor:
        """[Cout, Cin, k, k] -> fp32 [Cout, k*k, Cin] (dfu_conv_small_out)."""
        O, I = w.shape[0], w.shape[1]
        taps = w.numel() // (O * I)
        dst = torch.empty((O, taps, I), dtype=torch.float32, device=self.device)
        self._job(self._src(w), dst, O, I, taps, ld=I * taps)
        return dst

    def run(self):
        if not self.jobs:
            return
        n = len(self.jobs)
        arr = (_lib.PackJob * n)(*self.jobs)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        jobs_dev = host.to(self.device)
        prefix_dev = torch.tensor(self.prefix, dtype=torch.int64).to(self.device)
        with _Prof("pack_weights", 1):
            check(lib().dfu_pack_weights(jobs_dev.data_ptr(), prefix_dev.data_ptr(), n, self.total, _stream()),
                  "dfu_pack_weights")
        torch.cuda.current_stream().synchronize()  # the job table and the fp32 sources may be freed after this
        self.jobs, self.prefix, self.total, self._keep = [], [], 0, []


# torch statements of the same layouts, for single tensors (operands built by tests / scripts, and the layout tests that
# check dfu_pack_weights against them).  Model loading goes through Packer.
def split_f16(x: torch.Tensor, planes: int) -> torch.Tensor:
    """fp32 [...] -> fp16 [planes, ...]: plane 0 = RN(x), plane 1 = RN(x - plane0)."""
    hi = x.to(torch.float16)
    if planes == 1:
        return hi.unsqueeze(0).contiguous()
    lo = (x - hi.to(torch.float32)).to(torch.float16)
    return torch.stack([hi, lo], 0).contiguous()


def pack_linear_weight(w: torch.Tensor, planes: int, geglu: bool = False) -> torch.Tensor:
    """[N, K] fp32 -> fp16 [planes*N, K] (K-major).  geglu: interleave value/gate rows in blocks of 16."""
    if geglu:
        w = geglu_interleave(w)
    return split_f16(w.contiguous(), planes).reshape(planes * w.shape[0], w.shape[1])


def pack_conv_weight(w: torch.Tensor, planes: int) -> torch.Tensor:
    """[O, I, kh, kw] fp32 -> fp16 [planes*O, kh*kw*I] with k = tap*I + i (tap-major, channels innermost)."""
    O, I, kh, kw = w.shape
    wk = w.permute(0, 2, 3, 1).reshape(O, kh * kw * I)
    return split_f16(wk.contiguous(), planes).reshape(planes * O, kh * kw * I)


def geglu_interleave(w: torch.Tensor) -> torch.Tensor:
    """rows [a_0..a_{n-1}, g_0..g_{n-1}] -> blocks of 32 rows: 16 value rows then the matching 16 gate rows."""
    n2 = w.shape[0]
    n = n2 // 2
    a, g = w[:n], w[n:]
    rest = w.shape[1:]
    a = a.reshape(n // 16, 16, *rest)
    g = g.reshape(n // 16, 16, *rest)
    return torch.cat([a, g], dim=1).reshape(n2, *rest).contiguous()


# ---------------------------------------------------------------------------------------------
# tap tables
# ---------------------------------------------------------------------------------------------
def taps_3x3_s1():
    return [(0, ky - 1, kx - 1) for ky in range(3) for kx in range(3)]


def taps_3x3_s2(batch: int, pad_lo: int):
    """stride-2 3x3 over a space-to-depth operand [4 parity planes][B][H/2][W/2][C].

    input row = 2*yo + ky - pad_lo  ->  parity plane (row & 1), plane row yo + floor((ky - pad_lo) / 2)."""
    out = []
    for ky in range(3):
        for kx in range(3):
            oy, ox = ky - pad_lo, kx - pad_lo
            py, px = oy & 1, ox & 1
            out.append(((py * 2 + px) * batch, (oy - py) // 2, (ox - px) // 2))
    return out


def _fill_taps(op, taps):
    for i, (dn, dy, dx) in enumerate(taps):
        op.tap_dn[i], op.tap_dy[i], op.tap_dx[i] = dn, dy, dx


# ---------------------------------------------------------------------------------------------
# GEMM launch
# ---------------------------------------------------------------------------------------------
class Workspace:
    """Grow-only device scratch (split-K partials).  One per engine; never allocated inside the C-ABI."""

    def __init__(self, nbytes: int = 0, device="cuda"):
        self.device = device
        self.generation = 0  # bumped on every reallocation: captured CUDA graphs holding the old address are stale
        self.buf = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)
        # grid-barrier words (GEMM: [0:2], GroupNorm: [2:4]); zeroed once, every kernel leaves them zero
        self.sync = torch.zeros(64, dtype=torch.int32, device=device)  # [0:2] GEMM barrier, [2:4]+[8:56] GroupNorm

    def ensure(self, nbytes: int):
        if self.buf.numel() < nbytes:
            self.buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=self.device)
            self.generation += 1
        return self.buf


_default_ws = None


def default_workspace() -> Workspace:
    global _default_ws
    if _default_ws is None:
        _default_ws = Workspace(64 << 20)
    return _default_ws


def matrix_operand(op, a16: torch.Tensor, w16: torch.Tensor, K: int, planes: int):
    """a16: fp16 [planes, M, K] view (row stride a16.stride(1), plane stride a16.stride(0));
    w16: fp16 [planes*N, K] packed weights, or a [planes, N, K] view of an activation used as the B operand."""
    op.a = a16.data_ptr()
    op.a_mode = 0
    op.a_ld = a16.stride(1)
    op.a_plane = a16.stride(0) // a16.stride(1) if planes > 1 else a16.shape[1]
    op.a_rows = op.a_plane * (planes - 1) + a16.shape[1]
    op.b = w16.data_ptr()
    op.b_static = 0 if w16.dim() == 3 else 1  # packed weights never change inside a step; an activation used as B does
    if w16.dim() == 3:
        op.b_ld = w16.stride(1)
        op.b_plane = w16.stride(0) // w16.stride(1) if planes > 1 else w16.shape[1]
        op.b_rows = op.b_plane * (planes - 1) + w16.shape[1]
    else:
        op.b_rows = w16.shape[0]
        op.b_ld = w16.stride(0)
        op.b_plane = w16.shape[0] // planes
    op.ntaps = 1
    op.k_per_tap = K
    _fill_taps(op, [(0, 0, 0)])


def image_operand(op, a16: torch.Tensor, w16: torch.Tensor, planes: int, taps, imgs_per_plane: int):
    """a16: fp16 [imgs_total, h, w, c] (planes / parity planes stacked along dim 0), w16: [planes*N, ntaps*c]."""
    imgs, h, w, c = a16.shape
    op.a = a16.data_ptr()
    op.a_mode = 1
    op.a_rows = imgs
    op.a_h, op.a_w, op.a_c = h, w, c
    op.a_plane = imgs_per_plane
    op.b = w16.data_ptr()
    op.b_static = 1
    op.b_rows = w16.shape[0]
    op.b_ld = w16.shape[1]
    op.b_plane = w16.shape[0] // planes
    op.ntaps = len(taps)
    op.k_per_tap = c
    _fill_taps(op, taps)


# ---------------------------------------------------------------------------------------------
# per-shape tiling table (measured on B200 inside the captured step by scripts/tune_insitu.py; the C cost model is the
# fallback for shapes the table does not hold)
# ---------------------------------------------------------------------------------------------
TUNE_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tuning_b200.json")
TUNE_TABLE = {}
TUNER = None  # optional hook: a callable (d, ws, key) -> (block_n, splits, stages) consulted for shapes not in the table
if os.path.exists(TUNE_PATH):
    try:
        with open(TUNE_PATH) as _f:
            TUNE_TABLE = {k: tuple(v) for k, v in json.load(_f).get("table", {}).items()}
    except Exception:  # a corrupt table only costs performance
        TUNE_TABLE = {}


def gemm_key(d: Gemm) -> str:
    kb = sum(d.g[i].ntaps * (d.g[i].k_per_tap // 64) for i in range(d.ngroups)) * d.npass
    return f"{d.conv}:{d.m}:{d.n}:{kb}:{d.epi}" + (f":b{d.batch}" if d.batch > 1 else "")


def launch_gemm(d: Gemm, ws: Optional[Workspace] = None):
    L = lib()
    ws = ws or default_workspace()
    d.sync_words = ws.sync.data_ptr()
    if d.block_n == 0 and d.splits == 0 and d.kernel == 0:
        key = gemm_key(d)
        if TUNER is not None and key not in TUNE_TABLE:
            TUNE_TABLE[key] = TUNER(d, ws, key)
        t = TUNE_TABLE.get(key)
        if t is not None:  # (block_n, splits, stages[, kernel]); 3-entry rows were tuned for the tile-per-CTA kernel
            d.block_n, d.splits, d.stages = t[:3]
            d.kernel = t[3] if len(t) > 3 else 1
    need = L.dfu_gemm_workspace(C.byref(d))
    if need:
        buf = ws.ensure(need)
        d.workspace = buf.data_ptr()
        d.workspace_bytes = buf.numel()
    k = sum(d.g[i].ntaps * d.g[i].k_per_tap for i in range(d.ngroups))
    with _Prof("gemm_conv" if d.conv else "gemm_linear", 1, 2.0 * d.m * d.n * k):
        check(L.dfu_gemm(C.byref(d), _stream()), "dfu_gemm")


def gemm_stats():
    """{launches, split, fused_second_stage, reduce_launches, pair_kernel, last_grid} since process start."""
    out = (C.c_int64 * 6)()
    lib().dfu_gemm_stats(out)
    return dict(zip(("launches", "split", "fused_second_stage", "reduce_launches", "pair_kernel", "last_grid"), out))


def set_epilogue(d: Gemm, *, out_f32=None, out_f16=None, bias=None, rowvec=None, rows_per_sample=0, residual=None,
                 alpha: float = 1.0, geglu: bool = False, gelu: bool = False):
    d.alpha = alpha
    d.act = 1 if gelu else 0
    d.bias = _ptr(bias)
    if rowvec is not None:
        d.rowvec = rowvec.data_ptr()
        d.rowvec_ld = rowvec.stride(0)
        d.rows_per_sample = rows_per_sample
    if residual is not None:
        d.residual = residual.data_ptr()
        d.ldr = residual.stride(-2)
    if out_f32 is not None:
        d.epi = EPI_F32
        d.out_f32 = out_f32.data_ptr()
        d.ldo = out_f32.stride(-2)
    else:
        d.epi = EPI_GEGLU if geglu else EPI_F16
        d.out_f16 = out_f16.data_ptr()
        d.ldh = out_f16.stride(-2)
        d.out_planes = out_f16.shape[0]
        d.out_plane_stride = out_f16.stride(0)


def _set_tune(d: Gemm, tune):
    """tune = (block_n, splits, stages[, kernel]); zeros = let the library / tuning table decide."""
    d.block_n, d.splits, d.stages = tune[:3]
    d.kernel = tune[3] if len(tune) > 3 else 0


def linear(a16: torch.Tensor, w16: torch.Tensor, n: int, prec: int, *, tune: Tuple[int, ...] = (0, 0, 0),
           ws: Optional[Workspace] = None, batch: int = 0, a_batch_rows: int = 0, b_batch_rows: int = 0, **epi):
    """a16 [planes, M, K] fp16 operand; w16 packed [planes*n, K]; epilogue kwargs as set_epilogue.

    batch > 1: M = batch * m_b rows hold `batch` independent products; product b multiplies A rows
    [b*a_batch_rows, b*a_batch_rows + m_b) with B rows [b*b_batch_rows, b*b_batch_rows + n) (B then is an activation
    view [planes, rows, K]) and writes output rows [b*m_b, (b+1)*m_b)."""
    planes = planes_of(prec)
    d = Gemm()
    d.m, d.n = a16.shape[1], n
    d.batch, d.a_batch_rows, d.b_batch_rows = batch, a_batch_rows, b_batch_rows
    d.ngroups, d.npass = 1, npass_of(prec)
    matrix_operand(d.g[0], a16, w16, a16.shape[2], planes)
    set_epilogue(d, **epi)
    _set_tune(d, tune)
    launch_gemm(d, ws)


def conv(a16: torch.Tensor, w16: torch.Tensor, n: int, prec: int, out_grid: Tuple[int, int, int], taps, *,
         imgs_per_plane: Optional[int] = None, shortcut: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
         tune: Tuple[int, ...] = (0, 0, 0), ws: Optional[Workspace] = None, **epi):
    """Implicit-GEMM conv.  a16 [imgs_total, h, w, c] fp16 operand (planes stacked on dim 0), out_grid = (B, H, W).

    shortcut = (raw16 [imgs_total, H, W, c2], wsc16 [planes*n, c2]) fuses the ResnetBlock2D 1x1 conv_shortcut as a
    second operand group accumulating into the same TMEM tile."""
    planes = planes_of(prec)
    B, H, W = out_grid
    d = Gemm()
    d.m, d.n = B * H * W, n
    d.ngroups, d.npass = (2 if shortcut is not None else 1), npass_of(prec)
    d.conv, d.B, d.H, d.W = 1, B, H, W
    ipp = imgs_per_plane if imgs_per_plane is not None else a16.shape[0] // planes
    image_operand(d.g[0], a16, w16, planes, taps, ipp)
    if shortcut is not None:
        raw16, wsc16 = shortcut
        image_operand(d.g[1], raw16, wsc16, planes, [(0, 0, 0)], raw16.shape[0] // planes)
    set_epilogue(d, **epi)
    _set_tune(d, tune)
    launch_gemm(d, ws)


# ---------------------------------------------------------------------------------------------
# normalisation / casts / small kernels
# ---------------------------------------------------------------------------------------------
def _new_operand(planes: int, shape, device) -> torch.Tensor:
    return torch.empty((planes, *shape), dtype=torch.float16, device=device)


def groupnorm(src0: torch.Tensor, gamma, beta, eps: float, silu: bool, prec: int, *, src1: Optional[torch.Tensor] = None,
              out16: Optional[torch.Tensor] = None, out32: Optional[torch.Tensor] = None,
              raw16: Optional[torch.Tensor] = None, groups: int = 32, ws: Optional[Workspace] = None):
    """src0 [B,H,W,C0] (+ src1 [B,H,W,C1] concatenated on channels) fp32 NHWC.  out16/raw16: [planes,B,H,W,C]."""
    B, H, W, C0 = src0.shape
    C1 = 0 if src1 is None else src1.shape[-1]
    L = lib()
    need = L.dfu_groupnorm_workspace(B, H * W, C0 + C1, groups)
    ws = ws or default_workspace()
    buf = ws.ensure(need)
    o = out16 if out16 is not None else raw16
    planes = o.shape[0] if o is not None else 1
    pstride = o.stride(0) if o is not None else 0
    with _Prof('groupnorm', 3, 0.0, float(B * H * W * (C0 + C1)) * (8 + 2 * planes * ((out16 is not None) + (raw16 is not None)) + 4 * (out32 is not None))):
        check(L.dfu_groupnorm(src0.data_ptr(), C0, _ptr(src1), C1, B, H * W, groups, gamma.data_ptr(), beta.data_ptr(),
                              eps, int(silu), _ptr(out16), planes, pstride, _ptr(out32), _ptr(raw16), buf.data_ptr(),
                              buf.numel(), ws.sync[2:].data_ptr(), _stream()), "dfu_groupnorm")  # counters at sync[10:]


def layernorm(x: torch.Tensor, gamma, beta, eps: float, out16: Optional[torch.Tensor] = None,
              out32: Optional[torch.Tensor] = None):
    """x [M, C] fp32 -> out16 [planes, M, C] fp16 operand planes and / or out32 [M, C] fp32."""
    M, C = x.shape
    planes = out16.shape[0] if out16 is not None else 0
    with _Prof('layernorm', 1, 0.0, float(M * C) * (4 + 2 * planes + (4 if out32 is not None else 0))):
        check(lib().dfu_layernorm(x.data_ptr(), M, C, gamma.data_ptr(), beta.data_ptr(), eps, _ptr(out16), max(planes, 1),
                                  out16.stride(0) if out16 is not None else 0, _ptr(out32), _stream()), "dfu_layernorm")


def patchify_f16(x: torch.Tensor, patch: int, out16: torch.Tensor):
    """x [B, C, H, W] fp32 NCHW -> out16 [planes, B*(1 + (H/P)*(W/P)), C*P*P] (row 0 of each sample = zeros, the CLS slot)."""
    B, Cc, H, W = x.shape
    with _Prof('patchify', 1):
        check(lib().dfu_patchify_f16(x.data_ptr(), B, Cc, H, W, patch, out16.data_ptr(), out16.shape[0], out16.stride(0),
                                     _stream()), "dfu_patchify_f16")


CAST_PLAIN, CAST_UP2X, CAST_S2D = 0, 1, 2


def glue_preprocess(image_u8: torch.Tensor, window, bbox, out_size: int = 512, lat_factor: int = 8):
    """image_u8 uint8 [h, w, 3] on the device; window = (x_s, y_s, cw, ch) clipped to the image; bbox = (x0, y0, x1, y1)
    inclusive.  -> (image [3,S,S], masked [3,S,S], mask [1,S,S], mask_lat [1,S/f,S/f]) fp32 (dfu_glue_preprocess)."""
    h, w, c = image_u8.shape
    assert c == 3 and image_u8.dtype == torch.uint8 and image_u8.is_contiguous()
    S, dev = out_size, image_u8.device
    img = torch.empty((3, S, S), dtype=torch.float32, device=dev)
    msk_img = torch.empty((3, S, S), dtype=torch.float32, device=dev)
    mask = torch.empty((1, S, S), dtype=torch.float32, device=dev)
    mask_lat = torch.empty((1, S // lat_factor, S // lat_factor), dtype=torch.float32, device=dev)
    x_s, y_s, cw, ch = window
    with _Prof("glue_preprocess", 1):
        check(lib().dfu_glue_preprocess(image_u8.data_ptr(), h, w, x_s, y_s, cw, ch, *(int(v) for v in bbox), S, lat_factor,
                                        img.data_ptr(), msk_img.data_ptr(), mask.data_ptr(), mask_lat.data_ptr(),
                                        _stream()), "dfu_glue_preprocess")
    return img, msk_img, mask, mask_lat


def glyph_preprocess(image_u8: torch.Tensor, out: torch.Tensor):
    """image_u8 uint8 [h, w, 3] on the device -> out fp32 [3, S, S]: ViTImageProcessor (PIL bilinear resize to S x S,
    rescale 1/255, normalise 0.5 / 0.5), dfu_glyph_preprocess."""
    h, w, c = image_u8.shape
    S = out.shape[-1]
    assert c == 3 and image_u8.dtype == torch.uint8 and image_u8.is_contiguous() and out.shape == (3, S, S)
    need = lib().dfu_glyph_preprocess_workspace(h, w, S)
    ws = torch.empty((need,), dtype=torch.uint8, device=image_u8.device)
    with _Prof("glyph_preprocess", 3):
        check(lib().dfu_glyph_preprocess(image_u8.data_ptr(), h, w, S, ws.data_ptr(), need, out.data_ptr(), _stream()),
              "dfu_glyph_preprocess")
    return out


def glue_composite(decoded: torch.Tensor, image_u8: torch.Tensor, origin, size, bbox, wrap: bool = False) -> torch.Tensor:
    """decoded fp32 [3, S, S] in [-1, 1]; image_u8 uint8 [h, w, 3]; origin = (x_s, y_s), size = (r_w, r_h) of the pasted
    region; bbox = numpy-slice corners (end exclusive).  -> uint8 [h, w, 3] (dfu_glue_composite)."""
    h, w, _ = image_u8.shape
    S = decoded.shape[-1]
    assert decoded.dtype == torch.float32 and decoded.is_contiguous() and decoded.shape == (3, S, S)
    out = torch.empty_like(image_u8)
    with _Prof("glue_composite", 1):
        check(lib().dfu_glue_composite(decoded.data_ptr(), S, image_u8.data_ptr(), h, w, origin[0], origin[1], size[0],
                                       size[1], *(int(v) for v in bbox), int(wrap), out.data_ptr(), _stream()),
              "dfu_glue_composite")
    return out


def cast_f16(x: torch.Tensor, mode: int, out16: torch.Tensor):
    """x [B,H,W,C] fp32 NHWC -> out16 [planes, ...] (plain: B,H,W,C; up2x: B,2H,2W,C; s2d: 4*B,H/2,W/2,C)."""
    B, H, W, Cc = x.shape
    with _Prof('cast_f16', 1):
        check(lib().dfu_cast_f16(x.data_ptr(), B, H, W, Cc, mode, out16.data_ptr(), out16.shape[0], out16.stride(0),
                                 _stream()), "dfu_cast_f16")


def timestep_embedding(t: torch.Tensor, dim: int, flip_sin_to_cos: bool, freq_shift: float, out: torch.Tensor):
    with _Prof('timestep_embedding', 1):
        check(lib().dfu_timestep_embedding(t.data_ptr(), t.shape[0], dim, int(flip_sin_to_cos), float(freq_shift),
                                           out.data_ptr(), _stream()), "dfu_timestep_embedding")


def gemv(x: torch.Tensor, W: torch.Tensor, bias, out: torch.Tensor, silu_in: bool = False, silu_out: bool = False):
    B, K = x.shape
    N = W.shape[0]
    with _Prof('gemv', 1):
        check(lib().dfu_gemv(x.data_ptr(), B, K, x.stride(0), W.data_ptr(), _ptr(bias), N, int(silu_in), int(silu_out),
                             out.data_ptr(), out.stride(0), _stream()), "dfu_gemv")


def pack_small_in_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, k, k] -> fp32 [Cin*k*k, Cout] (transposed so consecutive output channels are contiguous)."""
    return w.reshape(w.shape[0], -1).t().contiguous()


def conv_small_in(srcs: Sequence[torch.Tensor], wt: torch.Tensor, bias, out: torch.Tensor, batch: int,
                  pre_scale: float = 1.0, nhwc: bool = False):
    """srcs: up to 3 fp32 tensors gathered on channels, NCHW [B or 1, c, H, W] (or NHWC [B,H,W,c] with nhwc=True);
    wt = pack_small_in_weight(w) [Cin*k*k, Cout]; out NHWC."""
    cin = sum((s.shape[-1] if nhwc else s.shape[1]) for s in srcs)
    ksz = 3 if wt.shape[0] == 9 * cin else 1
    assert wt.shape[0] == cin * ksz * ksz, (wt.shape, cin)
    if nhwc:
        H, W = srcs[0].shape[1:3]
    else:
        H, W = srcs[0].shape[-2:]
    a = []
    for i in range(3):
        if i < len(srcs):
            s = srcs[i]
            a += [s.data_ptr(), s.shape[-1] if nhwc else s.shape[1], 0 if s.shape[0] == 1 and batch > 1 else s.stride(0)]
        else:
            a += [None, 0, 0]
    with _Prof('conv_small_in', 1):
        check(lib().dfu_conv_small_in(*a, int(nhwc), batch, H, W, ksz, wt.data_ptr(), _ptr(bias), wt.shape[1],
                                      pre_scale, out.data_ptr(), _stream()), "dfu_conv_small_in")


def pack_small_out_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, k, k] -> fp32 [Cout, k*k, Cin]."""
    O, I, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(O, kh * kw, I).contiguous()


def conv_small_out(x: torch.Tensor, wp: torch.Tensor, bias, out: Optional[torch.Tensor], *, w2=None, b2=None,
                   sample=None, prev=None, coef=None, seed=None):
    """x [B,H,W,Cin] fp32 NHWC; wp packed [Cout, k*k, Cin]; out NCHW.  Optional fused 1x1 (w2,b2) / scheduler step
    (coef = device {cx, ce}; with seed = device int32 [2]: {cx, ce, sigma, step} and the ancestral noise is added)."""
    B, H, W, Cin = x.shape
    Cout, kk, _ = wp.shape
    ksz = 3 if kk == 9 else 1
    with _Prof('conv_small_out', 1):
        check(lib().dfu_conv_small_out(x.data_ptr(), B, H, W, Cin, ksz, wp.data_ptr(), _ptr(bias), Cout, _ptr(w2),
                                       _ptr(b2), 0 if w2 is None else w2.shape[0], _ptr(out), _ptr(sample), _ptr(prev),
                                       _ptr(coef), _ptr(seed), _stream()), "dfu_conv_small_out")


def philox_normal(seed: int, step: int, n: int, device="cuda", bits: bool = False):
    """-> fp32 [n] N(0,1) of dfu_philox_normal (and, with bits=True, the raw Philox words int32 [n, 2])."""
    out = torch.empty((n,), dtype=torch.float32, device=device)
    raw = torch.empty((n, 2), dtype=torch.int32, device=device) if bits else None
    with _Prof('philox_normal', 1):
        check(lib().dfu_philox_normal(seed & 0xFFFFFFFFFFFFFFFF, step, n, out.data_ptr(), _ptr(raw), _stream()),
              "dfu_philox_normal")
    return (out, raw) if bits else out


def axpbypcz(x, e, n, a: float, b: float, c: float, y):
    with _Prof('axpbypcz', 1):
        check(lib().dfu_axpbypcz(x.data_ptr(), e.data_ptr(), _ptr(n), a, b, c, y.data_ptr(), x.numel(), _stream()),
              "dfu_axpbypcz")


def gaussian_sample(moments: torch.Tensor, eps: Optional[torch.Tensor], scale: float, z: torch.Tensor):
    B, C2, h, w = moments.shape
    with _Prof('gaussian_sample', 1):
        check(lib().dfu_gaussian_sample(moments.data_ptr(), _ptr(eps), B, C2 // 2, h * w, scale, z.data_ptr(), _stream()),
              "dfu_gaussian_sample")


def softmax_rows(s: torch.Tensor, scale: float, p16: torch.Tensor):
    rows, n = s.shape
    with _Prof('softmax_rows', 1):
        check(lib().dfu_softmax_rows(s.data_ptr(), rows, n, s.stride(0), scale, p16.data_ptr(), p16.stride(1),
                                     p16.shape[0], p16.stride(0), _stream()), "dfu_softmax_rows")


def transpose_f16(x16: torch.Tensor, out16: torch.Tensor):
    """x16 [planes, rows, cols] (row stride may exceed cols) -> out16 [planes, cols, rows] contiguous."""
    planes, rows, cols = x16.shape
    with _Prof('transpose_f16', 1):
        check(lib().dfu_transpose_f16(x16.data_ptr(), planes, rows, cols, x16.stride(1), x16.stride(0), out16.data_ptr(),
                                      out16.stride(0), _stream()), "dfu_transpose_f16")


def attention(q16: torch.Tensor, q_col0: int, k16: torch.Tensor, k_col0: int, v16: torch.Tensor, v_col0: int,
              B: int, heads: int, Nq: int, Nk: int, scale: float, out16: torch.Tensor, kv_splits: int = 0,
              ws: Optional[Workspace] = None):
    """q16 [planes, B*Nq, ldq], k16/v16 [planes, B*Nk, ld] fp16 operands (k16 and v16 share the plane stride);
    out16 [planes, B*Nq, heads*64].  kv_splits = 0 lets the library decide (split-KV + merge for under-filled grids)."""
    planes = q16.shape[0]
    L = lib()
    need = L.dfu_attention_workspace(B, heads, Nq, Nk, kv_splits)
    wptr, wbytes = None, 0
    if need:
        ws = ws or default_workspace()
        buf = ws.ensure(need)
        wptr, wbytes = buf.data_ptr(), buf.numel()
    with _Prof('attention', 2 if need else 1, 4.0 * B * heads * Nq * Nk * 64):
        check(L.dfu_attention(q16.data_ptr(), q16.stride(1), q_col0, q16.stride(0), k16.data_ptr(), k16.stride(1),
                              k_col0, v16.data_ptr(), v16.stride(1), v_col0, k16.stride(0), B, heads, Nq, Nk, planes,
                              scale, out16.data_ptr(), out16.stride(1), out16.stride(0), kv_splits, wptr, wbytes,
                              _stream()), "dfu_attention")


def scheduler_step(x, m, noise, a0, a1, p0, d0, d1, sn, clip: bool, y, x0_out=None):
    with _Prof('scheduler_step', 1):
        check(lib().dfu_scheduler_step(x.data_ptr(), m.data_ptr(), _ptr(noise), a0, a1, p0, d0, d1, sn, int(clip),
                                       y.data_ptr(), _ptr(x0_out), x.numel(), _stream()), "dfu_scheduler_step")


def axpby_rows(x, e, ca, cb, y):
    B = x.shape[0]
    with _Prof('axpby_rows', 1):
        check(lib().dfu_axpby_rows(x.data_ptr(), e.data_ptr(), ca.data_ptr(), cb.data_ptr(), y.data_ptr(), B,
                                   x.numel() // B, _stream()), "dfu_axpby_rows")
