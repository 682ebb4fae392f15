"""Thin Python wrappers over the C-ABI: build descriptors from torch tensors and launch on torch's current stream.

torch is used here only for device memory and streams.  Every function launches hand-written kernels from
libdiffute_b200.so and raises DfuError on failure; nothing falls back to torch arithmetic.
"""
from __future__ import annotations

import ctypes as C
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


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ---------------------------------------------------------------------------------------------
# one-time weight packing (load time, not on the sampling path)
# ---------------------------------------------------------------------------------------------
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


def geglu_interleave(w: torch.Tensor) -> torch.Tensor:
    """rows [a_0..a_{n-1}, g_0..g_{n-1}] -> blocks of 32 rows: 16 value rows then the matching 16 gate rows."""
    n2 = w.shape[0]
    n = n2 // 2
    a, g = w[:n], w[n:]
    rest = w.shape[1:]
    a = a.reshape(n // 16, 16, *rest)
    g = g.reshape(n // 16, 16, *rest)
    return torch.cat([a, g], dim=1).reshape(n2, *rest).contiguous()


def pack_conv_weight(w: torch.Tensor, planes: int) -> torch.Tensor:
    """[O, I, kh, kw] fp32 -> fp16 [planes*O, kh*kw*I] with k = tap*I + i (tap-major, channels innermost)."""
    O, I, kh, kw = w.shape
    wk = w.permute(0, 2, 3, 1).reshape(O, kh * kw * I)
    return split_f16(wk.contiguous(), planes).reshape(planes * O, kh * kw * I)


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
        self.buf = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)

    def ensure(self, nbytes: int):
        if self.buf.numel() < nbytes:
            self.buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=self.device)
        return self.buf


_default_ws = None


def default_workspace() -> Workspace:
    global _default_ws
    if _default_ws is None:
        _default_ws = Workspace(64 << 20)
    return _default_ws


def matrix_operand(op, a16: torch.Tensor, w16: torch.Tensor, K: int, planes: int):
    """a16: fp16 [planes, M, K(ld)], w16: fp16 [planes*N, K]."""
    rows = a16.shape[0] * a16.shape[1]
    op.a = a16.data_ptr()
    op.a_mode = 0
    op.a_rows = rows
    op.a_ld = a16.stride(1)
    op.a_plane = a16.shape[1]
    op.b = w16.data_ptr()
    op.b_rows = w16.shape[0]
    op.b_ld = w16.shape[1]
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
    op.b_rows = w16.shape[0]
    op.b_ld = w16.shape[1]
    op.b_plane = w16.shape[0] // planes
    op.ntaps = len(taps)
    op.k_per_tap = c
    _fill_taps(op, taps)


def launch_gemm(d: Gemm, ws: Optional[Workspace] = None):
    L = lib()
    need = L.dfu_gemm_workspace(C.byref(d))
    if need:
        ws = ws or default_workspace()
        buf = ws.ensure(need)
        d.workspace = buf.data_ptr()
        d.workspace_bytes = buf.numel()
    check(L.dfu_gemm(C.byref(d), _stream()), "dfu_gemm")


def set_epilogue(d: Gemm, *, out_f32=None, out_f16=None, bias=None, rowvec=None, rows_per_sample=0, residual=None,
                 alpha: float = 1.0, geglu: bool = False):
    d.alpha = alpha
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


def linear(a16: torch.Tensor, w16: torch.Tensor, n: int, prec: int, *, tune: Tuple[int, int, int] = (0, 0, 0),
           ws: Optional[Workspace] = None, **epi):
    """a16 [planes, M, K] fp16 operand; w16 packed [planes*n, K]; epilogue kwargs as set_epilogue."""
    planes = planes_of(prec)
    d = Gemm()
    d.m, d.n = a16.shape[1], n
    d.ngroups, d.npass = 1, npass_of(prec)
    matrix_operand(d.g[0], a16, w16, a16.shape[2], planes)
    set_epilogue(d, **epi)
    d.block_n, d.splits, d.stages = tune
    launch_gemm(d, ws)


def conv(a16: torch.Tensor, w16: torch.Tensor, n: int, prec: int, out_grid: Tuple[int, int, int], taps, *,
         imgs_per_plane: Optional[int] = None, shortcut: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
         tune: Tuple[int, int, int] = (0, 0, 0), ws: Optional[Workspace] = None, **epi):
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
    d.block_n, d.splits, d.stages = tune
    launch_gemm(d, ws)
