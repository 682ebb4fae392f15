"""GPU parity of the tcgen05 contraction core (through the C-ABI) against fp64 torch on identical operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffute_b200 import ops as o
    return o


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


def _recombine(x16):
    return x16.double().sum(0)


# Operands are bit-identical on both sides, so what remains is the tensor core's fp32 accumulation: tcgen05
# adds each K=16 step into the TMEM accumulator with truncation, which shows up as a relative error that grows
# with the number of accumulation steps (measured on B200: ~6e-6 at 540 steps, ~2.6e-5 at 1260 steps), plus for
# FP16X2 the dropped lo*lo term (~2^-22).  The bound below is that accumulator behaviour, not operand rounding.
ACC_TOL = 6e-5


def _tol(prec, ops):
    return ACC_TOL


@pytest.mark.parametrize("prec", [1, 2])
@pytest.mark.parametrize("M,N,K,tune", [
    (4096, 320, 320, (0, 0, 0)),
    (4096, 320, 320, (160, 1, 3)),
    (1000, 640, 1024, (128, 1, 4)),      # ragged M
    (64, 1280, 1280, (0, 0, 0)),         # tiny M -> auto split-K
    (577, 640, 1024, (0, 0, 0)),         # glyph-token K/V projection shape
    (256, 1280, 5120, (256, 2, 4)),      # splits 2/4/8 reduce inside the kernel through a thread-block cluster
    (256, 1280, 1280, (128, 4, 3)),
    (64, 1280, 5120, (160, 8, 3)),
    (300, 640, 2560, (64, 8, 4)),        # ragged M + cluster reduce
    (64, 1280, 5120, (128, 16, 3)),      # 16-CTA clusters (non-portable size): half-full tile, 4 rows per CTA
    (200, 1280, 1280, (64, 16, 2)),      # ragged rows over 16 slices, some slices with a single k-block
    (256, 1280, 5120, (160, 6, 3)),      # other split counts go through the workspace + reduce kernel
    (128, 32, 64, (32, 1, 2)),           # single k-block
    # ---- persistent CTA-pair kernel (cta_group::2, two TMEM accumulators): tune[3] = 2 ----
    (4096, 320, 320, (160, 1, 0, 2)),    # 32 pair tiles
    (4096, 320, 320, (64, 1, 3, 2)),     # 80 pair tiles on 74 pairs: some pairs run two tiles (accumulator ping-pong)
    (1100, 640, 1024, (128, 1, 4, 2)),   # ragged M, odd number of 128-row tiles: the last pair's peer tile is empty
    (128, 32, 64, (32, 1, 2, 2)),        # one tile, one k-block, block_n = 32: half of the epilogue warps own no chunk
    (20000, 256, 512, (256, 1, 0, 2)),   # 79 tiles of 256 x 256: both accumulators = all 512 TMEM columns
    (40000, 320, 192, (80, 1, 0, 2)),    # 628 tiles: ~8.5 tiles per pair, 3 k-blocks each (ring wraps across tiles)
    (70000, 128, 128, (0, 0, 0)),        # automatic choice picks the pair kernel when its tiles alone fill the GPU
])
def test_linear_f32_epilogue(ops, prec, M, N, K, tune):
    planes = ops.planes_of(prec)
    a = _rand((M, K), 1)
    w = _rand((N, K), 2, K ** -0.5)
    bias = _rand((N,), 3)
    res = _rand((M, N), 4)
    a16 = ops.split_f16(a, planes)
    w16 = ops.pack_linear_weight(w, planes)
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.linear(a16, w16, N, prec, tune=tune, out_f32=out, bias=bias, residual=res, alpha=0.5)
    torch.cuda.synchronize()
    ref = 0.5 * (_recombine(a16) @ _recombine(w16.reshape(planes, N, K)).T) + bias.double() + res.double()
    if prec == 2:  # the kernel omits lo*lo
        ref = ref - 0.5 * (a16[1].double() @ w16.reshape(planes, N, K)[1].double().T)
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < _tol(prec, ops), err


@pytest.mark.parametrize("prec", [1, 2])
def test_linear_f16_and_geglu_epilogues(ops, prec):
    planes = ops.planes_of(prec)
    M, C = 1024, 640
    a = _rand((M, C), 5)
    a16 = ops.split_f16(a, planes)
    # plain f16 output (q/k/v projection), scaled
    w = _rand((3 * C, C), 6, C ** -0.5)
    w16 = ops.pack_linear_weight(w, planes)
    out16 = torch.zeros((planes, M, 3 * C), dtype=torch.float16, device="cuda")
    ops.linear(a16, w16, 3 * C, prec, out_f16=out16, alpha=0.125)
    ref = 0.125 * (_recombine(a16) @ w16.reshape(planes, 3 * C, C)[0].double().T)
    if prec == 2:
        ref = 0.125 * (_recombine(a16) @ _recombine(w16.reshape(planes, 3 * C, C)).T
                       - a16[1].double() @ w16.reshape(planes, 3 * C, C)[1].double().T)
    torch.cuda.synchronize()
    got = _recombine(out16)
    tol = 1e-3 if prec == 1 else ACC_TOL  # fp16 output rounding
    assert ((got - ref).abs().max() / ref.abs().max()).item() < tol
    out16.zero_()
    ops.linear(a16, w16, 3 * C, prec, tune=(192, 1, 0, 2), out_f16=out16, alpha=0.125)   # pair kernel, f16 planes out
    torch.cuda.synchronize()
    assert ((_recombine(out16) - ref).abs().max() / ref.abs().max()).item() < tol
    # GEGLU
    wg = _rand((8 * C, C), 7, C ** -0.5)
    bg = _rand((8 * C,), 8, 0.1)
    wg16 = ops.pack_linear_weight(wg, planes, geglu=True)
    bgi = ops.geglu_interleave(bg)
    wr = wg.to(torch.float16).double() if prec == 1 else ops.split_f16(wg, 2).double().sum(0)
    p = _recombine(a16) @ wr.T + bg.double()
    refg = p[:, :4 * C] * F.gelu(p[:, 4 * C:])
    tol = 1.5e-3 if prec == 1 else ACC_TOL
    # single pass, cluster split-K, workspace split-K, persistent CTA pairs (block_n 160: the pair splits the weight
    # tile at 80 rows, inside a 16/16 value/gate block; 256: full TMEM)
    for tune in ((0, 0, 0), (160, 2, 3), (128, 3, 3), (160, 1, 0, 2), (256, 1, 0, 2)):
        outg = torch.zeros((planes, M, 4 * C), dtype=torch.float16, device="cuda")
        ops.linear(a16, wg16, 8 * C, prec, tune=tune, out_f16=outg, bias=bgi, geglu=True)
        torch.cuda.synchronize()
        gotg = _recombine(outg)
        assert ((gotg - refg).abs().max() / refg.abs().max()).item() < tol, tune


def _conv_ref(x_nhwc, w, stride, pad):
    x = x_nhwc.permute(0, 3, 1, 2).double()
    if pad == "asym":
        x = F.pad(x, (0, 1, 0, 1))
        y = F.conv2d(x, w.double(), stride=2)
    else:
        y = F.conv2d(x, w.double(), stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize("prec", [1, 2])
@pytest.mark.parametrize("B,H,W,Cin,Cout", [
    (1, 64, 64, 320, 320),
    (2, 32, 32, 640, 640),
    (1, 16, 16, 1280, 1280),
    (1, 8, 8, 1280, 1280),      # one tile spans the whole (only) image: 64 valid rows
    (3, 8, 8, 640, 320),        # two images per tile, odd batch
    (1, 96, 96, 64, 64),        # 768-px latent: 96 of 128 tile rows used
    (1, 24, 24, 128, 128),
    (1, 256, 256, 128, 128),    # VAE-sized map: tile = half a row
])
def test_conv3x3_stride1(ops, prec, B, H, W, Cin, Cout):
    planes = ops.planes_of(prec)
    x = _rand((B, H, W, Cin), 11)
    w = _rand((Cout, Cin, 3, 3), 12, (9 * Cin) ** -0.5)
    bias = _rand((Cout,), 13)
    temb = _rand((B, Cout), 14)
    x16 = ops.split_f16(x, planes).reshape(planes * B, H, W, Cin)
    w16 = ops.pack_conv_weight(w, planes)
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda")
    xr = x16.reshape(planes, B, H, W, Cin).double().sum(0)
    wr = ops.split_f16(w, planes).double().sum(0)
    ref = _conv_ref(xr, wr, 1, 1) + bias.double() + temb.double()[:, None, None, :]
    bn = 160 if Cout % 160 == 0 else (128 if Cout % 128 == 0 else 64)
    # automatic tiling, a forced 4-way cluster split-K, the persistent CTA-pair kernel
    for tune in ((0, 0, 0), (bn, 4, 3), (bn, 1, 0, 2), (64, 1, 0, 2)):
        out.fill_(float("nan"))
        ops.conv(x16, w16, Cout, prec, (B, H, W), ops.taps_3x3_s1(), tune=tune, out_f32=out.view(B * H * W, Cout),
                 bias=bias, rowvec=temb, rows_per_sample=H * W)
        torch.cuda.synchronize()
        err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
        assert err < ACC_TOL, (tune, err)


@pytest.mark.parametrize("prec", [1, 2])
@pytest.mark.parametrize("mode", ["unet", "vae"])
def test_conv3x3_stride2_parity_planes(ops, prec, mode):
    planes = ops.planes_of(prec)
    B, H, W, C = 2, 32, 32, 128
    x = _rand((B, H, W, C), 21)
    w = _rand((C, C, 3, 3), 22, (9 * C) ** -0.5)
    # space-to-depth operand [planes][py][px][B][H/2][W/2][C]
    xs = ops.split_f16(x, planes)  # [planes,B,H,W,C]
    s2d = torch.stack([xs[:, :, py::2, px::2, :] for py in (0, 1) for px in (0, 1)], dim=1).contiguous()
    a16 = s2d.reshape(planes * 4 * B, H // 2, W // 2, C)
    w16 = ops.pack_conv_weight(w, planes)
    out = torch.full((B, H // 2, W // 2, C), float("nan"), device="cuda")
    pad_lo = 1 if mode == "unet" else 0
    ref = _conv_ref(xs.double().sum(0), ops.split_f16(w, planes).double().sum(0), 2, 1 if mode == "unet" else "asym")
    for tune in ((0, 0, 0), (128, 1, 0, 2)):
        out.fill_(float("nan"))
        ops.conv(a16, w16, C, prec, (B, H // 2, W // 2), ops.taps_3x3_s2(B, pad_lo), imgs_per_plane=4 * B, tune=tune,
                 out_f32=out.view(-1, C))
        torch.cuda.synchronize()
        err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
        assert err < ACC_TOL, (tune, err)


@pytest.mark.parametrize("prec", [1, 2])
def test_conv_with_fused_shortcut_and_residual(ops, prec):
    planes = ops.planes_of(prec)
    B, H, W, Cin, Cout = 1, 32, 32, 960, 640
    h = _rand((B, H, W, Cout), 31)     # conv2 input (already normalised/activated upstream)
    xraw = _rand((B, H, W, Cin), 32)   # block input for the 1x1 shortcut
    w2 = _rand((Cout, Cout, 3, 3), 33, (9 * Cout) ** -0.5)
    wsc = _rand((Cout, Cin, 1, 1), 34, Cin ** -0.5)
    bias = _rand((Cout,), 35)
    h16 = ops.split_f16(h, planes).reshape(planes * B, H, W, Cout)
    x16 = ops.split_f16(xraw, planes).reshape(planes * B, H, W, Cin)
    out = torch.full((B * H * W, Cout), float("nan"), device="cuda")
    rs = lambda t: ops.split_f16(t, planes).double().sum(0)
    ref = _conv_ref(rs(h), rs(w2), 1, 1) + _conv_ref(rs(xraw), rs(wsc), 1, 0) + bias.double()
    for tune in ((0, 0, 0), (160, 1, 0, 2)):   # second operand group through both kernels
        out.fill_(float("nan"))
        ops.conv(h16, ops.pack_conv_weight(w2, planes), Cout, prec, (B, H, W), ops.taps_3x3_s1(), tune=tune,
                 shortcut=(x16, ops.pack_conv_weight(wsc, planes)), out_f32=out, bias=bias)
        torch.cuda.synchronize()
        err = ((out.view(B, H, W, Cout).double() - ref).abs().max() / ref.abs().max()).item()
        assert err < ACC_TOL, (tune, err)


@pytest.mark.parametrize("prec", [1, 2])
@pytest.mark.parametrize("tune", [(0, 0, 0), (64, 2, 3), (128, 5, 2), (128, 1, 0, 2), (256, 1, 0, 2)])
def test_batched_products(ops, prec, tune):
    """DfuGemm.batch: independent products in one launch (per-sample Q K^T of the VAE attention): ragged rows per batch
    (200 = one full + one partial 128-row tile), B operand an activation view with its own batch stride, both kernels,
    cluster and workspace split-K."""
    planes = ops.planes_of(prec)
    nb, mb, n, K = 3, 200, 256, 192
    a = _rand((nb * mb, K), 41)
    b = _rand((nb * 300, K), 42, K ** -0.5)        # 300 rows per batch, only the first n = 256 are used
    res = _rand((nb * mb, n), 43)
    a16 = ops.split_f16(a, planes)
    b16 = ops.split_f16(b, planes)                  # [planes, nb*300, K] activation-style B
    out = torch.full((nb * mb, n), float("nan"), device="cuda")
    ops.linear(a16, b16, n, prec, tune=tune, batch=nb, a_batch_rows=mb, b_batch_rows=300, out_f32=out, residual=res)
    torch.cuda.synchronize()
    ar, br = _recombine(a16), _recombine(b16)
    ref = torch.cat([ar[i * mb:(i + 1) * mb] @ br[i * 300:i * 300 + n].T for i in range(nb)], 0) + res.double()
    if prec == 2:
        ref = ref - torch.cat([a16[1, i * mb:(i + 1) * mb].double() @ b16[1, i * 300:i * 300 + n].double().T
                               for i in range(nb)], 0)
    err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    assert err < ACC_TOL, err
