"""GPU parity of the memory-bound / small kernels (through the C-ABI) against fp64 torch."""
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


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


@pytest.mark.parametrize("planes", [1, 2])
@pytest.mark.parametrize("B,H,W,C0,C1,silu,eps", [
    (1, 64, 64, 320, 0, True, 1e-5),
    (2, 32, 32, 640, 320, True, 1e-5),     # concat, group boundary straddles the two sources (960/32 = 30)
    (1, 16, 16, 1280, 640, True, 1e-5),    # 1920 channels
    (1, 8, 8, 1280, 1280, True, 1e-5),     # 2560 channels
    (2, 16, 16, 1280, 0, False, 1e-6),     # Transformer2DModel.norm
    (1, 256, 256, 128, 0, True, 1e-6),     # VAE-sized
    # slabs staged in shared memory (gn_cluster_smem_kernel): batched UNet levels and mid-sized VAE maps
    (8, 64, 64, 320, 0, True, 1e-5),
    (8, 32, 32, 640, 320, True, 1e-5),     # two-source concat through the cp.async path
    (5, 64, 64, 640, 320, True, 1e-5),     # 960 channels at level 0 (odd batch)
    (2, 128, 128, 512, 0, True, 1e-6),     # VAE decoder map: used to take the two-launch path
    (3, 96, 96, 320, 0, False, 1e-6),      # 768-px latent, no SiLU
])
def test_groupnorm(ops, planes, B, H, W, C0, C1, silu, eps):
    C = C0 + C1
    x0 = _rand((B, H, W, C0), 1) * 3 + 0.7
    x1 = _rand((B, H, W, C1), 2) * 0.5 - 1.0 if C1 else None
    gamma = 1 + 0.1 * _rand((C,), 3)
    beta = 0.1 * _rand((C,), 4)
    out16 = torch.zeros((planes, B, H, W, C), dtype=torch.float16, device="cuda")
    raw16 = torch.zeros_like(out16)
    out32 = torch.zeros((B, H, W, C), device="cuda")
    ops.groupnorm(x0, gamma, beta, eps, silu, planes, src1=x1, out16=out16, out32=out32, raw16=raw16)
    torch.cuda.synchronize()
    x = x0 if x1 is None else torch.cat([x0, x1], -1)
    ref = F.group_norm(x.permute(0, 3, 1, 2).double(), 32, gamma.double(), beta.double(), eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 3, 1)
    assert _rel(out32, ref) < 5e-6
    tol16 = 6e-4 if planes == 1 else 2e-6
    assert _rel(out16.double().sum(0), ref) < tol16
    assert _rel(raw16.double().sum(0), x) < tol16


@pytest.mark.parametrize("planes", [1, 2])
@pytest.mark.parametrize("M,C", [(4096, 320), (1024, 640), (77, 1280), (32768, 320), (9473, 640), (1154, 1024)])
def test_layernorm(ops, planes, M, C):
    x = _rand((M, C), 5) * 2 + 0.3
    g = 1 + 0.1 * _rand((C,), 6)
    b = 0.1 * _rand((C,), 7)
    out16 = torch.zeros((planes, M, C), dtype=torch.float16, device="cuda")
    ops.layernorm(x, g, b, 1e-5, out16)
    torch.cuda.synchronize()
    ref = F.layer_norm(x.double(), (C,), g.double(), b.double(), 1e-5)
    assert _rel(out16.double().sum(0), ref) < (6e-4 if planes == 1 else 2e-6)


@pytest.mark.parametrize("planes", [1, 2])
def test_casts(ops, planes):
    B, H, W, C = 2, 16, 16, 64
    x = _rand((B, H, W, C), 8)
    tol = 6e-4 if planes == 1 else 2e-6
    o = torch.zeros((planes, B, H, W, C), dtype=torch.float16, device="cuda")
    ops.cast_f16(x, ops.CAST_PLAIN, o)
    assert _rel(o.double().sum(0), x) < tol
    u = torch.zeros((planes, B, 2 * H, 2 * W, C), dtype=torch.float16, device="cuda")
    ops.cast_f16(x, ops.CAST_UP2X, u)
    ref = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert _rel(u.double().sum(0), ref) < tol
    s = torch.zeros((planes, 4 * B, H // 2, W // 2, C), dtype=torch.float16, device="cuda")
    ops.cast_f16(x, ops.CAST_S2D, s)
    ref = torch.stack([x[:, py::2, px::2] for py in (0, 1) for px in (0, 1)], 0).reshape(4 * B, H // 2, W // 2, C)
    assert _rel(s.double().sum(0), ref) < tol


def test_timestep_embedding_and_gemv(ops):
    from oracle.unet import timestep_sincos
    t = torch.tensor([981.0, 1.0, 500.0], device="cuda")
    out = torch.zeros((3, 320), device="cuda")
    ops.timestep_embedding(t, 320, True, 0.0, out)
    ref = timestep_sincos(t.cpu(), 320, True, 0.0)
    assert (out.cpu() - ref).abs().max().item() < 2e-4  # fp32 sin/cos of arguments up to ~1e3
    x = _rand((3, 1280), 9)
    W = _rand((2048, 1280), 10, 1280 ** -0.5)
    b = _rand((2048,), 11)
    o = torch.zeros((3, 2048), device="cuda")
    ops.gemv(x, W, b, o, silu_in=True, silu_out=True)
    ref = F.silu(F.linear(F.silu(x.double()), W.double(), b.double()))
    assert _rel(o, ref) < 1e-5


def test_conv_small_in_out(ops):
    B, H, W = 2, 32, 32
    lat, mask, ml = _rand((B, 4, H, W), 12), (_rand((1, 1, H, W), 13) > 0).float(), _rand((B, 4, H, W), 14)
    w = _rand((320, 9, 3, 3), 15, 81 ** -0.5)
    bias = _rand((320,), 16)
    out = torch.zeros((B, H, W, 320), device="cuda")
    ops.conv_small_in([lat, mask, ml], ops.pack_small_in_weight(w), bias, out, B)
    x = torch.cat([lat, mask.expand(B, -1, -1, -1), ml], 1)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    assert _rel(out, ref) < 1e-5
    # 1x1 with pre-scale (post_quant_conv on latents / scaling_factor)
    w1 = _rand((4, 4, 1, 1), 17)
    o1 = torch.zeros((B, H, W, 4), device="cuda")
    ops.conv_small_in([lat], ops.pack_small_in_weight(w1), None, o1, B, pre_scale=1 / 0.18215)
    ref = F.conv2d(lat.double() / 0.18215, w1.double()).permute(0, 2, 3, 1)
    assert _rel(o1, ref) < 1e-5
    # few-output conv with fused scheduler step and trailing 1x1
    xin = _rand((B, H, W, 320), 18)
    wo = _rand((4, 320, 3, 3), 19, 2880 ** -0.5)
    bo = _rand((4,), 20)
    eps_out = torch.zeros((B, 4, H, W), device="cuda")
    prev = torch.zeros_like(eps_out)
    coef = torch.tensor([1.01, -0.03], device="cuda")
    ops.conv_small_out(xin, ops.pack_small_out_weight(wo), bo, eps_out, sample=lat, prev=prev, coef=coef)
    ref = F.conv2d(xin.permute(0, 3, 1, 2).double(), wo.double(), bo.double(), padding=1)
    assert _rel(eps_out, ref) < 1e-5
    assert _rel(prev, 1.01 * lat.double() + (-0.03) * ref) < 1e-5
    w2 = _rand((8, 8, 1, 1), 21)
    b2 = _rand((8,), 22)
    wo8 = _rand((8, 320, 3, 3), 23, 2880 ** -0.5)
    mom = torch.zeros((B, 8, H, W), device="cuda")
    ops.conv_small_out(xin, ops.pack_small_out_weight(wo8), None, mom, w2=w2.reshape(8, 8).contiguous(), b2=b2)
    ref = F.conv2d(F.conv2d(xin.permute(0, 3, 1, 2).double(), wo8.double(), padding=1), w2.double(), b2.double())
    assert _rel(mom, ref) < 1e-5


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 512, 512, 32, 3), (2, 300, 200, 8, 8), (1, 256, 256, 64, 1), (3, 40, 40, 32, 3)])
def test_conv1x1_few_channels(ops, B, H, W, Cin, Cout):
    """1x1 conv with few input channels (the selection conv behind the VAE decoder's tensor-core conv_out, the quant
    convs): large maps take the pixel-per-thread kernel, small ones the warp-per-pixel kernel; same result."""
    x = _rand((B, H, W, Cin), 31)
    w = _rand((Cout, Cin, 1, 1), 32, Cin ** -0.5)
    b = _rand((Cout,), 33)
    out = torch.full((B, Cout, H, W), float("nan"), device="cuda")
    ops.conv_small_out(x, ops.pack_small_out_weight(w), b, out)
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double())
    assert _rel(out, ref) < 1e-5


def test_elementwise_and_softmax(ops):
    x, e, n = _rand((2, 4, 64, 64), 24), _rand((2, 4, 64, 64), 25), _rand((2, 4, 64, 64), 26)
    y = torch.zeros_like(x)
    ops.axpbypcz(x, e, n, 1.02, -0.05, 0.3, y)
    assert _rel(y, 1.02 * x.double() - 0.05 * e.double() + 0.3 * n.double()) < 1e-6
    mom = _rand((2, 8, 16, 16), 27) * 3
    eps = _rand((2, 4, 16, 16), 28)
    z = torch.zeros((2, 4, 16, 16), device="cuda")
    ops.gaussian_sample(mom, eps, 0.18215, z)
    mean, lv = mom.double().chunk(2, 1)
    ref = (mean + torch.exp(0.5 * lv.clamp(-30, 20)) * eps.double()) * 0.18215
    assert _rel(z, ref) < 1e-5
    ops.gaussian_sample(mom, None, 1.0, z)
    assert _rel(z, mean) < 1e-7
    for planes in (1, 2):
        s = _rand((300, 1000), 29) * 4
        p16 = torch.zeros((planes, 300, 1000), dtype=torch.float16, device="cuda")
        ops.softmax_rows(s, 0.5, p16)
        ref = torch.softmax(s.double() * 0.5, -1)
        assert (p16.double().sum(0) - ref).abs().max().item() < (3e-4 if planes == 1 else 1e-6)
        t = torch.zeros((planes, 1000, 300), dtype=torch.float16, device="cuda")
        ops.transpose_f16(p16, t)
        assert torch.equal(t, p16.transpose(1, 2))


def test_groupnorm_two_launch_path_subprocess():
    """The statistics + apply pair (used for maps too large for the cluster kernel, e.g. VAE 512x512) on the standard
    shapes: DFU_GN_CLUSTER=0 is read once per process, so the same test runs again in a child process."""
    import os, subprocess, sys
    if os.environ.get("DFU_GN_CLUSTER") == "0":
        pytest.skip("already the child")
    env = dict(os.environ, DFU_GN_CLUSTER="0")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", __file__, "-k", "test_groupnorm and not subprocess"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_pack_weights_kernel_matches_torch_layouts():
    """dfu_pack_weights (one launch for a whole job list) against the torch statements of every layout it produces."""
    from diffute_b200 import ops
    g = torch.Generator().manual_seed(9)
    r = lambda *s: torch.randn(s, generator=g)
    wc, wl, wg, bg = r(96, 64, 3, 3), r(128, 192), r(256, 64), r(256)
    q, k, v = r(64, 128), r(64, 128), r(64, 128)
    b1, b2, wi, wo = r(96), r(96), r(320, 9, 3, 3), r(4, 320, 3, 3)
    for planes in (1, 2):
        pk = ops.Packer("cuda")
        c16 = pk.weight16(wc, planes)
        l16 = pk.weight16(wl, planes)
        g16 = pk.weight16(wg, planes, geglu=True)
        gb = pk.f32(bg.reshape(-1, 1), geglu=True).reshape(-1)
        qkv = torch.empty((planes * 192, 128), dtype=torch.float16, device="cuda")
        for i, t in enumerate((q, k, v)):
            pk.weight16(t, planes, into=qkv, row0=64 * i, total_rows=192)
        bsum = pk.f32(b1, add=b2)
        si, so = pk.small_in(wi), pk.small_out(wo)
        stack = torch.empty((192, 128), dtype=torch.float32, device="cuda")
        pk.f32(q, into=stack, row0=0)
        pk.f32(k, into=stack, row0=64)
        pk.f32(v, into=stack, row0=128)
        pk.run()
        assert torch.equal(c16.cpu(), ops.pack_conv_weight(wc, planes))
        assert torch.equal(l16.cpu(), ops.pack_linear_weight(wl, planes))
        assert torch.equal(g16.cpu(), ops.pack_linear_weight(wg, planes, geglu=True))
        assert torch.equal(gb.cpu(), ops.geglu_interleave(bg))
        assert torch.equal(qkv.cpu(), ops.pack_linear_weight(torch.cat([q, k, v], 0), planes))
        assert torch.equal(bsum.cpu(), b1 + b2)
        assert torch.equal(si.cpu(), ops.pack_small_in_weight(wi))
        assert torch.equal(so.cpu(), ops.pack_small_out_weight(wo))
        assert torch.equal(stack.cpu(), torch.cat([q, k, v], 0))
