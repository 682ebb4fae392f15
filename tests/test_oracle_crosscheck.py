"""CPU: pin `oracle/` to something it did not generate.

The reference holds no golden vectors for the UNet / VAE and diffusers cannot be installed here, so the
oracle is checked against `tests/refmath.py` — a second restatement of the same published architecture in a different
form (numpy float64, flat functional walk over the state dict, explicit im2col / softmax / normalisation).  The checks
go primitive by primitive, block by block and end to end, on the same seeded synthetic weights the GPU tests use.
Bar: the oracle computes in fp32, refmath in fp64 — agreement to a few 1e-5 relative (fp32 accumulation noise).

When a real diffusers is importable (e.g. the driver provides one under baseline/_ref), the last test activates and
compares the oracle with diffusers' own UNet2DConditionModel / AutoencoderKL on the same state dict.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refmath as R  # noqa: E402
from diffute_b200 import arch, synthetic  # noqa: E402
from oracle import DDIMOracle, UNetOracle, VAEOracle  # noqa: E402
from oracle import unet as ou  # noqa: E402


def rel(a, b):
    a = a.detach().double().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().double().numpy() if torch.is_tensor(b) else np.asarray(b)
    return float(np.abs(a - b).max() / np.abs(b).max())


def nhwc(t):
    return R.f64(t).transpose(0, 2, 3, 1)


def nchw(a):
    return a.transpose(0, 3, 1, 2)


@pytest.fixture(scope="module")
def unet():
    sd = synthetic.make_state_dict(arch.unet_param_shapes())
    u = UNetOracle()
    u.load_state_dict(sd)
    return sd, u


@pytest.fixture(scope="module")
def vae():
    sd = synthetic.make_state_dict(arch.vae_param_shapes())
    v = VAEOracle()
    v.load_state_dict(sd)
    return sd, v


def test_primitives():
    g = torch.Generator().manual_seed(0)
    x = torch.randn((2, 64, 9, 7), generator=g)
    w = torch.randn((48, 64, 3, 3), generator=g) * 0.05
    b = torch.randn((48,), generator=g)
    for stride, pad_t, pad_r in [(1, 1, (1, 1, 1, 1)), (2, 1, (1, 1, 1, 1))]:
        assert rel(F.conv2d(x, w, b, stride=stride, padding=pad_t),
                   nchw(R.conv2d(nhwc(x), R.f64(w), R.f64(b), stride, pad_r))) < 1e-5
    # VAE encoder downsample: F.pad(0,1,0,1) then stride 2, no padding
    xe = torch.randn((1, 64, 8, 8), generator=g)
    assert rel(F.conv2d(F.pad(xe, (0, 1, 0, 1)), w, b, stride=2),
               nchw(R.conv2d(nhwc(xe), R.f64(w), R.f64(b), 2, (0, 1, 0, 1)))) < 1e-5
    w1 = torch.randn((48, 64, 1, 1), generator=g)
    assert rel(F.conv2d(x, w1, None), nchw(R.conv2d(nhwc(x), R.f64(w1), None, 1, (0, 0, 0, 0)))) < 1e-5
    ga, be = torch.randn((64,), generator=g), torch.randn((64,), generator=g)
    assert rel(F.group_norm(x, 32, ga, be, 1e-5), nchw(R.group_norm(nhwc(x), R.f64(ga), R.f64(be), 32, 1e-5))) < 1e-5
    t = torch.randn((3, 11, 64), generator=g) * 3
    assert rel(F.layer_norm(t, (64,), ga, be, 1e-5), R.layer_norm(R.f64(t), R.f64(ga), R.f64(be))) < 1e-5
    assert rel(F.gelu(t), R.gelu_erf(R.f64(t))) < 1e-6 and rel(F.silu(t), R.silu(R.f64(t))) < 1e-6
    assert rel(F.interpolate(x, scale_factor=2.0, mode="nearest"), nchw(R.upsample_nearest2x(nhwc(x)))) == 0.0
    q, k, v = (torch.randn((2, n, 128), generator=g) for n in (10, 13, 13))
    assert rel(ou.attention_core(q, k, v, 2, 64 ** -0.5), R.attention(R.f64(q), R.f64(k), R.f64(v), 2)) < 1e-5
    ts = torch.tensor([981.0, 1.0, 500.0])
    assert rel(ou.timestep_sincos(ts, 320, True, 0.0), R.timestep_sincos(ts.numpy(), 320)) < 1e-4  # fp32 sin of ~1e3


def test_resnet_and_transformer_blocks(unet):
    sd, u = unet
    P = R.Net(sd)
    g = torch.Generator().manual_seed(1)
    temb = torch.randn((2, 1280), generator=g)
    ctx = torch.randn((2, 21, 1024), generator=g)
    # a resnet with a 1x1 shortcut (320 -> 640), one without, and an up-block resnet fed by a channel concat
    for key, mod, cin in [("down_blocks.1.resnets.0", u.down_blocks[1].resnets[0], 320),
                          ("down_blocks.0.resnets.1", u.down_blocks[0].resnets[1], 320),
                          ("up_blocks.3.resnets.0", u.up_blocks[3].resnets[0], 960)]:
        x = torch.randn((2, cin, 6, 6), generator=g)
        e = rel(mod(x, temb), nchw(R.resnet(P, key, nhwc(x), R.f64(temb))))
        print(f"{key}: {e:.2e}")
        assert e < 2e-5
    for key, mod, c, heads in [("down_blocks.0.attentions.0", u.down_blocks[0].attentions[0], 320, 5),
                               ("mid_block.attentions.0", u.mid_block.attentions[0], 1280, 20)]:
        x = torch.randn((2, c, 4, 4), generator=g)
        e = rel(mod(x, ctx), nchw(R.transformer(P, key, nhwc(x), R.f64(ctx), heads)))
        print(f"{key}: {e:.2e}")
        assert e < 2e-5


def test_unet_forward_end_to_end(unet):
    """Whole UNet2DConditionModel.forward (app.ipynb:814) at an 8x8 latent, batch 2, per-sample timesteps."""
    sd, u = unet
    g = torch.Generator().manual_seed(2)
    sample = torch.randn((2, 9, 8, 8), generator=g)
    ctx = torch.randn((2, 37, 1024), generator=g)
    t = torch.tensor([981, 21])
    ref = R.unet_forward(sd, sample, t.numpy(), ctx)
    e = rel(u(sample, t, ctx).sample, ref)
    print(f"UNet forward, oracle (fp32) vs refmath (fp64): maxrel {e:.3e}")
    assert e < 5e-5


def test_vae_end_to_end(vae):
    sd, v = vae
    g = torch.Generator().manual_seed(3)
    x = torch.rand((1, 3, 32, 32), generator=g) * 2 - 1
    mom = v.encode(x).latent_dist.parameters
    e_enc = rel(mom, R.vae_encode_moments(sd, x))
    z = torch.randn((1, 4, 4, 4), generator=g)
    e_dec = rel(v.decode(z).sample, R.vae_decode(sd, z))
    print(f"VAE oracle vs refmath: encode moments {e_enc:.3e}, decode {e_dec:.3e}")
    assert e_enc < 5e-5 and e_dec < 5e-5


def test_ddim_against_the_papers_update():
    s = DDIMOracle()
    s.set_timesteps(50)
    assert [int(t) for t in s.timesteps] == R.ddim_timesteps(50)
    g = torch.Generator().manual_seed(4)
    x, eps = torch.randn((1, 4, 8, 8), generator=g), torch.randn((1, 4, 8, 8), generator=g)
    for t in (981, 501, 21, 1):
        assert rel(s.step(eps, t, x).prev_sample, R.ddim_step(R.f64(x), R.f64(eps), t, 50)) < 1e-5


def test_against_real_diffusers_when_available(unet, vae):
    """Auto-activating pin: needs the real third-party package the reference calls (not installable offline here).
    Command once it exists:  python -m pytest tests/test_oracle_crosscheck.py -k real_diffusers"""
    diffusers = pytest.importorskip("diffusers")
    usd, u = unet
    vsd, v = vae
    g = torch.Generator().manual_seed(5)
    du = diffusers.UNet2DConditionModel(**{k: (list(x) if isinstance(x, tuple) else x)
                                           for k, x in arch.SD2_INPAINT_UNET_CONFIG.items()})
    du.load_state_dict(usd)
    du.eval()
    sample, ctx = torch.randn((1, 9, 16, 16), generator=g), torch.randn((1, 577, 1024), generator=g)
    with torch.no_grad():
        assert rel(u(sample, 981, ctx).sample, du(sample, 981, ctx).sample) < 1e-4
    cfg = dict(arch.SD2_VAE_CONFIG)
    dv = diffusers.AutoencoderKL(in_channels=3, out_channels=3, block_out_channels=list(cfg["block_out_channels"]),
                                 layers_per_block=2, latent_channels=4, norm_num_groups=32,
                                 down_block_types=["DownEncoderBlock2D"] * 4, up_block_types=["UpDecoderBlock2D"] * 4)
    dv.load_state_dict(vsd)
    dv.eval()
    x = torch.rand((1, 3, 64, 64), generator=g) * 2 - 1
    with torch.no_grad():
        assert rel(v.encode(x).latent_dist.parameters, dv.encode(x).latent_dist.parameters) < 1e-4
        z = torch.randn((1, 4, 8, 8), generator=g)
        assert rel(v.decode(z).sample, dv.decode(z).sample) < 1e-4
    ds = diffusers.DDIMScheduler(**arch.SD2_SCHEDULER_CONFIG)
    ds.set_timesteps(50)
    o = DDIMOracle()
    o.set_timesteps(50)
    assert [int(t) for t in ds.timesteps] == [int(t) for t in o.timesteps]
    eps, lat = torch.randn((1, 4, 8, 8), generator=g), torch.randn((1, 4, 8, 8), generator=g)
    assert rel(o.step(eps, 501, lat).prev_sample, ds.step(eps, 501, lat).prev_sample) < 1e-5
