"""CPU: structural pins of the oracle (parameter identities, key inventory) and agreement with the committed golden
vectors (so the oracle cannot drift unnoticed)."""
import json
import os

import pytest
import torch

from diffute_b200 import arch, synthetic
from oracle import DDIMOracle, UNetOracle, VAEOracle, sample_loop

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.json")


@pytest.fixture(scope="module")
def models():
    u, v = UNetOracle(), VAEOracle()
    u.load_state_dict(synthetic.make_state_dict(arch.unet_param_shapes()))
    v.load_state_dict(synthetic.make_state_dict(arch.vae_param_shapes()))
    return u, v


def test_parameter_identities(models):
    u, v = models
    assert sum(p.numel() for p in u.parameters()) == 865_925_124
    assert sum(p.numel() for p in v.parameters()) == 83_653_863
    enc = sum(p.numel() for p in v.encoder.parameters()) + sum(p.numel() for p in v.quant_conv.parameters())
    assert enc == 34_163_664 and 83_653_863 - enc == 49_490_199
    assert arch.count(arch.unet_param_shapes()) == 865_925_124
    assert arch.count(arch.vae_param_shapes()) == 83_653_863


def test_key_inventory_matches_product_arch(models):
    u, v = models
    assert {k: tuple(t.shape) for k, t in u.state_dict().items()} == dict(arch.unet_param_shapes())
    assert {k: tuple(t.shape) for k, t in v.state_dict().items()} == dict(arch.vae_param_shapes())
    # up-block resnet input widths (hidden first, skip second)
    assert [[r.conv1.in_channels for r in b.resnets] for b in u.up_blocks] == \
        [[2560, 2560, 2560], [2560, 2560, 1920], [1920, 1280, 960], [960, 640, 640]]


def test_oracle_matches_golden(models):
    u, v = models
    g = json.load(open(GOLD))
    inp = synthetic.make_inputs(1, 64, 64)
    x = torch.cat([inp["latents"], inp["mask"][:, :, ::8, ::8], inp["latents"] * 0.5], 1)
    eps = u(x, 981, inp["glyph_embeds"]).sample.flatten()
    ref = torch.tensor(g["unet_L8_t981"])
    assert ((eps - ref).abs().max() / ref.abs().max()).item() < 1e-4   # thread-count dependent summation order
    post = v.encode(inp["masked_image"]).latent_dist
    ref = torch.tensor(g["vae_moments_64px"])
    assert ((post.parameters.flatten() - ref).abs().max() / ref.abs().max()).item() < 1e-4
    rgb = sample_loop(u, v, DDIMOracle(), inp["masked_image"], inp["mask"], inp["glyph_embeds"], inp["latents"], 4,
                      posterior_noise=inp["posterior_noise"]).flatten()[::7]
    ref = torch.tensor(g["loop_64px_4steps_rgb"])
    assert ((rgb - ref).abs().max() / ref.abs().max()).item() < 1e-4


def test_timestep_forms_agree(models):
    u, _ = models
    inp = synthetic.make_inputs(2, 64, 64)
    x = torch.cat([inp["latents"], inp["mask"][:, :, ::8, ::8], inp["latents"]], 1)
    a = u(x, 21, inp["glyph_embeds"]).sample
    b = u(x, torch.tensor(21), inp["glyph_embeds"]).sample
    c = u(x, torch.tensor([21, 21]), inp["glyph_embeds"])["sample"]
    assert torch.equal(a, b) and torch.equal(a, c)
