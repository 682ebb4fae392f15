"""GPU parity of the whole UNet denoising step (diffusers call signature) against the fp32 CPU oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffute_b200 import arch, synthetic
    from oracle import UNetOracle
    sd = synthetic.make_state_dict(arch.unet_param_shapes())
    orc = UNetOracle()
    orc.load_state_dict(sd)
    return sd, orc


def _inputs(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    sample = torch.randn((B, 9, L, L), generator=g)
    ehs = torch.randn((B, 577, 1024), generator=g)
    return sample, ehs


@pytest.mark.parametrize("precision,tol", [("fp16", 4e-3), ("fp16x2", 2e-4)])
@pytest.mark.parametrize("graph", [False, True])
def test_unet_step_parity(setup, precision, tol, graph):
    from diffute_b200.unet import UNet2DConditionModel
    sd, orc = setup
    unet = UNet2DConditionModel(sd, precision=precision, use_cuda_graph=graph)
    for (B, L, t) in [(1, 32, 981), (2, 16, torch.tensor([500, 21]))]:
        sample, ehs = _inputs(B, L, 7 + B)
        ref = orc(sample, t, ehs).sample
        for rep in range(2):  # second call exercises the cached context / graph replay
            got = unet(sample.cuda(), t, ehs.cuda() if rep == 0 else ehs).sample.cpu()
            err = ((got - ref).abs().max() / ref.abs().max()).item()
            print(f"unet parity {precision} graph={graph} B={B} L={L} rep={rep}: maxrel {err:.3e}")
            assert err < tol, err
    # a 0-d tensor timestep and return_dict=False, as the reference passes them (app.ipynb:814)
    sample, ehs = _inputs(1, 16, 3)
    out = unet(sample.cuda(), torch.tensor(981), ehs.cuda(), return_dict=False)[0]
    ref = orc(sample, torch.tensor(981), ehs).sample
    assert ((out.cpu() - ref).abs().max() / ref.abs().max()).item() < tol


@pytest.mark.parametrize("precision,tol", [("fp16", 4e-3), ("fp16x2", 2e-4)])
def test_unet_step_parity_at_baseline_size(setup, precision, tol):
    """L=64 (512 px): the tilings bench.py runs — 64-wide / 2-row conv boxes, 160-tile stream-K attention, the level-0
    split-K clusters — which the L<=32 cases never reach."""
    from diffute_b200.unet import UNet2DConditionModel
    sd, orc = setup
    unet = UNet2DConditionModel(sd, precision=precision, use_cuda_graph=True)
    sample, ehs = _inputs(1, 64, 21)
    ref = orc(sample, 981, ehs).sample
    for rep in range(2):
        got = unet(sample.cuda(), 981, ehs.cuda()).sample.cpu()
        err = ((got - ref).abs().max() / ref.abs().max()).item()
        print(f"unet parity {precision} B=1 L=64 rep={rep}: maxrel {err:.3e}")
        assert err < tol, err


def test_unet_context_length_change_recaptures_graph(setup):
    """A captured step bakes the glyph-context length (Nk and the K/V row strides) into its kernel arguments: a call
    with a shorter encoder_hidden_states must not replay the graph captured for the longer one."""
    from diffute_b200.unet import UNet2DConditionModel
    sd, orc = setup
    unet = UNet2DConditionModel(sd, precision="fp16x2", use_cuda_graph=True)
    g = torch.Generator().manual_seed(5)
    sample = torch.randn((1, 9, 16, 16), generator=g)
    for T in (577, 77, 577, 130):
        ehs = torch.randn((1, T, 1024), generator=g)
        ref = orc(sample, 500, ehs).sample
        got = unet(sample.cuda(), 500, ehs.cuda()).sample.cpu()
        err = ((got - ref).abs().max() / ref.abs().max()).item()
        print(f"unet ctx T={T}: maxrel {err:.3e}")
        assert err < 2e-4, (T, err)


@pytest.mark.parametrize("safe", [True, False])
def test_unet_checkpoint_folder_round_trip(setup, tmp_path, safe):
    """f1: the on-disk format the reference writes with `unet.save_pretrained(<out>/unet)` (train_diffute_v1.py:669)
    and reads with `UNet2DConditionModel.from_pretrained(path, subfolder="unet")` (app.ipynb:551-553,
    train_diffute_v1.py:684), as .safetensors and as .bin; plus the nn.Module surface of the load hook
    (`register_to_config(**loaded.config)`, `load_state_dict(loaded.state_dict())`, train_diffute_v1.py:687-689)."""
    import os
    from diffute_b200.unet import UNet2DConditionModel
    sd, orc = setup
    a = UNet2DConditionModel(sd, precision="fp16x2")
    a.save_pretrained(os.path.join(tmp_path, "unet"), safe_serialization=safe)
    assert os.path.isfile(tmp_path / "unet" / ("diffusion_pytorch_model.safetensors" if safe else "diffusion_pytorch_model.bin"))
    b = UNet2DConditionModel.from_pretrained(str(tmp_path), subfolder="unet", precision="fp16x2")
    assert b.config.cross_attention_dim == 1024 and list(b.config.block_out_channels) == [320, 640, 1280, 1280]
    got_sd = b.state_dict()
    assert list(got_sd) == list(sd) and all(torch.equal(got_sd[k].cpu(), sd[k]) for k in sd)
    assert sum(p.numel() for p in b.parameters()) == 865_925_124
    sample, ehs = _inputs(1, 16, 31)
    ref = orc(sample, 500, ehs).sample
    err = ((b(sample.cuda(), 500, ehs.cuda()).sample.cpu() - ref).abs().max() / ref.abs().max()).item()
    print(f"from_pretrained(safe={safe}) vs oracle: maxrel {err:.3e}")
    assert err < 2e-4
    if safe:
        # the accelerate load hook: copy config + weights of a freshly loaded model into a live one
        sd2 = {k: (v * 1.5 if k.endswith("conv_out.weight") else v) for k, v in sd.items()}
        c = UNet2DConditionModel(sd2, precision="fp16x2")
        out_before = c(sample.cuda(), 500, ehs.cuda()).sample.cpu()
        assert ((out_before - ref).abs().max() / ref.abs().max()).item() > 1e-2
        c.register_to_config(**b.config)
        c.load_state_dict(b.state_dict())
        err = ((c(sample.cuda(), 500, ehs.cuda()).sample.cpu() - ref).abs().max() / ref.abs().max()).item()
        assert err < 2e-4, err
        with pytest.raises(ValueError):
            c.register_to_config(cross_attention_dim=768)
        with pytest.raises(KeyError):
            c.load_state_dict({k: v for k, v in sd.items() if k != "conv_in.bias"})


def test_unet_rejects_bad_shapes(setup):
    from diffute_b200.unet import UNet2DConditionModel
    sd, _ = setup
    unet = UNet2DConditionModel(sd, precision="fp16", use_cuda_graph=False)
    with pytest.raises(ValueError):
        unet(torch.zeros(1, 4, 16, 16).cuda(), 1, torch.zeros(1, 577, 1024).cuda())
    with pytest.raises(ValueError):
        unet(torch.zeros(1, 9, 12, 12).cuda(), 1, torch.zeros(1, 577, 1024).cuda())
    with pytest.raises(ValueError):
        unet(torch.zeros(1, 9, 16, 16).cuda(), 1, torch.zeros(1, 577, 768).cuda())
