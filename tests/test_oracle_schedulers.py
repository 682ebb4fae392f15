"""Pin the scheduler oracle against six upstream diffusers known-answer constants."""
import pytest
import torch

from oracle.schedulers import DDIMOracle, DDPMOracle


def _dummy_sample():
    n = 4 * 3 * 8 * 8
    return (torch.arange(n).reshape(3, 8, 8, 4) / n).permute(3, 0, 1, 2)


def _model(x, t):
    return x * t / (t + 1)


def _ddim_loop(**cfg):
    base = dict(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon")
    base.update(cfg)
    s = DDIMOracle(**base)
    s.set_timesteps(10)
    x = _dummy_sample()
    for t in s.timesteps:
        x = s.step(_model(x, t), t, x, eta=0.0).prev_sample
    return x.abs().sum().item(), x.abs().mean().item()


@pytest.mark.parametrize("cfg,exp_sum,exp_mean", [
    (dict(), 172.0067, 0.223967),
    (dict(prediction_type="v_prediction"), 52.5302, 0.0684),
    (dict(set_alpha_to_one=True, beta_start=0.01), 149.8295, 0.1951),
    (dict(set_alpha_to_one=False, beta_start=0.01), 149.0784, 0.1941),
])
def test_ddim_full_loop_kat(cfg, exp_sum, exp_mean):
    s, m = _ddim_loop(**cfg)
    assert abs(s - exp_sum) < 1e-2
    assert abs(m - exp_mean) < 1e-3


@pytest.mark.parametrize("pt,exp_sum,exp_mean", [("epsilon", 258.9606, 0.3372), ("v_prediction", 202.0296, 0.2631)])
def test_ddpm_full_loop_kat(pt, exp_sum, exp_mean):
    s = DDPMOracle(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                   clip_sample=True, prediction_type=pt)
    g = torch.manual_seed(0)
    x = _dummy_sample()
    for t in reversed(range(1000)):
        x = s.step(_model(x, t), t, x, generator=g).prev_sample
    assert abs(x.abs().sum().item() - exp_sum) < 1e-2
    assert abs(x.abs().mean().item() - exp_mean) < 1e-3


def test_sd2_ddim_timesteps_and_collapse():
    s = DDIMOracle()
    s.set_timesteps(50)
    assert s.timesteps.tolist() == list(range(981, 0, -20))
    # collapsed 2-coefficient form == full step for every SD2 timestep
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
    e = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
    s.alphas_cumprod = s.alphas_cumprod.double()
    s.final_alpha_cumprod = s.alphas_cumprod[0]
    for t in s.timesteps:
        cx, ce = s.collapsed_coeffs(int(t))
        ref = s.step(e, t, x).prev_sample
        assert torch.allclose(cx * x + ce * e, ref, atol=1e-12)


def test_add_noise_velocity():
    s = DDPMOracle()
    x0 = torch.randn(3, 4, 8, 8)
    n = torch.randn(3, 4, 8, 8)
    t = torch.tensor([0, 500, 999])
    a = s.alphas_cumprod[t].view(3, 1, 1, 1)
    assert torch.allclose(s.add_noise(x0, n, t), a.sqrt() * x0 + (1 - a).sqrt() * n)
    assert torch.allclose(s.get_velocity(x0, n, t), a.sqrt() * n - (1 - a).sqrt() * x0)
