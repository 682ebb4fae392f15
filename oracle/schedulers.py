"""CPU restatement of diffusers DDIMScheduler / DDPMScheduler (TEST INFRASTRUCTURE).

Reference call sites: app.ipynb:545 (from_pretrained), :800 (init_noise_sigma),
:803-804 (set_timesteps / timesteps), :810 (scale_model_input), :816 (step);
train_diffute_v1.py:892-907 (add_noise / get_velocity / prediction_type).
Math follows SURVEY.md Appendix A.3.  PINNED by six upstream diffusers
known-answer constants (tests/test_oracle_schedulers.py):
  DDIM full loop eps-pred  172.0067 / 0.223967,  v-pred 52.5302 / 0.0684,
  set_alpha_to_one True/False 149.8295 / 0.1951, 149.0784 / 0.1941,
  DDPM full loop 258.9606 / 0.3372, v-pred 202.0296 / 0.2631.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

SD2_SCHEDULER_CONFIG = dict(
    num_train_timesteps=1000,
    beta_start=0.00085,
    beta_end=0.012,
    beta_schedule="scaled_linear",
    set_alpha_to_one=False,
    steps_offset=1,
    clip_sample=False,
    prediction_type="epsilon",
)


def make_betas(num_train_timesteps, beta_start, beta_end, beta_schedule):
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    raise ValueError(beta_schedule)


@dataclass
class StepOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class _Base:
    def __init__(self, **cfg):
        c = dict(SD2_SCHEDULER_CONFIG)
        c.update(cfg)
        self.config = c
        self.betas = make_betas(c["num_train_timesteps"], c["beta_start"], c["beta_end"], c["beta_schedule"])
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.init_noise_sigma = 1.0
        self.num_train_timesteps = c["num_train_timesteps"]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, c["num_train_timesteps"])[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def add_noise(self, x0, noise, timesteps):
        ac = self.alphas_cumprod.to(x0.dtype)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        while sa.dim() < x0.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * x0 + sb * noise

    def get_velocity(self, x0, noise, timesteps):
        ac = self.alphas_cumprod.to(x0.dtype)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        while sa.dim() < x0.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * noise - sb * x0

    def _x0_eps(self, model_output, sample, a_t):
        b_t = 1 - a_t
        pt = self.config["prediction_type"]
        if pt == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        elif pt == "sample":
            x0 = model_output
            eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        elif pt == "v_prediction":
            x0 = a_t ** 0.5 * sample - b_t ** 0.5 * model_output
            eps = a_t ** 0.5 * model_output + b_t ** 0.5 * sample
        else:
            raise ValueError(pt)
        return x0, eps


class DDIMOracle(_Base):
    def __init__(self, **cfg):
        super().__init__(**cfg)
        self.final_alpha_cumprod = torch.tensor(1.0) if self.config["set_alpha_to_one"] else self.alphas_cumprod[0]

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        ratio = self.config["num_train_timesteps"] // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + self.config["steps_offset"]
        self.timesteps = torch.from_numpy(ts)

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True):
        t = int(timestep)
        p = t - self.config["num_train_timesteps"] // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[p] if p >= 0 else self.final_alpha_cumprod
        x0, eps = self._x0_eps(model_output, sample, a_t)
        if self.config["clip_sample"]:
            x0 = x0.clamp(-1, 1)
        var = (1 - a_p) / (1 - a_t) * (1 - a_t / a_p)
        sigma = eta * var ** 0.5
        if use_clipped_model_output:
            eps = (sample - a_t ** 0.5 * x0) / (1 - a_t) ** 0.5
        direction = (1 - a_p - sigma ** 2) ** 0.5 * eps
        prev = a_p ** 0.5 * x0 + direction
        if eta > 0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            prev = prev + sigma * variance_noise
        return StepOutput(prev, x0) if return_dict else (prev,)

    def collapsed_coeffs(self, t: int):
        """(cx, ce) with x' = cx*x + ce*eps for eps-prediction, no clip, eta 0 (fp64)."""
        ac = self.alphas_cumprod.double()
        p = t - self.config["num_train_timesteps"] // self.num_inference_steps
        a_t = ac[t]
        a_p = ac[p] if p >= 0 else self.final_alpha_cumprod.double()
        cx = (a_p / a_t) ** 0.5
        ce = (1 - a_p) ** 0.5 - (a_p * (1 - a_t) / a_t) ** 0.5
        return float(cx), float(ce)


class DDPMOracle(_Base):
    """Later-diffusers DDPM (current_alpha_t = abar_t/abar_prev; exactly N timesteps). Appendix A.3."""

    def __init__(self, variance_type: str = "fixed_small", **cfg):
        super().__init__(**cfg)
        self.variance_type = variance_type

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        ratio = self.config["num_train_timesteps"] // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts)

    def step(self, model_output, timestep, sample, generator=None, noise=None, return_dict: bool = True):
        t = int(timestep)
        n = self.num_inference_steps or self.config["num_train_timesteps"]
        p = t - self.config["num_train_timesteps"] // n
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[p] if p >= 0 else torch.tensor(1.0)
        b_t, b_p = 1 - a_t, 1 - a_p
        cur_a = a_t / a_p
        cur_b = 1 - cur_a
        x0, _ = self._x0_eps(model_output, sample, a_t)
        if self.config["clip_sample"]:
            x0 = x0.clamp(-1, 1)
        c0 = (a_p ** 0.5 * cur_b) / b_t
        c1 = cur_a ** 0.5 * b_p / b_t
        prev = c0 * x0 + c1 * sample
        if t > 0:
            if noise is None:
                noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            var = torch.clamp(b_p / b_t * cur_b, min=1e-20)
            prev = prev + var ** 0.5 * noise
        return StepOutput(prev, x0) if return_dict else (prev,)
