"""DDIMScheduler / DDPMScheduler with the diffusers API surface the reference uses.

Reference call sites: app.ipynb:545 (`DDPMScheduler.from_pretrained(path, subfolder="scheduler")`), :800
(`init_noise_sigma`), :803-804 (`set_timesteps`, `timesteps`), :810 (`scale_model_input`), :816
(`step(noise_pred, t, latents).prev_sample`); train_diffute_v1.py:892-907 (`add_noise`, `get_velocity`,
`config.prediction_type`, `num_train_timesteps`).  Semantics: SURVEY.md Appendix A.3.

The beta / alpha-bar tables are a few thousand host floats computed exactly the way diffusers does (fp32 torch on
the CPU).  Each `step` collapses to per-step scalar coefficients on the host and ONE elementwise CUDA kernel
(dfu_scheduler_step) on the device; nothing is computed with torch on the device.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Union

import numpy as np
import torch

from . import arch, ops
from .unet import _Config


@dataclass
class SchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None

    def __getitem__(self, k):
        if k in (0, "prev_sample"):
            return self.prev_sample
        if k in (1, "pred_original_sample"):
            return self.pred_original_sample
        raise KeyError(k)


def _betas(cfg) -> torch.Tensor:
    n, b0, b1, kind = cfg["num_train_timesteps"], cfg["beta_start"], cfg["beta_end"], cfg["beta_schedule"]
    if cfg.get("trained_betas") is not None:
        return torch.tensor(cfg["trained_betas"], dtype=torch.float32)
    if kind == "linear":
        return torch.linspace(b0, b1, n, dtype=torch.float32)
    if kind == "scaled_linear":
        return torch.linspace(b0 ** 0.5, b1 ** 0.5, n, dtype=torch.float32) ** 2
    raise NotImplementedError(f"beta_schedule {kind!r}")


class _SchedulerBase:
    _defaults = dict(arch.SD2_SCHEDULER_CONFIG)
    order = 1

    def __init__(self, **kwargs):
        cfg = dict(self._defaults)
        cfg.update({k: v for k, v in kwargs.items() if not k.startswith("_")})
        self.config = _Config(cfg)
        self.betas = _betas(cfg)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.init_noise_sigma = 1.0
        self.num_train_timesteps = cfg["num_train_timesteps"]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, cfg["num_train_timesteps"])[::-1].copy().astype(np.int64))

    @classmethod
    def from_pretrained(cls, path, subfolder: Optional[str] = "scheduler", **kw):
        from .checkpoint import load_scheduler_config
        cfg = load_scheduler_config(path, subfolder)
        cfg.update(kw)
        return cls.from_config(cfg)

    @classmethod
    def from_config(cls, config, **kw):
        known = set(cls._defaults) | {"trained_betas"}
        cfg = {k: v for k, v in dict(config).items() if k in known}
        cfg.update(kw)
        return cls(**cfg)

    def __len__(self):
        return self.config["num_train_timesteps"]

    def scale_model_input(self, sample: torch.Tensor, timestep=None) -> torch.Tensor:
        return sample

    def _rows(self, fn, x0, noise, timesteps):
        t = timesteps.to("cpu", torch.int64).reshape(-1)
        ac = self.alphas_cumprod[t]
        ca, cb = fn(ac ** 0.5, (1 - ac) ** 0.5)
        dev = x0.device
        y = torch.empty_like(x0, dtype=torch.float32)
        ops.axpby_rows(x0.float().contiguous(), noise.float().contiguous(), ca.float().to(dev), cb.float().to(dev), y)
        return y

    def add_noise(self, original_samples, noise, timesteps):
        """sqrt(abar_t) x0 + sqrt(1-abar_t) noise, per-sample t (train_diffute_v1.py:897)."""
        return self._rows(lambda sa, sb: (sa, sb), original_samples, noise, timesteps)

    def get_velocity(self, sample, noise, timesteps):
        """sqrt(abar_t) noise - sqrt(1-abar_t) x0 (train_diffute_v1.py:907)."""
        y = self._rows(lambda sa, sb: (-sb, sa), sample, noise, timesteps)
        return y

    # coefficients of x0 = a0*x + a1*m and eps = e0*x + e1*m for the configured prediction type
    def _pred_coeffs(self, a_t: float):
        sa, sb = math.sqrt(a_t), math.sqrt(1.0 - a_t)
        pt = self.config["prediction_type"]
        if pt == "epsilon":
            return (1.0 / sa, -sb / sa), (0.0, 1.0)
        if pt == "sample":
            return (0.0, 1.0), (1.0 / sb, -sa / sb)
        if pt == "v_prediction":
            return (sa, -sb), (sb, sa)
        raise ValueError(f"prediction_type {pt!r}")


def _t_int(timestep) -> int:
    if torch.is_tensor(timestep):
        return int(timestep.item()) if timestep.numel() == 1 else int(timestep.reshape(-1)[0].item())
    return int(timestep)


class DDIMScheduler(_SchedulerBase):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.final_alpha_cumprod = (torch.tensor(1.0) if self.config["set_alpha_to_one"] else self.alphas_cumprod[0])

    def set_timesteps(self, num_inference_steps: int, device=None):
        n_train = self.config["num_train_timesteps"]
        if num_inference_steps > n_train:
            raise ValueError(f"num_inference_steps {num_inference_steps} > num_train_timesteps {n_train}")
        self.num_inference_steps = num_inference_steps
        ratio = n_train // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts + self.config["steps_offset"])
        if device is not None:
            self.timesteps = self.timesteps.to(device)

    def step_coefficients(self, t: int, eta: float = 0.0):
        """(a0, a1, p0, d0, d1, sigma, clip): prev = p0*clip?(a0 x + a1 m) + d0 x + d1 m + sigma*noise.  fp64 host math
        on the fp32 alpha-bar table, exactly the quantities of Appendix A.3."""
        if self.num_inference_steps is None:
            raise ValueError("call set_timesteps() before step()")
        p = t - self.config["num_train_timesteps"] // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_p = float(self.alphas_cumprod[p]) if p >= 0 else float(self.final_alpha_cumprod)
        (a0, a1), (e0, e1) = self._pred_coeffs(a_t)
        var = (1 - a_p) / (1 - a_t) * (1 - a_t / a_p)
        sigma = eta * math.sqrt(max(var, 0.0))
        dirc = math.sqrt(max(1 - a_p - sigma ** 2, 0.0))
        return a0, a1, math.sqrt(a_p), dirc * e0, dirc * e1, sigma, bool(self.config["clip_sample"])

    def collapsed_coefficients(self, t: int):
        """(cx, ce) with prev = cx*x + ce*m when no clipping and eta = 0 — used by the fused conv_out."""
        a0, a1, p0, d0, d1, sigma, clip = self.step_coefficients(t, 0.0)
        if clip:
            raise ValueError("collapsed form needs clip_sample=False")
        return p0 * a0 + d0, p0 * a1 + d1

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True):
        t = _t_int(timestep)
        a0, a1, p0, d0, d1, sigma, clip = self.step_coefficients(t, eta)
        if use_clipped_model_output and clip:
            raise NotImplementedError("use_clipped_model_output with clip_sample is not used by the reference")
        x = sample.float().contiguous()
        m = model_output.float().contiguous()
        noise = None
        if eta > 0:
            if variance_noise is None:
                dev = generator.device if generator is not None else x.device
                variance_noise = torch.randn(m.shape, generator=generator, device=dev, dtype=torch.float32)
            noise = variance_noise.to(x.device, torch.float32).contiguous()
        prev = torch.empty_like(x)
        x0 = torch.empty_like(x) if return_dict else None
        ops.scheduler_step(x, m, noise, a0, a1, p0, d0, d1, sigma, clip, prev, x0)
        return SchedulerOutput(prev, x0) if return_dict else (prev,)


class DDPMScheduler(_SchedulerBase):
    """What the reference actually runs (app.ipynb:545, 150 steps by default :914).  Later-diffusers semantics:
    alpha_t = abar_t / abar_prev, exactly N timesteps (SURVEY A.3 notes the <=0.15 variant)."""
    _defaults = dict(arch.SD2_SCHEDULER_CONFIG, variance_type="fixed_small")

    def set_timesteps(self, num_inference_steps: int, device=None):
        n_train = self.config["num_train_timesteps"]
        if num_inference_steps > n_train:
            raise ValueError(f"num_inference_steps {num_inference_steps} > num_train_timesteps {n_train}")
        self.num_inference_steps = num_inference_steps
        ratio = n_train // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts)
        if device is not None:
            self.timesteps = self.timesteps.to(device)

    def step_coefficients(self, t: int):
        n = self.num_inference_steps or self.config["num_train_timesteps"]
        p = t - self.config["num_train_timesteps"] // n
        a_t = float(self.alphas_cumprod[t])
        a_p = float(self.alphas_cumprod[p]) if p >= 0 else 1.0
        cur_a = a_t / a_p
        cur_b = 1.0 - cur_a
        (a0, a1), _ = self._pred_coeffs(a_t)
        c0 = math.sqrt(a_p) * cur_b / (1.0 - a_t)
        c1 = math.sqrt(cur_a) * (1.0 - a_p) / (1.0 - a_t)
        var = max((1.0 - a_p) / (1.0 - a_t) * cur_b, 1e-20)
        if self.config["variance_type"] not in ("fixed_small",):
            raise NotImplementedError(f"variance_type {self.config['variance_type']!r}")
        sigma = math.sqrt(var) if t > 0 else 0.0
        return a0, a1, c0, c1, 0.0, sigma, bool(self.config["clip_sample"])

    def collapsed_coefficients(self, t: int):
        """(cx, ce, sigma) with prev = cx*x + ce*m + sigma*z when no clipping -- the fused conv_out epilogue's form."""
        a0, a1, p0, d0, d1, sigma, clip = self.step_coefficients(t)
        if clip:
            raise ValueError("collapsed form needs clip_sample=False")
        return p0 * a0 + d0, p0 * a1 + d1, sigma

    def step(self, model_output, timestep, sample, generator=None, return_dict: bool = True):
        t = _t_int(timestep)
        a0, a1, p0, d0, d1, sigma, clip = self.step_coefficients(t)
        x = sample.float().contiguous()
        m = model_output.float().contiguous()
        noise = None
        if t > 0:
            dev = generator.device if generator is not None else x.device
            noise = torch.randn(m.shape, generator=generator, device=dev, dtype=torch.float32).to(x.device)
        prev = torch.empty_like(x)
        x0 = torch.empty_like(x) if return_dict else None
        ops.scheduler_step(x, m, noise, a0, a1, p0, d0, d1, sigma, clip, prev, x0)
        return SchedulerOutput(prev, x0) if return_dict else (prev,)
