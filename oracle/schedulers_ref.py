"""ORACLE (test infrastructure): DDIM / inverse-DDIM schedulers for the I2VGen-XL config.

Follows the reference's vendored copy of the diffusers class, /root/reference/consisti2v/ddim_inverse_scheduler.py:
  betas_for_alpha_bar :49-90, rescale_zero_terminal_snr :94-127, set_timesteps :253-289, step :291-373.
DDIMScheduler.step (diffusers, un-vendored) is the mirrored formula with eta = 0.
Scheduler config is pinned by i2vgen-xl/demo.ipynb:1209-1225.

Rounding model: on the reference's CUDA path `alpha ** 0.5` is a 0-dim fp32 CPU tensor and `coef * fp16_cuda_tensor`
is evaluated as fp32(coef) * fp32(x) rounded to fp16, one rounding per PyTorch op.  `_mul` / `_add` below spell that
out explicitly so the oracle gives the same bits on CPU and on CUDA, in fp16; in fp32/fp64 they are plain ops.
"""
from __future__ import annotations

import math

import numpy as np
import torch

CONFIG = dict(num_train_timesteps=1000, beta_schedule="squaredcos_cap_v2", beta_start=1e-4, beta_end=0.02,
              prediction_type="v_prediction", rescale_betas_zero_snr=True, clip_sample=False, set_alpha_to_one=True,
              steps_offset=1, timestep_spacing="leading")


def _cosine_betas(n: int, max_beta: float = 0.999) -> torch.Tensor:
    bar = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    return torch.tensor([min(1 - bar((i + 1) / n) / bar(i / n), max_beta) for i in range(n)], dtype=torch.float32)


def _zero_terminal_snr(betas: torch.Tensor) -> torch.Tensor:
    abar_sqrt = torch.cumprod(1.0 - betas, dim=0).sqrt()
    first, last = abar_sqrt[0].clone(), abar_sqrt[-1].clone()
    abar_sqrt -= last
    abar_sqrt *= first / (first - last)
    abar = abar_sqrt ** 2
    alphas = torch.cat([abar[0:1], abar[1:] / abar[:-1]])
    return 1 - alphas


def alphas_cumprod() -> torch.Tensor:
    betas = _zero_terminal_snr(_cosine_betas(CONFIG["num_train_timesteps"]))
    return torch.cumprod(1.0 - betas, dim=0)


def _mul(coef: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    if x.dtype == torch.float16:
        return (coef.float() * x.float()).half()
    return coef.to(x.dtype) * x


def _add(a: torch.Tensor, b: torch.Tensor, sign: float = 1.0) -> torch.Tensor:
    if a.dtype == torch.float16:
        return (a.float() + sign * b.float()).half()
    return a + sign * b


class _Base:
    def __init__(self):
        self.alphas_cumprod = alphas_cumprod()
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.initial_alpha_cumprod = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.order = 1
        self.num_inference_steps = None
        self.timesteps = None

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _leading(self, n: int) -> np.ndarray:
        ratio = CONFIG["num_train_timesteps"] // n
        return (np.arange(0, n) * ratio).round().astype(np.int64) + CONFIG["steps_offset"]

    @staticmethod
    def _vpred_update(x, v, a_in: torch.Tensor, a_out: torch.Tensor):
        ca, cb = a_in ** 0.5, (1 - a_in) ** 0.5
        x0 = _add(_mul(ca, x), _mul(cb, v), -1.0)
        eps = _add(_mul(ca, v), _mul(cb, x))
        direction = _mul((1 - a_out) ** 0.5, eps)
        return _add(_mul(a_out ** 0.5, x0), direction), x0


class DDIMInverseScheduler(_Base):
    def set_timesteps(self, n: int, device=None):
        self.num_inference_steps = n
        self.timesteps = torch.from_numpy(self._leading(n).copy()).to(device)  # ascending [1, 21, ...]

    def step(self, model_output, timestep, sample):
        t_next = int(timestep)
        t_cur = min(t_next - CONFIG["num_train_timesteps"] // self.num_inference_steps, CONFIG["num_train_timesteps"] - 1)
        a_cur = self.alphas_cumprod[t_cur] if t_cur >= 0 else self.initial_alpha_cumprod
        a_next = self.alphas_cumprod[t_next]
        out, x0 = self._vpred_update(sample, model_output, a_cur, a_next)
        return out, x0

    def coefficients(self, timestep):
        t_next = int(timestep)
        t_cur = min(t_next - CONFIG["num_train_timesteps"] // self.num_inference_steps, CONFIG["num_train_timesteps"] - 1)
        a_in = self.alphas_cumprod[t_cur] if t_cur >= 0 else self.initial_alpha_cumprod
        a_out = self.alphas_cumprod[t_next]
        return tuple(float(c) for c in (a_in ** 0.5, (1 - a_in) ** 0.5, a_out ** 0.5, (1 - a_out) ** 0.5))


class DDIMScheduler(_Base):
    def set_timesteps(self, n: int, device=None):
        self.num_inference_steps = n
        self.timesteps = torch.from_numpy(self._leading(n)[::-1].copy()).to(device)  # descending [981, 961, ...]

    def step(self, model_output, timestep, sample, eta: float = 0.0):
        assert eta == 0.0, "the reference path always samples with eta = 0"
        t = int(timestep)
        t_prev = t - CONFIG["num_train_timesteps"] // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[t_prev] if t_prev >= 0 else self.final_alpha_cumprod
        out, x0 = self._vpred_update(sample, model_output, a_t, a_prev)
        return out, x0

    def coefficients(self, timestep):
        t = int(timestep)
        t_prev = t - CONFIG["num_train_timesteps"] // self.num_inference_steps
        a_in = self.alphas_cumprod[t]
        a_out = self.alphas_cumprod[t_prev] if t_prev >= 0 else self.final_alpha_cumprod
        return tuple(float(c) for c in (a_in ** 0.5, (1 - a_in) ** 0.5, a_out ** 0.5, (1 - a_out) ** 0.5))


def cfg_combine(neg, edit, guidance: float):
    """pipeline_i2vgen_xl.py:1162 — neg + g * (edit - neg), one fp16 rounding per op."""
    if neg.dtype == torch.float16:
        d = (edit.float() - neg.float()).half()
        d = (guidance * d.float()).half()
        return (neg.float() + d.float()).half()
    return neg + guidance * (edit - neg)
