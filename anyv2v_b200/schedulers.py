"""DDIM / inverse-DDIM schedulers for the I2VGen-XL scheduler config, host side + fused device step.

API of the diffusers classes the reference uses (``set_timesteps``, ``timesteps``, ``step(...).prev_sample``,
``scale_model_input``, ``init_noise_sigma``, ``order``; run_group_pnp_edit.py:69-72, pipeline_i2vgen_xl.py:1104,1173,
1359,1418).  Host side: the beta schedule, zero-terminal-SNR rescale and "leading" timestep spacing follow
consisti2v/ddim_inverse_scheduler.py:49-127, 253-289 (the reference's vendored copy of the diffusers class), config
pinned at i2vgen-xl/demo.ipynb:1209-1225.  Device side: ``step`` is ONE kernel launch (csrc/elementwise.cu) that also
folds in classifier-free guidance when given both model outputs, reproducing the reference's fp16 rounding sequence
bit for bit (each PyTorch op of scheduler.step / pipeline :1162 rounds to fp16 separately).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

from . import ops

DEFAULT_CONFIG = dict(num_train_timesteps=1000, beta_schedule="squaredcos_cap_v2", beta_start=1e-4, beta_end=0.02,
                      prediction_type="v_prediction", rescale_betas_zero_snr=True, clip_sample=False,
                      set_alpha_to_one=True, steps_offset=1, timestep_spacing="leading", thresholding=False)


def _alphas_cumprod(cfg) -> torch.Tensor:
    n = cfg["num_train_timesteps"]
    if cfg["beta_schedule"] != "squaredcos_cap_v2":
        raise ValueError("only the I2VGen-XL beta schedule (squaredcos_cap_v2) is supported")
    bar = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    betas = torch.tensor([min(1 - bar((i + 1) / n) / bar(i / n), 0.999) for i in range(n)], dtype=torch.float32)
    if cfg["rescale_betas_zero_snr"]:
        s = torch.cumprod(1.0 - betas, dim=0).sqrt()
        s0, sT = s[0].clone(), s[-1].clone()
        s = (s - sT) * (s0 / (s0 - sT))
        abar = s ** 2
        betas = 1 - torch.cat([abar[0:1], abar[1:] / abar[:-1]])
    return torch.cumprod(1.0 - betas, dim=0)


class _DDIMBase:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, **config):
        cfg = dict(DEFAULT_CONFIG)
        cfg.update(config)
        if cfg["prediction_type"] != "v_prediction" or cfg["clip_sample"] or cfg["thresholding"]:
            raise ValueError("only the I2VGen-XL scheduler config (v_prediction, no clipping) is supported")
        self.config = SimpleNamespace(**cfg)
        self.alphas_cumprod = _alphas_cumprod(cfg)  # fp32, CPU
        one = torch.tensor(1.0)
        self.final_alpha_cumprod = one if cfg["set_alpha_to_one"] else self.alphas_cumprod[0]
        self.initial_alpha_cumprod = self.final_alpha_cumprod
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, cfg["num_train_timesteps"])[::-1].copy().astype(np.int64))

    @classmethod
    def from_pretrained(cls, *_args, **kwargs):
        """The reference loads the scheduler config from the hub; offline we use the pinned config."""
        kwargs.pop("subfolder", None)
        return cls(**kwargs)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _leading(self, n: int) -> np.ndarray:
        if n > self.config.num_train_timesteps:
            raise ValueError(f"`num_inference_steps`: {n} cannot be larger than {self.config.num_train_timesteps}")
        if self.config.timestep_spacing != "leading":
            raise ValueError("only timestep_spacing='leading' is supported")
        ratio = self.config.num_train_timesteps // n
        return (np.arange(0, n) * ratio).round().astype(np.int64) + self.config.steps_offset

    def _alpha_pair(self, timestep):
        raise NotImplementedError

    def coefficients(self, timestep):
        """(sqrt(a_in), sqrt(1-a_in), sqrt(a_out), sqrt(1-a_out)) as Python floats holding exact fp32 values,
        computed with the same fp32 torch ops as the reference (`alpha ** 0.5`)."""
        a_in, a_out = self._alpha_pair(int(timestep))
        return tuple(float(c) for c in (a_in ** 0.5, (1 - a_in) ** 0.5, a_out ** 0.5, (1 - a_out) ** 0.5))

    def step(self, model_output, timestep, sample, eta: float = 0.0, return_dict: bool = True, *,
             model_output_cond=None, guidance_scale: float = 1.0, out=None, coef_dev=None, **_unused):
        """x_t -> x_{t-1} (DDIM) or x_t -> x_{t+1} (inverse).  With ``model_output_cond`` the CFG combine
        ``uncond + g*(cond-uncond)`` (pipeline :1162) is fused into the same launch."""
        if eta != 0.0:
            raise ValueError("the AnyV2V path samples with eta = 0")
        ca, cb, cc, cd = (0.0, 0.0, 0.0, 0.0) if coef_dev is not None else self.coefficients(timestep)
        prev = ops.ddim_step(sample.contiguous(), model_output.contiguous(),
                             None if model_output_cond is None else model_output_cond.contiguous(),
                             float(guidance_scale), ca, cb, cc, cd, out=out, inverse=self._inverse, coef_dev=coef_dev)
        prev = prev.view(sample.shape)
        if not return_dict:
            return (prev,)
        return SimpleNamespace(prev_sample=prev)


    def coefficient_table(self, timesteps, guidance_scale: float, device) -> torch.Tensor:
        """[len(timesteps), 5] fp32 device table {ca, cb, cc, cd, guidance}: the per-step scalars of ``step`` as DATA,
        so a captured CUDA graph of one loop iteration can be replayed for every timestep."""
        rows = [list(self.coefficients(t)) + [float(guidance_scale)] for t in timesteps]
        return torch.tensor(rows, dtype=torch.float32).to(device)


class DDIMScheduler(_DDIMBase):
    _inverse = False

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        # kept on the host: the loops index Python ints, never `.item()` a device tensor (pipeline :1143 does)
        self.timesteps = torch.from_numpy(self._leading(num_inference_steps)[::-1].copy())

    def _alpha_pair(self, t):
        t_prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_prev = self.alphas_cumprod[t_prev] if t_prev >= 0 else self.final_alpha_cumprod
        return self.alphas_cumprod[t], a_prev


class DDIMInverseScheduler(_DDIMBase):
    _inverse = True

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.from_numpy(self._leading(num_inference_steps).copy())

    def _alpha_pair(self, t_next):
        t_cur = min(t_next - self.config.num_train_timesteps // self.num_inference_steps,
                    self.config.num_train_timesteps - 1)
        a_cur = self.alphas_cumprod[t_cur] if t_cur >= 0 else self.initial_alpha_cumprod
        return a_cur, self.alphas_cumprod[t_next]
