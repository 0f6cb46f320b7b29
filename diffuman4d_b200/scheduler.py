"""Host-side DDIM tables for the fused CFG+DDIM kernel.

Mirrors what the reference obtains from ``self.scheduler.set_timesteps(n)`` + per-frame deep copies
(pipeline_diffuman4d.py:265-271) for upstream diffusers==0.33.1 ``DDIMScheduler``: the ``timesteps`` vector and
``alphas_cumprod`` / ``final_alpha_cumprod``.  Only table construction lives here (numpy, host); the update itself
runs on the GPU (csrc/elementwise.cu ``cfg_ddim_kernel``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import D4DSched
from .config import SchedulerConfig

_PRED = {"epsilon": 0, "v_prediction": 1, "sample": 2}


class DDIMTables:
    init_noise_sigma = 1.0  # DDIM: scale_model_input is the identity, init sigma 1 (reference PIPE:189,376)

    def __init__(self, cfg: SchedulerConfig = None, device="cuda:0"):
        self.config = cfg or SchedulerConfig()
        c = self.config
        T = c.num_train_timesteps
        if c.beta_schedule == "scaled_linear":
            betas = torch.linspace(c.beta_start ** 0.5, c.beta_end ** 0.5, T, dtype=torch.float32) ** 2
        elif c.beta_schedule == "linear":
            betas = torch.linspace(c.beta_start, c.beta_end, T, dtype=torch.float32)
        else:
            raise ValueError(f"{c.beta_schedule} is not implemented")
        if c.prediction_type not in _PRED:
            raise ValueError(f"prediction_type given as {c.prediction_type} must be one of {list(_PRED)}")
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = 1.0 if c.set_alpha_to_one else float(self.alphas_cumprod[0])
        self.device = torch.device(device)
        self._alphas_dev = None
        self.num_inference_steps = None
        self.timesteps = None          # host int64, like scheduler.timesteps
        self._timesteps_dev = None

    def set_timesteps(self, n: int, device=None):
        c = self.config
        T = c.num_train_timesteps
        if n > T:
            raise ValueError(f"`num_inference_steps`: {n} cannot be larger than `self.config.train_timesteps`: {T}")
        self.num_inference_steps = n
        if c.timestep_spacing == "leading":
            ts = (np.arange(0, n) * (T // n)).round()[::-1].copy().astype(np.int64) + c.steps_offset
        elif c.timestep_spacing == "trailing":
            ts = np.round(np.arange(T, 0, -T / n)).astype(np.int64) - 1
        elif c.timestep_spacing == "linspace":
            ts = np.linspace(0, T - 1, n).round()[::-1].copy().astype(np.int64)
        else:
            raise ValueError(f"{c.timestep_spacing} is not supported")
        self.timesteps = torch.from_numpy(ts)
        self._timesteps_dev = None
        return self.timesteps

    def c_struct(self, emulate_bf16: bool = False) -> D4DSched:
        if self.timesteps is None:
            raise ValueError("call set_timesteps first")
        if self._alphas_dev is None:
            self._alphas_dev = self.alphas_cumprod.to(self.device)
        if self._timesteps_dev is None:
            self._timesteps_dev = self.timesteps.to(self.device)
        c = self.config
        s = D4DSched()
        s.timesteps_table = self._timesteps_dev.data_ptr()
        s.alphas_cumprod = self._alphas_dev.data_ptr()
        s.n_steps = int(self.num_inference_steps)
        s.num_train_timesteps = int(c.num_train_timesteps)
        s.final_alpha_cumprod = float(self.final_alpha_cumprod)
        s.prediction_type = _PRED[c.prediction_type]
        s.clip_sample = int(c.clip_sample)
        s.clip_sample_range = float(c.clip_sample_range)
        s.emulate_bf16 = int(emulate_bf16)
        return s
