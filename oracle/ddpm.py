"""Builder-owned DDPM oracle.  TEST INFRASTRUCTURE ONLY.

**PARITY UNPINNED — NOT FROM THE REFERENCE.**  The reference repository contains no diffusion
code (its only trace is an ignored launcher name at ``/root/reference/src/.gitignore:10``).
BASELINE.json's north star nevertheless names "the diffusion denoiser's noising, timestep
embedding and reverse-sampling loop", so the spec below (frozen in DESIGN.md)
is the builder's, restated in plain torch fp32/fp64 on CPU.  The CUDA kernels are checked against
THIS file; nothing here can be checked against the reference.

Spec (standard DDPM, Ho et al. 2020, x0-prediction as in MDM):
  betas      = linspace(1e-4, 2e-2, T=1000) in fp64
  abar       = cumprod(1 - betas);  abar_prev = [1, abar[:-1]]
  q_sample   : x_t = sqrt(abar_t) x0 + sqrt(1 - abar_t) eps
  posterior  : mu = c1_t x0_hat + c2_t x_t,  c1 = beta sqrt(abar_prev)/(1-abar),
               c2 = (1-abar_prev) sqrt(alpha)/(1-abar)
               x_{t-1} = mu + [t>0] exp(0.5 logvar_t) z,  logvar = log(max(post_var, post_var[1]))
  all tables computed in fp64 and stored as fp32; elementwise math in fp32.
"""
from __future__ import annotations

from typing import Callable, Dict

import numpy as np
import torch

NUM_TIMESTEPS = 1000
BETA_START = 1e-4
BETA_END = 2e-2


def make_schedule(num_timesteps: int = NUM_TIMESTEPS, beta_start: float = BETA_START,
                  beta_end: float = BETA_END) -> Dict[str, torch.Tensor]:
    betas = np.linspace(beta_start, beta_end, num_timesteps, dtype=np.float64)
    alphas = 1.0 - betas
    abar = np.cumprod(alphas)
    abar_prev = np.concatenate([[1.0], abar[:-1]])
    post_var = betas * (1.0 - abar_prev) / (1.0 - abar)
    post_logvar = np.log(np.concatenate([post_var[1:2], post_var[1:]]))
    tab = dict(
        sqrt_abar=np.sqrt(abar), sqrt_one_minus_abar=np.sqrt(1.0 - abar),
        coef_x0=betas * np.sqrt(abar_prev) / (1.0 - abar),
        coef_xt=(1.0 - abar_prev) * np.sqrt(alphas) / (1.0 - abar),
        sigma=np.exp(0.5 * post_logvar),
    )
    return {k: torch.from_numpy(v.astype(np.float32)) for k, v in tab.items()}


def q_sample(sched, x0: torch.Tensor, t: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    a = sched["sqrt_abar"][t].view(-1, 1, 1)
    b = sched["sqrt_one_minus_abar"][t].view(-1, 1, 1)
    return a * x0 + b * eps


def posterior_step(sched, x0_hat: torch.Tensor, x_t: torch.Tensor, t: int, z: torch.Tensor) -> torch.Tensor:
    mu = sched["coef_x0"][t] * x0_hat + sched["coef_xt"][t] * x_t
    if t > 0:
        return mu + sched["sigma"][t] * z
    return mu


def sample_loop(sched, denoise: Callable[[torch.Tensor, int], torch.Tensor], x_T: torch.Tensor,
                noise: Callable[[int], torch.Tensor], steps=None) -> torch.Tensor:
    """x_T → x_0 with the given per-step noise supplier (parity tests feed identical noise to the
    CUDA path).  ``steps`` defaults to all T steps (999 … 0)."""
    x = x_T
    ts = range(len(sched["sigma"]) - 1, -1, -1) if steps is None else steps
    for t in ts:
        x = posterior_step(sched, denoise(x, t), x, t, noise(t))
    return x
