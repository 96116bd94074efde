"""CPU restatement of the reference diffusion process (test infrastructure).

Follows ``/root/reference/src/DDPM.py``:
  schedule   :31-43   beta = linspace(1e-4, 0.02, T) fp32, alpha = 1-beta, alpha_bar = cumprod
  gather     :12-19
  q_sample   :46-68   x_t = sqrt(abar_t) x0 + sqrt(1-abar_t) eps
  p_sample   :71-96   mean = alpha_t^-1/2 (x_t - (1-alpha_t)/sqrt(1-abar_t) eps_theta); + sqrt(beta_t) z unless t[0]==0
  sample     :98-130  reverse loop with classifier-free guidance via torch.lerp
and the LatentDiffusionModel schedule, ``src/LatentDiffusionModel.py:41-55``.

All RNG is injected (``noise`` arguments): the north-star compares on fixed noise.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch


def make_schedule(n_steps: int) -> Dict[str, torch.Tensor]:
    """src/DDPM.py:31-43 (all fp32, exactly the reference's op sequence)."""
    beta = torch.linspace(0.0001, 0.02, n_steps)
    alpha = 1.0 - beta
    alpha_bar = torch.cumprod(alpha, dim=0)
    return {"beta": beta, "alpha": alpha, "alpha_bar": alpha_bar, "sigma2": beta}


def make_ldm_schedule(n_steps: int, linear_start: float, linear_end: float) -> Dict[str, torch.Tensor]:
    """src/LatentDiffusionModel.py:41-55: sqrt-linear beta in fp64, cast to fp32."""
    beta = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_steps, dtype=torch.float64) ** 2
    alpha = 1.0 - beta
    alpha_bar = torch.cumprod(alpha, dim=0)
    return {"beta": beta.to(torch.float32), "alpha_bar": alpha_bar.to(torch.float32)}


def _g(v: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    return v.gather(-1, t).reshape(-1, 1, 1, 1)  # src/DDPM.py:12-19


def q_sample(sched, x0: torch.Tensor, t: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """src/DDPM.py:46-68."""
    abar = _g(sched["alpha_bar"].to(x0.dtype), t)
    mean = abar ** 0.5 * x0
    var = 1 - abar
    return mean + (var ** 0.5) * eps


def p_sample(sched, xt: torch.Tensor, t: torch.Tensor, eps_theta: torch.Tensor,
             noise: Optional[torch.Tensor]) -> torch.Tensor:
    """src/DDPM.py:71-96; ``noise`` replaces the reference's torch.randn draw."""
    dt = xt.dtype
    alpha_bar = _g(sched["alpha_bar"].to(dt), t)
    alpha = _g(sched["alpha"].to(dt), t)
    eps_coef = (1 - alpha) / (1 - alpha_bar) ** 0.5
    mean = 1 / (alpha ** 0.5) * (xt - eps_coef * eps_theta)
    if int(t[0]) == 0:  # the reference branches on t[0] for the whole batch (:85)
        return mean
    var = _g(sched["sigma2"].to(dt), t)
    return mean + (var ** 0.5) * noise


def cfg_combine(eps_cond: torch.Tensor, eps_uncond: torch.Tensor, cfg_scale: float) -> torch.Tensor:
    """torch.lerp(uncond, cond, w) = uncond + w (cond - uncond)  -- src/DDPM.py:124."""
    return torch.lerp(eps_uncond, eps_cond, cfg_scale)


def sample_loop(sched, eps_model: Callable, classes, x_T: torch.Tensor, noises,
                cfg_scale: float = 3.0, n_steps: Optional[int] = None,
                on_step: Optional[Callable] = None) -> torch.Tensor:
    """src/DDPM.py:98-130 with injected x_T and per-step noise.

    ``noises`` is indexable by the timestep t (shape of x_T); ``noises[t]`` is used at step t>0.
    ``eps_model(x, t_long[B], classes_or_None)`` is the reference's duck-typed protocol.
    """
    T = n_steps if n_steps is not None else sched["beta"].numel()
    xt = x_T
    B = xt.shape[0]
    for step in reversed(range(T)):
        t = torch.full((B,), step, dtype=torch.long, device=xt.device)
        eps = eps_model(xt, t, classes)
        if cfg_scale > 0:
            eps_u = eps_model(xt, t, None)
            eps = cfg_combine(eps, eps_u, cfg_scale)
        z = noises[step] if step > 0 else None
        xt = p_sample(sched, xt, t, eps, z)
        if on_step is not None:
            on_step(step, xt)
    return xt
