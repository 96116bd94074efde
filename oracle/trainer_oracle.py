"""CPU oracle for the steps either side of the denoising hot path.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py): the optimizer step, the validation loss with classifier-free guidance and the uint8 output
stage, restated with plain tensor / numpy arithmetic.  Pinned by tests/golden/g7_adam.npz, g8_output.npz and
g9_val_loss.npz, which oracle/make_golden_next.py produced by running the reference's own call sites
(torch.optim.Adam as src/Trainer.py:68-71 builds it, src/transforms.py:22-35, src/utils.py:121-130 ->
torchvision.utils.save_image, src/DiffusionModelTrainer.py:94-107)."""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .ddpm_oracle import q_sample, cfg_combine
from .unet_oracle import unet_forward


def adam_step(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int,
              lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
    """One torch.optim.Adam update with default flags, in place (src/Trainer.py:68-71: Adam(params, lr=cfg.lr)).
    m += (1-b1)(g-m);  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)."""
    exp_avg += (1.0 - beta1) * (grad - exp_avg)
    exp_avg_sq *= beta2
    exp_avg_sq += (1.0 - beta2) * grad * grad
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = exp_avg_sq.sqrt() / math.sqrt(bc2) + eps
    param -= (lr / bc1) * (exp_avg / denom)


def images_to_uint8(x_nchw: np.ndarray, convention: str) -> np.ndarray:
    """fp32 [B,C,H,W] -> uint8 [B,H,W,C].
    'save_image'        : torchvision.utils.save_image on the raw tensor (src/utils.py:121-130): x*255 + 0.5, clamp
                          to [0,255], truncate.
    'reverse_transform' : src/transforms.py:22-35: ((x+1)/2)*255, numpy astype(uint8) (truncate, keep the low byte)."""
    x = np.asarray(x_nchw, dtype=np.float32)
    if convention == "save_image":
        v = np.clip(x * np.float32(255.0) + np.float32(0.5), 0, 255)
        out = np.trunc(v).astype(np.int64)
    elif convention == "reverse_transform":
        v = ((x + np.float32(1.0)) / np.float32(2.0)) * np.float32(255.0)
        out = np.trunc(v).astype(np.int64) & 255
    else:
        raise ValueError(convention)
    return np.ascontiguousarray(out.astype(np.uint8).transpose(0, 2, 3, 1))


def val_loss(sd: Dict[str, torch.Tensor], sched, x0: torch.Tensor, noise: torch.Tensor, t: torch.Tensor,
             y: Optional[torch.Tensor], cfg_scale: float) -> torch.Tensor:
    """The loss of one `_val_epoch` batch (src/DiffusionModelTrainer.py:94-107) for given noise and t."""
    xt = q_sample(sched, x0, t, noise)
    eps = unet_forward(sd, xt, t, y)
    if cfg_scale > 0:
        eps = cfg_combine(eps, unet_forward(sd, xt, t, None), cfg_scale)
    return torch.nn.functional.mse_loss(noise, eps)
