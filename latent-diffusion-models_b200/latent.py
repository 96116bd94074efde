"""Drop-in ``LatentDiffusionModel`` / ``DiffusionWrapper`` (reference: src/LatentDiffusionModel.py:7-81).

The class surface and state_dict stay as the reference's: ``model.diffusion_model.*`` (the UNet),
``autoencoder.*`` and the frozen ``beta`` / ``alpha_bar`` parameters of the sqrt-linear schedule.  The
eps-prediction is the native UNet on the latent shape; ``Diffusion.sample`` recognises the wrapper and
runs the graph sampler on the wrapped UNet.  ``autoencoder_decode`` routes around the reference's
``self.first_stage_model`` AttributeError (src/LatentDiffusionModel.py:72) by using ``self.autoencoder``.
"""
from __future__ import annotations

import torch
from torch import nn


class DiffusionWrapper(nn.Module):
    def __init__(self, diffusion_model: nn.Module):
        super().__init__()
        self.diffusion_model = diffusion_model

    def forward(self, x: torch.Tensor, time_steps: torch.Tensor, targets: torch.Tensor = None):
        return self.diffusion_model(x, time_steps, targets)


class LatentDiffusionModel(nn.Module):
    def __init__(self, eps_model, autoencoder, latent_scaling_factor: float, n_steps: int,
                 linear_start: float, linear_end: float):
        super().__init__()
        self.model = DiffusionWrapper(eps_model)
        self.autoencoder = autoencoder
        self.latent_scaling_factor = latent_scaling_factor
        self.n_steps = n_steps
        # sqrt-linear schedule computed in fp64 and stored as frozen fp32 parameters (:41-55)
        beta = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_steps, dtype=torch.float64) ** 2
        self.beta = nn.Parameter(beta.to(torch.float32), requires_grad=False)
        alpha_bar = torch.cumprod(1.0 - beta, dim=0)
        self.alpha_bar = nn.Parameter(alpha_bar.to(torch.float32), requires_grad=False)

    @property
    def num_classes(self):
        return getattr(self.model.diffusion_model, "num_classes", None)

    def make_diffusion(self, device):
        """A Diffusion process carrying this model's schedule (the reference never wires one up)."""
        from .ddpm import Diffusion
        d = Diffusion(self.n_steps, device)
        d.set_schedule(self.beta.data, self.alpha_bar.data)
        return d

    def autoencoder_encode(self, image: torch.Tensor, epsilon: torch.Tensor = None):
        """Scaled latent of the image (:57-65).  ``epsilon`` (keyword, ours) fixes the reparameterisation noise for parity
        tests; it is only passed on when given, so any autoencoder with the reference's ``encode(image)`` works."""
        dist = self.autoencoder.encode(image) if epsilon is None else self.autoencoder.encode(image, epsilon)
        return self.latent_scaling_factor * dist.sample()

    def autoencoder_decode(self, z: torch.Tensor):
        return self.autoencoder.decode(z / self.latent_scaling_factor)

    def forward(self, x: torch.Tensor, t: torch.Tensor, targets: torch.Tensor = None):
        return self.model(x, t, targets)
