"""ldm_b200: the DDPM denoising hot path of JohanLundberg12/latent-diffusion-models on sm_100a.

Drop-in classes (same names, constructor arguments, methods and state_dict as the reference's
src/UNet.py, src/DDPM.py, src/LatentDiffusionModel.py and src/Autoencoder.py); all tensor work happens in
``libldm_b200.so`` (hand-written CUDA for B200) through the C ABI of ``include/ldm_b200.h``.
"""
from . import _lib  # noqa: F401
from .unet import UNet  # noqa: F401
from .ddpm import Diffusion  # noqa: F401
from .latent import LatentDiffusionModel, DiffusionWrapper  # noqa: F401
from .autoencoder import Autoencoder, GaussianDistribution  # noqa: F401

__all__ = ["UNet", "Diffusion", "LatentDiffusionModel", "DiffusionWrapper", "Autoencoder", "GaussianDistribution"]
