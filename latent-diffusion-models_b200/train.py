"""Training-step path (autograd bridge).  Filled in by the backward kernels; see DESIGN.md."""
from __future__ import annotations


def unet_autograd_forward(model, x, t, y):
    raise NotImplementedError(
        "ldm_b200.UNet: the backward kernels are not built into this library; call under torch.no_grad() "
        "for eps-prediction / sampling")
