"""CPU oracle for the first-stage autoencoder.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
A functional restatement of src/Autoencoder.py over a plain state_dict; pinned by tests/golden/g10_autoencoder_*.npz, which
oracle/make_golden_next.py produced by running the unmodified reference modules."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn.functional as F


def _gn(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)       # src/Autoencoder.py:9-11


def _swish(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(x)                                                   # :14-18


def _conv(sd, p: str, x: torch.Tensor, stride: int = 1, padding: int = 0) -> torch.Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def resnet_block(sd, p: str, x: torch.Tensor) -> torch.Tensor:                    # :46-84
    h = _conv(sd, p + ".conv1", _swish(_gn(sd, p + ".norm1", x)), padding=1)
    h = _conv(sd, p + ".conv2", _swish(_gn(sd, p + ".norm2", h)), padding=1)
    sc = _conv(sd, p + ".nin_shortcut", x) if (p + ".nin_shortcut.weight") in sd else x
    return sc + h


def attn_block(sd, p: str, x: torch.Tensor) -> torch.Tensor:                      # :87-139
    xn = _gn(sd, p + ".norm", x)
    b, c, h, w = x.shape
    q = _conv(sd, p + ".q", xn).reshape(b, c, h * w)
    k = _conv(sd, p + ".k", xn).reshape(b, c, h * w)
    v = _conv(sd, p + ".v", xn).reshape(b, c, h * w)
    attn = torch.softmax(torch.einsum("bci,bcj->bij", q, k) * (c ** -0.5), dim=2)
    out = torch.einsum("bij,bcj->bci", attn, v).reshape(b, c, h, w)
    return x + _conv(sd, p + ".proj_out", out)


def _count(sd, prefix: str) -> int:
    n = 0
    while any(k.startswith(f"{prefix}.{n}.") for k in sd):
        n += 1
    return n


def encoder(sd, img: torch.Tensor, p: str = "encoder") -> torch.Tensor:          # :183-291
    x = _conv(sd, p + ".conv_in", img, padding=1)
    for i in range(_count(sd, p + ".down")):
        for j in range(_count(sd, f"{p}.down.{i}.block")):
            x = resnet_block(sd, f"{p}.down.{i}.block.{j}", x)
        if f"{p}.down.{i}.downsample.conv.weight" in sd:                          # :160-180
            x = _conv(sd, f"{p}.down.{i}.downsample.conv", F.pad(x, (0, 1, 0, 1)), stride=2)
    x = resnet_block(sd, p + ".mid.block_1", x)
    x = attn_block(sd, p + ".mid.attn_1", x)
    x = resnet_block(sd, p + ".mid.block_2", x)
    return _conv(sd, p + ".conv_out", _swish(_gn(sd, p + ".norm_out", x)), padding=1)


def decoder(sd, z: torch.Tensor, p: str = "decoder") -> torch.Tensor:            # :294-385
    h = _conv(sd, p + ".conv_in", z, padding=1)
    h = resnet_block(sd, p + ".mid.block_1", h)
    h = attn_block(sd, p + ".mid.attn_1", h)
    h = resnet_block(sd, p + ".mid.block_2", h)
    for i in reversed(range(_count(sd, p + ".up"))):
        for j in range(_count(sd, f"{p}.up.{i}.block")):
            h = resnet_block(sd, f"{p}.up.{i}.block.{j}", h)
        if f"{p}.up.{i}.upsample.conv.weight" in sd:                              # :142-157
            h = _conv(sd, f"{p}.up.{i}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"), padding=1)
    return _conv(sd, p + ".conv_out", _swish(_gn(sd, p + ".norm_out", h)), padding=1)


def encode_moments(sd, img: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(mu, log_var) of Autoencoder.encode (:432-440)."""
    mu, log_var = torch.chunk(_conv(sd, "quant_conv", encoder(sd, img)), 2, dim=1)
    return mu, log_var


def gaussian_sample(mu: torch.Tensor, log_var: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:   # :21-43
    return mu + torch.exp(log_var / 2) * eps


def decode(sd, z: torch.Tensor) -> torch.Tensor:                                  # :442-450
    return decoder(sd, _conv(sd, "post_quant_conv", z))
