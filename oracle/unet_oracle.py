"""Functional CPU restatement of the reference UNet eps-model (test infrastructure).

Everything here follows ``/root/reference/src/UNet.py`` and is written against a
flat ``state_dict`` (the 200-tensor contract of SURVEY.md App. B-5) instead of
an ``nn.Module`` tree, so that it can be run in fp32 or fp64 on any host.

Reference anchors (file:line in /root/reference):
  SinusoidalPosEmb     src/UNet.py:23-44
  Block                src/UNet.py:47-58
  ResNetBlock          src/UNet.py:61-99
  PreNorm / Residual   src/UNet.py:14-20,102-110
  Attention            src/UNet.py:113-136
  LinearAttention      src/UNet.py:139-164
  Encoder / Decoder    src/UNet.py:167-248
  TimeEmbedding        src/UNet.py:251-273
  BottleNeck           src/UNet.py:276-290   (time embedding is NOT used there)
  UNet.forward         src/UNet.py:361-389
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

HEADS = 4
DIM_HEAD = 32
HIDDEN = HEADS * DIM_HEAD


# --------------------------------------------------------------------------
# parameter inventory (creation order == reference construction order, so the
# same torch seed gives the same default initialisation as the reference)
# --------------------------------------------------------------------------
def _resblock_spec(prefix: str, cin: int, cout: int, temb: Optional[int]):
    spec = []
    if temb is not None:
        spec += [(f"{prefix}.mlp_t.1.weight", (cout, temb), "linear_w"),
                 (f"{prefix}.mlp_t.1.bias", (cout,), "bias")]
    spec += [(f"{prefix}.block1.norm.weight", (cin,), "ones"),
             (f"{prefix}.block1.norm.bias", (cin,), "zeros"),
             (f"{prefix}.block1.conv2d.weight", (cout, cin, 3, 3), "conv_w"),
             (f"{prefix}.block1.conv2d.bias", (cout,), "bias"),
             (f"{prefix}.block2.norm.weight", (cout,), "ones"),
             (f"{prefix}.block2.norm.bias", (cout,), "zeros"),
             (f"{prefix}.block2.conv2d.weight", (cout, cout, 3, 3), "conv_w"),
             (f"{prefix}.block2.conv2d.bias", (cout,), "bias")]
    if cin != cout:
        spec += [(f"{prefix}.shortcut.weight", (cout, cin, 1, 1), "conv_w"),
                 (f"{prefix}.shortcut.bias", (cout,), "bias")]
    return spec


def _linattn_spec(prefix: str, dim: int):
    # Residual(PreNorm(dim, LinearAttention(dim))): fn.fn = attention, fn.norm = prenorm
    return [(f"{prefix}.fn.fn.to_qkv.weight", (3 * HIDDEN, dim, 1, 1), "conv_w"),
            (f"{prefix}.fn.fn.to_out.0.weight", (dim, HIDDEN, 1, 1), "conv_w"),
            (f"{prefix}.fn.fn.to_out.0.bias", (dim,), "bias"),
            (f"{prefix}.fn.fn.to_out.1.weight", (dim,), "ones"),
            (f"{prefix}.fn.fn.to_out.1.bias", (dim,), "zeros"),
            (f"{prefix}.fn.norm.weight", (dim,), "ones"),
            (f"{prefix}.fn.norm.bias", (dim,), "zeros")]


def _attn_spec(prefix: str, dim: int):
    return [(f"{prefix}.fn.fn.to_qkv.weight", (3 * HIDDEN, dim, 1, 1), "conv_w"),
            (f"{prefix}.fn.fn.to_out.weight", (dim, HIDDEN, 1, 1), "conv_w"),
            (f"{prefix}.fn.fn.to_out.bias", (dim,), "bias"),
            (f"{prefix}.fn.norm.weight", (dim,), "ones"),
            (f"{prefix}.fn.norm.bias", (dim,), "zeros")]


def unet_key_shapes(in_channels: int, out_channels: int, channels: int = 64,
                    channel_multipliers: Sequence[int] = (1, 2, 4, 8),
                    with_time_emb: bool = True, num_classes: Optional[int] = None):
    """(key, shape, init-kind) triples in reference state_dict order (src/UNet.py:293-348)."""
    dims = [channels] + [channels * m for m in channel_multipliers]
    temb = channels * 4 if with_time_emb else None
    spec = []
    if with_time_emb:
        spec += [("time_emb.time_mlp.1.weight", (temb, temb // 4), "linear_w"),
                 ("time_emb.time_mlp.1.bias", (temb,), "bias"),
                 ("time_emb.time_mlp.3.weight", (temb, temb), "linear_w"),
                 ("time_emb.time_mlp.3.bias", (temb,), "bias")]
    if num_classes is not None:
        spec += [("label_emb.weight", (num_classes, temb), "normal")]
    spec += [("initial_conv.weight", (channels, in_channels, 3, 3), "conv_w"),
             ("initial_conv.bias", (channels,), "bias")]
    for i in range(len(dims) - 1):
        spec += _resblock_spec(f"encoder.downs.{i}.0", dims[i], dims[i + 1], temb)
        spec += _linattn_spec(f"encoder.downs.{i}.1", dims[i + 1])
    c = dims[-1]
    spec += _resblock_spec("bottleneck.res1", c, c, temb)
    spec += _attn_spec("bottleneck.attn", c)
    spec += _resblock_spec("bottleneck.res2", c, c, temb)
    rd = list(reversed(dims))
    for i in range(len(rd) - 1):
        spec += _resblock_spec(f"decoder.ups.{i}.0", rd[i] + rd[i + 1], rd[i + 1], temb)
        spec += _linattn_spec(f"decoder.ups.{i}.1", rd[i + 1])
        spec += [(f"decoder.ups.{i}.2.weight", (rd[i], rd[i + 1], 2, 2), "conv_w"),
                 (f"decoder.ups.{i}.2.bias", (rd[i + 1],), "bias")]
    spec += _resblock_spec("final_conv.0", channels, channels, None)
    spec += [("final_conv.1.weight", (out_channels, channels, 1, 1), "conv_w"),
             ("final_conv.1.bias", (out_channels,), "bias")]
    return spec


def init_state_dict(seed: int, in_channels: int, out_channels: int, channels: int = 64,
                    channel_multipliers: Sequence[int] = (1, 2, 4, 8),
                    with_time_emb: bool = True, num_classes: Optional[int] = None,
                    ) -> Dict[str, torch.Tensor]:
    """PyTorch-default initialisation drawn in the reference's construction order.

    nn.Linear / nn.Conv2d / nn.ConvTranspose2d: kaiming_uniform_(a=sqrt(5)) on the
    weight, then U(-1/sqrt(fan_in), 1/sqrt(fan_in)) on the bias; nn.Embedding:
    N(0,1); nn.GroupNorm: ones / zeros.  With the same ``torch.manual_seed`` this
    reproduces ``UNet(...)`` of the reference bit for bit (checked in tests).
    """
    torch.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    last_fan_in = 1
    for key, shape, kind in unet_key_shapes(in_channels, out_channels, channels,
                                            channel_multipliers, with_time_emb, num_classes):
        t = torch.empty(shape, dtype=torch.float32)
        if kind in ("conv_w", "linear_w"):
            torch.nn.init.kaiming_uniform_(t, a=math.sqrt(5))
            last_fan_in = t[0].numel()  # size(1) * receptive field (also right for ConvTranspose IOHW)
        elif kind == "bias":
            bound = 1.0 / math.sqrt(last_fan_in)
            torch.nn.init.uniform_(t, -bound, bound)
        elif kind == "normal":
            torch.nn.init.normal_(t)
        elif kind == "ones":
            t.fill_(1.0)
        elif kind == "zeros":
            t.zero_()
        sd[key] = t
    return sd


# --------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------
def sinusoidal_pos_emb(t: torch.Tensor, dim: int, dtype) -> torch.Tensor:
    """src/UNet.py:32-44.  The frequency table is built in fp32 in the reference."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, device=t.device) * -e)  # fp32, as the reference
    arg = t[:, None].to(freqs.dtype) * freqs[None, :]
    if dtype == torch.float64:  # fp64 oracle: keep the fp32-rounded table, evaluate sin/cos wide
        arg = t[:, None].to(torch.float64) * freqs[None, :].to(torch.float64)
    return torch.cat((arg.sin(), arg.cos()), dim=-1).to(dtype)


def time_embedding(sd, t: torch.Tensor, dtype) -> torch.Tensor:
    """src/UNet.py:263-273: sinusoid -> Linear -> GELU(erf) -> Linear."""
    w1, b1 = sd["time_emb.time_mlp.1.weight"], sd["time_emb.time_mlp.1.bias"]
    w3, b3 = sd["time_emb.time_mlp.3.weight"], sd["time_emb.time_mlp.3.bias"]
    e = sinusoidal_pos_emb(t, w1.shape[1], dtype)
    return F.linear(F.gelu(F.linear(e, w1, b1)), w3, b3)


def block(sd, p: str, x: torch.Tensor, groups: int = 8) -> torch.Tensor:
    """src/UNet.py:47-58: conv3x3(SiLU(GroupNorm(8, C)(x)))."""
    h = F.group_norm(x, groups, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], eps=1e-5)
    return F.conv2d(F.silu(h), sd[f"{p}.conv2d.weight"], sd[f"{p}.conv2d.bias"], padding=1)


def resnet_block(sd, p: str, x: torch.Tensor, temb: Optional[torch.Tensor]) -> torch.Tensor:
    """src/UNet.py:85-99."""
    h = block(sd, f"{p}.block1", x)
    if temb is not None and f"{p}.mlp_t.1.weight" in sd:
        proj = F.linear(F.silu(temb), sd[f"{p}.mlp_t.1.weight"], sd[f"{p}.mlp_t.1.bias"])
        h = proj[:, :, None, None] + h
    h = block(sd, f"{p}.block2", h)
    if f"{p}.shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{p}.shortcut.weight"], sd[f"{p}.shortcut.bias"])
    return h + x


def _split_heads(qkv: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    # "b (h c) x y -> b h c (x y)" on each contiguous third (src/UNet.py:124-127,150-153)
    b, _, hh, ww = qkv.shape
    q, k, v = qkv.chunk(3, dim=1)
    f = lambda z: z.reshape(b, HEADS, DIM_HEAD, hh * ww)
    return f(q), f(k), f(v)


def linear_attention(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """src/UNet.py:149-164 (p = '<...>.fn.fn')."""
    b, c, hh, ww = x.shape
    q, k, v = _split_heads(F.conv2d(x, sd[f"{p}.to_qkv.weight"]))
    q = q.softmax(dim=-2) * DIM_HEAD ** -0.5
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b, HIDDEN, hh, ww)
    out = F.conv2d(out, sd[f"{p}.to_out.0.weight"], sd[f"{p}.to_out.0.bias"])
    return F.group_norm(out, 1, sd[f"{p}.to_out.1.weight"], sd[f"{p}.to_out.1.bias"], eps=1e-5)


def attention(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """src/UNet.py:122-136 (p = '<...>.fn.fn')."""
    b, c, hh, ww = x.shape
    q, k, v = _split_heads(F.conv2d(x, sd[f"{p}.to_qkv.weight"]))
    q = q * DIM_HEAD ** -0.5
    sim = torch.einsum("bhdi,bhdj->bhij", q, k)
    sim = sim - sim.amax(dim=-1, keepdim=True)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhdj->bhid", attn, v)            # [b,h,n,d]
    out = out.permute(0, 1, 3, 2).reshape(b, HIDDEN, hh, ww)  # "b h (x y) d -> b (h d) x y"
    return F.conv2d(out, sd[f"{p}.to_out.weight"], sd[f"{p}.to_out.bias"])


def prenorm_residual(sd, p: str, x: torch.Tensor, fn) -> torch.Tensor:
    """Residual(PreNorm(dim, fn)) -- src/UNet.py:14-20,102-110 (p = '<...>' owning .fn.norm)."""
    h = F.group_norm(x, 1, sd[f"{p}.fn.norm.weight"], sd[f"{p}.fn.norm.bias"], eps=1e-5)
    return fn(sd, f"{p}.fn.fn", h) + x


def unet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, t: torch.Tensor,
                 y: Optional[torch.Tensor] = None, *, taps: Optional[dict] = None) -> torch.Tensor:
    """eps = UNet(x_noisy, t, y) -- src/UNet.py:361-389.  dtype follows ``x``.

    ``sd`` must hold tensors of x's dtype (use ``cast_state_dict``).  ``taps`` (optional
    dict) receives named intermediates for kernel-level parity tests.
    """
    dtype = x.dtype
    n_levels = sum(1 for k in sd if k.startswith("encoder.downs.") and k.endswith(".0.block1.conv2d.weight"))
    temb = None
    if "time_emb.time_mlp.1.weight" in sd:
        temb = time_embedding(sd, t, dtype)
        if y is not None:
            temb = temb + sd["label_emb.weight"][y]  # length-1 y broadcasts (src/UNet.py:375-376)
    if taps is not None:
        taps["temb"] = temb
    h = F.conv2d(x, sd["initial_conv.weight"], sd["initial_conv.bias"], padding=1)
    if taps is not None:
        taps["initial"] = h
    skips: List[torch.Tensor] = []
    for i in range(n_levels):
        h = resnet_block(sd, f"encoder.downs.{i}.0", h, temb)
        if taps is not None:
            taps[f"enc{i}.res"] = h
        h = prenorm_residual(sd, f"encoder.downs.{i}.1", h, linear_attention)
        if taps is not None:
            taps[f"enc{i}.attn"] = h
        skips.append(h)
        h = F.max_pool2d(h, 2, 2)
    # BottleNeck.forward ignores t (src/UNet.py:287-288)
    h = resnet_block(sd, "bottleneck.res1", h, None)
    h = prenorm_residual(sd, "bottleneck.attn", h, attention)
    h = resnet_block(sd, "bottleneck.res2", h, None)
    if taps is not None:
        taps["bottleneck"] = h
    for i in range(n_levels):
        h = F.conv_transpose2d(h, sd[f"decoder.ups.{i}.2.weight"], sd[f"decoder.ups.{i}.2.bias"], stride=2)
        h = torch.cat((h, skips.pop()), dim=1)
        h = resnet_block(sd, f"decoder.ups.{i}.0", h, temb)
        h = prenorm_residual(sd, f"decoder.ups.{i}.1", h, linear_attention)
        if taps is not None:
            taps[f"dec{i}"] = h
    h = resnet_block(sd, "final_conv.0", h, None)
    return F.conv2d(h, sd["final_conv.1.weight"], sd["final_conv.1.bias"])


def cast_state_dict(sd: Dict[str, torch.Tensor], dtype=None, device=None) -> Dict[str, torch.Tensor]:
    return {k: v.to(dtype=dtype, device=device) for k, v in sd.items()}
