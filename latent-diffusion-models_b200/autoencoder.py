"""Drop-in ``Autoencoder`` (reference: src/Autoencoder.py:183-462) on the C ABI: the first stage of the
LatentDiffusionModel (BASELINE config 4: encode -> latent-space DDPM sampling -> decode).

Same constructor arguments, attribute tree and ``state_dict`` keys as the reference (``encoder.down.{i}.block.{j}.norm1.weight``,
``encoder.mid.attn_1.q.weight``, ``quant_conv.*``, ``decoder.up.{i}.upsample.conv.*`` ...).  The ``torch.nn`` modules below are
parameter HOLDERS only -- created in the reference's order, so ``torch.manual_seed(s); Autoencoder(...)`` draws the same
default-init weights -- and are never called: every tensor op runs in ``libldm_b200.so`` on NHWC activations:

* 3x3 / 1x1 convolutions: the UNet's implicit-GEMM kernels (tcgen05 when both channel counts are multiples of 64, the FFMA
  kernel for the few narrow ones -- image / latent / moment channels are zero-padded to a multiple of 16);
* ``DownSample`` (pad (0,1,0,1), stride 2) = the pad-1 conv at full resolution followed by ``ldm_downsample_pick``;
  ``UpSample`` = ``ldm_upsample_nearest2x`` then the conv;
* GroupNorm(32, C, eps=1e-6) + swish: ``ldm_group_norm``;  ``AttnBlock``: q | k | v as ONE 1x1 conv, then
  ``ldm_attention_single_head``, then ``proj_out`` with the residual in its epilogue;
* ``GaussianDistribution``: ``ldm_gaussian_distribution``.

CUDA only: there is no CPU path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _lib, ops

_EPS = 1e-6
_GROUPS = 32


def _norm(ch: int) -> nn.GroupNorm:              # src/Autoencoder.py:9-11
    return nn.GroupNorm(num_groups=_GROUPS, num_channels=ch, eps=_EPS)


class ResnetBlock(nn.Module):                    # :46-84
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.norm1 = _norm(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, stride=1, padding=1)
        self.norm2 = _norm(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, stride=1, padding=1)
        self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else nn.Identity()


class AttnBlock(nn.Module):                      # :87-139
    def __init__(self, channels: int):
        super().__init__()
        self.norm = _norm(channels)
        self.q = nn.Conv2d(channels, channels, 1)
        self.k = nn.Conv2d(channels, channels, 1)
        self.v = nn.Conv2d(channels, channels, 1)
        self.proj_out = nn.Conv2d(channels, channels, 1)
        self.scale = channels ** -0.5


class UpSample(nn.Module):                       # :142-157
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)


class DownSample(nn.Module):                     # :160-180
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=0)


class Encoder(nn.Module):                        # :183-291
    def __init__(self, channels: int = 64, channel_multipliers: List[int] = (1, 2, 4, 8), n_resnet_blocks: int = 2,
                 in_channels: int = 1, z_channels: int = 512):
        super().__init__()
        n_res = len(channel_multipliers)
        self.conv_in = nn.Conv2d(in_channels, channels, 3, stride=1, padding=1)
        chans = [m * channels for m in [1] + list(channel_multipliers)]
        self.down = nn.ModuleList()
        for i in range(n_res):
            blocks = nn.ModuleList()
            for _ in range(n_resnet_blocks):
                blocks.append(ResnetBlock(channels, chans[i + 1]))
                channels = chans[i + 1]
            down = nn.Module()
            down.block = blocks
            down.downsample = DownSample(channels) if i != n_res - 1 else nn.Identity()
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(channels, channels)
        self.mid.attn_1 = AttnBlock(channels)
        self.mid.block_2 = ResnetBlock(channels, channels)
        self.norm_out = _norm(channels)
        self.conv_out = nn.Conv2d(channels, 2 * z_channels, 3, stride=1, padding=1)


class Decoder(nn.Module):                        # :294-385
    def __init__(self, channels: int = 64, channel_multipliers: List[int] = (1, 2, 4, 8), n_resnet_blocks: int = 2,
                 out_channels: int = 1, z_channels: int = 512):
        super().__init__()
        n_res = len(channel_multipliers)
        chans = [m * channels for m in channel_multipliers]
        channels = chans[-1]
        self.conv_in = nn.Conv2d(z_channels, channels, 3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(channels, channels)
        self.mid.attn_1 = AttnBlock(channels)
        self.mid.block_2 = ResnetBlock(channels, channels)
        self.up = nn.ModuleList()
        for i in reversed(range(n_res)):
            blocks = nn.ModuleList()
            for _ in range(n_resnet_blocks + 1):
                blocks.append(ResnetBlock(channels, chans[i]))
                channels = chans[i]
            up = nn.Module()
            up.block = blocks
            up.upsample = UpSample(channels) if i != 0 else nn.Identity()
            self.up.insert(0, up)
        self.norm_out = _norm(channels)
        self.conv_out = nn.Conv2d(channels, out_channels, 3, stride=1, padding=1)


class GaussianDistribution:                      # :21-43
    """mu, log_var, sigma, epsilon as fp32 NCHW tensors; ``sample()`` = mu + sigma * epsilon (computed in the same launch)."""

    def __init__(self, moments_nhwc: torch.Tensor, z_channels: int, epsilon: Optional[torch.Tensor] = None):
        B, H, W, ld = moments_nhwc.shape
        dev = moments_nhwc.device
        shape = (B, z_channels, H, W)
        self.epsilon = (torch.randn(shape, device=dev, dtype=torch.float32) if epsilon is None
                        else epsilon.detach().to(dev, torch.float32).contiguous())
        if tuple(self.epsilon.shape) != shape:
            raise ValueError(f"epsilon must have shape {shape}")
        self.mu, self.log_var, self.sigma, self._z = (torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(4))
        if self.mu.numel():
            with torch.cuda.device(dev):
                _lib.check(_lib.load().ldm_gaussian_distribution(
                    moments_nhwc.data_ptr(), moments_nhwc.stride(2), self.epsilon.data_ptr(), self.mu.data_ptr(),
                    self.log_var.data_ptr(), self.sigma.data_ptr(), self._z.data_ptr(), B, z_channels, H * W,
                    ops._dt(moments_nhwc), _lib.stream_ptr()))

    def sample(self) -> torch.Tensor:
        return self._z


def _pad16(n: int) -> int:
    return (n + 15) // 16 * 16


class Autoencoder(nn.Module):                    # :388-462
    def __init__(self, in_channels: int = 1, z_channels: int = 512, out_channels: int = 1, channels: int = 64,
                 channel_multipliers: List[int] = (1, 2, 4, 8), n_resnet_blocks: int = 2, *, dtype: str = "bf16"):
        super().__init__()
        if dtype not in ("bf16", "fp32"):
            raise ValueError("dtype must be 'bf16' or 'fp32'")
        self.z_channels, self.in_channels, self.out_channels = z_channels, in_channels, out_channels
        self.compute_dtype = dtype
        self.encoder = Encoder(channels, list(channel_multipliers), n_resnet_blocks, in_channels, z_channels)
        self.quant_conv = nn.Conv2d(z_channels * 2, z_channels * 2, 1)
        self.decoder = Decoder(channels, list(channel_multipliers), n_resnet_blocks, out_channels, z_channels)
        self.post_quant_conv = nn.Conv2d(z_channels, z_channels, 1)
        self._packed: Dict[Tuple, Tuple] = {}
        self.last_launches = 0

    # ------------------------------------------------------------------ packed weights (cached per parameter version)
    def _weights(self, convs: Tuple[nn.Conv2d, ...], cin_pad: int, cout_pad: int):
        """Pack (and zero-pad) the OIHW filters of one conv, or of several convs stacked along the output channels."""
        fp = tuple((c.weight.data_ptr(), c.weight._version, c.bias.data_ptr(), c.bias._version) for c in convs)
        key = (tuple(id(c) for c in convs), cin_pad, cout_pad, self.compute_dtype)
        hit = self._packed.get(key)
        if hit is not None and hit[0] == fp:
            return hit[1], hit[2]
        w = torch.cat([c.weight.detach() for c in convs]).to(torch.float32)
        b = torch.cat([c.bias.detach() for c in convs]).to(torch.float32)
        cout, cin, k, _ = w.shape
        if cin_pad != cin or cout_pad != cout:
            wp = torch.zeros(cout_pad, cin_pad, k, k, dtype=torch.float32, device=w.device)
            wp[:cout, :cin] = w
            bp = torch.zeros(cout_pad, dtype=torch.float32, device=w.device)
            bp[:cout] = b
            w, b = wp, bp
        packed = ops.pack_conv_weight(w.contiguous(), self.compute_dtype)
        self._packed[key] = (fp, packed, b.contiguous())
        return packed, b

    def _conv(self, x: torch.Tensor, convs, cout_pad: Optional[int] = None, res: Optional[torch.Tensor] = None) -> torch.Tensor:
        convs = convs if isinstance(convs, tuple) else (convs,)
        k = convs[0].kernel_size[0]
        cin = x.shape[3]
        cout = sum(c.out_channels for c in convs)
        cout_pad = cout_pad or cout
        wp, b = self._weights(convs, cin, cout_pad)
        tc = self.compute_dtype == "bf16" and cin % 64 == 0 and cout_pad % 64 == 0
        return ops.conv2d(x, wp, k, bias=b, res=res, impl=0 if tc else 1)

    def _gn(self, x: torch.Tensor, norm: nn.GroupNorm, silu: bool) -> torch.Tensor:
        return ops.group_norm(x, norm.weight.detach(), norm.bias.detach(), norm.num_groups, eps=norm.eps, silu=silu)

    def _resnet(self, x: torch.Tensor, blk: ResnetBlock) -> torch.Tensor:
        h = self._conv(self._gn(x, blk.norm1, True), blk.conv1)
        h = self._gn(h, blk.norm2, True)
        sc = x if isinstance(blk.nin_shortcut, nn.Identity) else self._conv(x, blk.nin_shortcut)
        return self._conv(h, blk.conv2, res=sc)          # nin_shortcut(x) + h in the conv epilogue

    def _attn(self, x: torch.Tensor, blk: AttnBlock) -> torch.Tensor:
        B, H, W, Cc = x.shape
        qkv = self._conv(self._gn(x, blk.norm, False), (blk.q, blk.k, blk.v))
        a = torch.empty(B, H, W, Cc, dtype=x.dtype, device=x.device)
        _lib.check(_lib.load().ldm_attention_single_head(qkv.data_ptr(), a.data_ptr(), B, H * W, Cc, ops._dt(x), _lib.stream_ptr()))
        return self._conv(a, blk.proj_out, res=x)

    def _up(self, x: torch.Tensor, up: UpSample) -> torch.Tensor:
        B, H, W, Cc = x.shape
        y = torch.empty(B, 2 * H, 2 * W, Cc, dtype=x.dtype, device=x.device)
        _lib.check(_lib.load().ldm_upsample_nearest2x(x.data_ptr(), x.stride(2), y.data_ptr(), Cc, B, H, W, Cc, ops._dt(x),
                                                      _lib.stream_ptr()))
        return self._conv(y, up.conv)

    def _down(self, x: torch.Tensor, down: DownSample) -> torch.Tensor:
        B, H, W, Cc = x.shape
        full = self._conv(x, down.conv)                  # pad-1, stride-1: its odd positions are the stride-2 outputs
        y = torch.empty(B, H // 2, W // 2, Cc, dtype=x.dtype, device=x.device)
        _lib.check(_lib.load().ldm_downsample_pick(full.data_ptr(), full.stride(2), y.data_ptr(), Cc, B, H, W, Cc, ops._dt(x),
                                                   _lib.stream_ptr()))
        return y

    def _to_nhwc_padded(self, x_nchw: torch.Tensor, channels: int) -> torch.Tensor:
        if not x_nchw.is_cuda:
            raise _lib.LdmError("ldm_b200.Autoencoder runs on CUDA tensors only (no CPU fallback)")
        if x_nchw.shape[1] != channels:
            raise ValueError(f"expected {channels} channels, got {x_nchw.shape[1]}")
        x = x_nchw.detach().to(torch.float32)
        pad = _pad16(channels) - channels
        if pad:
            x = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, pad))
        return ops.to_nhwc(x.contiguous(), self.compute_dtype)

    @staticmethod
    def _require_cuda(t: torch.Tensor) -> None:
        if not t.is_cuda:
            raise _lib.LdmError("ldm_b200.Autoencoder runs on CUDA tensors only (no CPU fallback)")

    # ------------------------------------------------------------------ reference surface
    @torch.no_grad()
    def _encode_moments(self, img: torch.Tensor) -> torch.Tensor:
        e = self.encoder
        self._require_cuda(img)
        with torch.cuda.device(img.device):
            x = self._conv(self._to_nhwc_padded(img, self.in_channels), e.conv_in)
            for down in e.down:
                for blk in down.block:
                    x = self._resnet(x, blk)
                if not isinstance(down.downsample, nn.Identity):
                    x = self._down(x, down.downsample)
            x = self._resnet(x, e.mid.block_1)
            x = self._attn(x, e.mid.attn_1)
            x = self._resnet(x, e.mid.block_2)
            x = self._conv(self._gn(x, e.norm_out, True), e.conv_out, cout_pad=_pad16(2 * self.z_channels))
            return self._conv(x, self.quant_conv, cout_pad=_pad16(2 * self.z_channels))

    @torch.no_grad()
    def encode(self, img: torch.Tensor, epsilon: Optional[torch.Tensor] = None) -> GaussianDistribution:
        self._require_cuda(img)
        before = _lib.launch_count()
        with torch.cuda.device(img.device):
            dist = GaussianDistribution(self._encode_moments(img), self.z_channels, epsilon)
        self.last_launches = _lib.launch_count() - before
        return dist

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        d = self.decoder
        self._require_cuda(z)
        before = _lib.launch_count()
        with torch.cuda.device(z.device):
            h = self._conv(self._to_nhwc_padded(z, self.z_channels), self.post_quant_conv, cout_pad=_pad16(self.z_channels))
            h = self._conv(h, d.conv_in)
            h = self._resnet(h, d.mid.block_1)
            h = self._attn(h, d.mid.attn_1)
            h = self._resnet(h, d.mid.block_2)
            for up in reversed(d.up):
                for blk in up.block:
                    h = self._resnet(h, blk)
                if not isinstance(up.upsample, nn.Identity):
                    h = self._up(h, up.upsample)
            cpad = (self.out_channels + 3) // 4 * 4
            img = self._conv(self._gn(h, d.norm_out, True), d.conv_out, cout_pad=cpad)
            out = ops.to_nchw(img, channels=self.out_channels, ld=cpad)
        self.last_launches = _lib.launch_count() - before
        return out

    def forward(self, img: torch.Tensor, epsilon: Optional[torch.Tensor] = None):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # the reference never trains its autoencoder on a live path (train_autoencoder.py / AutoencoderTrainer are dead
            # code); this class is inference-only and says so instead of returning detached tensors
            raise _lib.LdmError("ldm_b200.Autoencoder is inference-only (no backward kernels): call it under torch.no_grad() "
                                "or freeze its parameters with requires_grad_(False)")
        self.distribution = self.encode(img, epsilon)
        return self.decode(self.distribution.sample()), self.distribution.mu, self.distribution.log_var
