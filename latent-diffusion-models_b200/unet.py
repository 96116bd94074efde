"""Drop-in ``UNet`` eps-model backed by the sm_100a kernels of libldm_b200.so.

Same constructor, call protocol, attributes and 200-key fp32 ``state_dict`` as the reference's
``src/UNet.py:293-389``; the module tree below only *holds parameters* (created in the reference's
construction order so that the same ``torch.manual_seed`` yields bit-identical default weights).
``forward`` hands device pointers to ``ldm_unet_forward`` -- no PyTorch compute, no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
from typing import Dict, List, Optional, Tuple, Union

import torch
from torch import nn

from . import _lib

_HIDDEN = 128  # 4 heads x 32 (src/UNet.py:114,140)
_UID = itertools.count(1)   # identity of a UNet object's native state (never reused, unlike id() / handle addresses)


# ----------------------------------------------------------------------------- parameter holders
def _seq_with_gaps(*mods) -> nn.Sequential:
    """nn.Sequential whose parameter-free slots are Identity, to reproduce key indices such as 'mlp_t.1'."""
    return nn.Sequential(*[m if m is not None else nn.Identity() for m in mods])


class _BlockParams(nn.Module):  # keys: norm.*, conv2d.*   (src/UNet.py:50-54)
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.norm = nn.GroupNorm(8, cin)
        self.conv2d = nn.Conv2d(cin, cout, 3, padding=1)


class _ResParams(nn.Module):  # keys: mlp_t.1.*, block1.*, block2.*, shortcut.*   (src/UNet.py:64-83)
    def __init__(self, cin: int, cout: int, temb: Optional[int]):
        super().__init__()
        if temb is not None:
            self.mlp_t = _seq_with_gaps(None, nn.Linear(temb, cout))
        self.block1 = _BlockParams(cin, cout)
        self.block2 = _BlockParams(cout, cout)
        if cin != cout:
            self.shortcut = nn.Conv2d(cin, cout, 1)


class _LinAttnParams(nn.Module):  # keys: to_qkv.weight, to_out.0.*, to_out.1.*   (src/UNet.py:140-147)
    def __init__(self, dim: int):
        super().__init__()
        self.to_qkv = nn.Conv2d(dim, 3 * _HIDDEN, 1, bias=False)
        self.to_out = nn.Sequential(nn.Conv2d(_HIDDEN, dim, 1), nn.GroupNorm(1, dim))


class _AttnParams(nn.Module):  # keys: to_qkv.weight, to_out.*   (src/UNet.py:114-120)
    def __init__(self, dim: int):
        super().__init__()
        self.to_qkv = nn.Conv2d(dim, 3 * _HIDDEN, 1, bias=False)
        self.to_out = nn.Conv2d(_HIDDEN, dim, 1)


class _PreNormParams(nn.Module):  # keys: fn.*, norm.*   (src/UNet.py:103-106)
    def __init__(self, dim: int, fn: nn.Module):
        super().__init__()
        self.fn = fn
        self.norm = nn.GroupNorm(1, dim)


class _ResidualParams(nn.Module):  # key prefix: fn.   (src/UNet.py:15-17)
    def __init__(self, fn: nn.Module):
        super().__init__()
        self.fn = fn


class _Holder(nn.Module):
    pass


def _attn_site(dim: int, linear: bool) -> nn.Module:
    return _ResidualParams(_PreNormParams(dim, _LinAttnParams(dim) if linear else _AttnParams(dim)))


class UNet(nn.Module):
    """eps_theta(x_t, t, y).  Extra keyword ``dtype`` ('bf16' | 'fp32') picks the kernel precision."""

    def __init__(self, in_channels: int, out_channels: int, channels: int = 64,
                 channel_multipliers: Union[Tuple[int, ...], List[int]] = (1, 2, 4, 8),
                 with_time_emb: bool = True, num_classes: Optional[int] = None, *,
                 dtype: Optional[str] = None, conv_impl: int = 0) -> None:
        super().__init__()
        self.in_channels, self.out_channels, self.channels = in_channels, out_channels, channels
        self.channel_multipliers = tuple(channel_multipliers)
        self.channels_list = [channels] + [channels * m for m in channel_multipliers]
        self.num_classes = num_classes
        self.with_time_emb = with_time_emb
        self.compute_dtype = (dtype or os.environ.get("LDM_B200_DTYPE", "bf16")).lower()
        if self.compute_dtype not in _lib.DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_lib.DTYPES)}")
        self.conv_impl = int(conv_impl)
        dims = self.channels_list
        temb = channels * 4 if with_time_emb else None
        # ---- parameters, in the reference's construction order (src/UNet.py:320-348)
        if with_time_emb:
            self.time_emb = _Holder()
            self.time_emb.time_mlp = _seq_with_gaps(None, nn.Linear(temb // 4, temb), None, nn.Linear(temb, temb))
        else:
            self.time_emb = None
        if num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, temb)
        self.initial_conv = nn.Conv2d(in_channels, channels, 3, padding=1)
        self.encoder = _Holder()
        self.encoder.downs = nn.ModuleList([
            nn.ModuleList([_ResParams(dims[i], dims[i + 1], temb), _attn_site(dims[i + 1], True)])
            for i in range(len(dims) - 1)])
        self.bottleneck = _Holder()
        self.bottleneck.res1 = _ResParams(dims[-1], dims[-1], temb)
        self.bottleneck.attn = _attn_site(dims[-1], False)
        self.bottleneck.res2 = _ResParams(dims[-1], dims[-1], temb)
        rd = list(reversed(dims))
        self.decoder = _Holder()
        self.decoder.ups = nn.ModuleList([
            nn.ModuleList([_ResParams(rd[i] + rd[i + 1], rd[i + 1], temb), _attn_site(rd[i + 1], True),
                           nn.ConvTranspose2d(rd[i], rd[i + 1], 2, 2)])
            for i in range(len(rd) - 1)])
        self.final_conv = nn.Sequential(_ResParams(channels, channels, None), nn.Conv2d(channels, out_channels, 1))
        self._reset_native()

    def _reset_native(self) -> None:
        """Native state (not part of state_dict): owned by exactly one Python object, rebuilt lazily."""
        self._handles: Dict[Tuple[int, int], int] = {}      # (image_size, device index) -> ldm_unet*
        self._loaded: Dict[int, tuple] = {}                 # handle -> parameter fingerprint
        self._ws: Dict[int, torch.Tensor] = {}               # device index -> workspace bytes
        self._uid = next(_UID)
        self.last_launches = 0

    def _destroy_native(self) -> None:
        try:
            for h in getattr(self, "_handles", {}).values():
                _lib.destroy_native("ldm_unet_destroy", h)     # parked while a CUDA-graph capture is in progress (cudaFree)
            if getattr(self, "_handles", None):
                self._handles = {}
        except Exception:
            pass

    def set_compute_dtype(self, dtype: str) -> None:
        """Switch the kernel precision ('bf16' | 'fp32'); native handles are rebuilt on the next call."""
        dtype = dtype.lower()
        if dtype not in _lib.DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_lib.DTYPES)}")
        if _lib.DTYPES[dtype] != _lib.DTYPES[self.compute_dtype]:
            self._destroy_native()
            self._reset_native()
        self.compute_dtype = dtype

    def __getstate__(self):
        # copy.deepcopy / pickle / torch.save(model): the copy must not share ldm_unet* handles (two owners would repack each
        # other's weights behind stale fingerprints and double-free in __del__); it builds its own on first use
        state = self.__dict__.copy()
        for k in ("_handles", "_loaded", "_ws", "_uid"):
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._reset_native()

    # ------------------------------------------------------------------ native plumbing
    def _handle(self, image_size: int, device: torch.device) -> int:
        key = (image_size, device.index if device.index is not None else torch.cuda.current_device())
        h = self._handles.get(key)
        if h is None:
            _lib.flush_graveyard()              # handles whose owners died during a stream capture
            lib = _lib.load()
            d = _lib.UNetDesc()
            d.in_channels, d.out_channels, d.channels = self.in_channels, self.out_channels, self.channels
            d.n_levels = len(self.channel_multipliers)
            for i, m in enumerate(self.channel_multipliers):
                d.channel_multipliers[i] = int(m)
            d.with_time_emb = int(self.with_time_emb)
            d.num_classes = int(self.num_classes or 0)
            d.image_size = int(image_size)
            d.dtype = _lib.DTYPES[self.compute_dtype]
            d.conv_impl = self.conv_impl
            out = C.c_void_p()
            with torch.cuda.device(device):
                _lib.check(lib.ldm_unet_create(C.byref(d), C.byref(out)))
            h = out.value
            names = [lib.ldm_unet_param_name(h, i).decode() for i in range(lib.ldm_unet_num_params(h))]
            mine = [k for k, _ in self.named_parameters()]
            if names != mine:
                raise _lib.LdmError("state_dict key order of the Python module and the native handle differ")
            self._handles[key] = h
        return h

    def _sync_params(self, h: int, device: torch.device) -> None:
        """Re-pack weights into the kernels' layouts when any parameter changed (training / load_state_dict)."""
        params = list(self.parameters())
        fp = tuple((p.data_ptr(), p._version) for p in params)
        if self._loaded.get(h) == fp:
            return
        for p in params:
            if p.device != device or p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.LdmError("UNet parameters must be contiguous fp32 tensors on the input's CUDA device "
                                    "(call model.to(device)); there is no CPU path")
        arr = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
        _lib.check(_lib.load().ldm_unet_load_params(h, arr, len(params), _lib.stream_ptr()))
        self._loaded[h] = fp

    def _workspace(self, nbytes: int, device: torch.device) -> torch.Tensor:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        ws = self._ws.get(idx)
        if ws is None or ws.numel() < nbytes + 1024:
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            self._ws[idx] = ws
        off = (-ws.data_ptr()) % 1024
        return ws[off:off + nbytes]

    def native(self, image_size: int, device: torch.device) -> int:
        """Native handle with up-to-date packed weights (used by Diffusion.sample's graph sampler)."""
        h = self._handle(image_size, device)
        self._sync_params(h, device)
        return h

    def set_tap(self, name: Optional[str], out: Optional[torch.Tensor], image_size: int) -> None:
        h = self._handle(image_size, out.device if out is not None else next(self.parameters()).device)
        _lib.check(_lib.load().ldm_unet_set_tap(h, name.encode() if name else None, _lib.ptr(out),
                                                out.numel() if out is not None else 0))

    # ------------------------------------------------------------------ nn.Module protocol
    def forward(self, x_noisy: torch.Tensor, t: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not x_noisy.is_cuda:
            raise _lib.LdmError("ldm_b200.UNet runs on CUDA tensors only (no CPU fallback)")
        if torch.is_grad_enabled() and x_noisy.requires_grad:
            raise _lib.LdmError("ldm_b200.UNet has no gradient with respect to its input x_noisy (the training step never needs "
                                "it: src/DiffusionModelTrainer.py:41-63); detach the input")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .train import unet_autograd_forward
            return unet_autograd_forward(self, x_noisy, t, y)
        return self._forward_nograd(x_noisy, t, y)

    def _forward_nograd(self, x_noisy, t, y=None, y_rows: Optional[int] = None):
        dev = x_noisy.device
        B, Cin, H, W = x_noisy.shape
        if Cin != self.in_channels or H != W:
            raise ValueError(f"expected [B,{self.in_channels},S,S] input, got {tuple(x_noisy.shape)}")
        x = x_noisy.detach().to(torch.float32).contiguous()
        t = t.detach().to(device=dev, dtype=torch.int64).contiguous()
        if t.numel() != B:
            raise ValueError("t must have one entry per sample")
        y_len = 0
        if y is not None:
            if self.num_classes is None:
                raise ValueError("labels given but num_classes is None")
            y = y.detach().to(device=dev, dtype=torch.int64).contiguous()
            y_len = y.numel()
        with torch.cuda.device(dev):
            lib = _lib.load()
            h = self.native(H, dev)
            nbytes = lib.ldm_unet_workspace_bytes(h, B)
            ws = self._workspace(nbytes, dev)
            out = torch.empty(B, self.out_channels, H, W, dtype=torch.float32, device=dev)
            before = _lib.launch_count()
            _lib.check(lib.ldm_unet_forward(h, x.data_ptr(), t.data_ptr(), None, _lib.ptr(y), y_len,
                                            y_rows if y_rows is not None else B, B, out.data_ptr(),
                                            ws.data_ptr(), nbytes, _lib.stream_ptr()))
            self.last_launches = _lib.launch_count() - before
        return out

    @torch.no_grad()
    def profile(self, x_noisy: torch.Tensor, t: torch.Tensor, y: Optional[torch.Tensor] = None,
                y_rows: Optional[int] = None) -> Dict[str, Dict[str, float]]:
        """One forward with a CUDA-event pair around every launch; per-family {ms, flops, bytes, launches}."""
        dev = x_noisy.device
        B, _, H, _ = x_noisy.shape
        x = x_noisy.detach().to(torch.float32).contiguous()
        t = t.detach().to(device=dev, dtype=torch.int64).contiguous()
        y = y.detach().to(device=dev, dtype=torch.int64).contiguous() if y is not None else None
        with torch.cuda.device(dev):
            lib = _lib.load()
            h = self.native(H, dev)
            nbytes = lib.ldm_unet_workspace_bytes(h, B)
            ws = self._workspace(nbytes, dev)
            out = torch.empty(B, self.out_channels, H, H, dtype=torch.float32, device=dev)
            prof = _lib.Profile()
            _lib.check(lib.ldm_unet_profile(h, x.data_ptr(), t.data_ptr(), _lib.ptr(y), y.numel() if y is not None else 0,
                                            y_rows if y_rows is not None else B, B, out.data_ptr(), ws.data_ptr(), nbytes,
                                            _lib.stream_ptr(), C.byref(prof)))
        return {name: {"ms": prof.family[i].ms, "flops": prof.family[i].flops, "bytes": prof.family[i].bytes,
                       "launches": int(prof.family[i].launches)}
                for i, name in enumerate(_lib.FAMILIES) if prof.family[i].launches}

    def __del__(self):
        self._destroy_native()
