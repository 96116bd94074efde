"""Kernel-level Python wrappers over the C ABI (NHWC tensors in the kernels' compute dtype).

These mirror the PyTorch calls the reference makes inside src/UNet.py (F.group_norm + SiLU, F.conv2d,
F.conv_transpose2d, the attention einsums, F.max_pool2d) one kernel at a time.  They exist for unit
parity tests, profiling and benchmarks; the UNet handle launches the same kernels natively.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

_TORCH_DT = {"fp32": torch.float32, "bf16": torch.bfloat16}


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.LdmError("CUDA tensors only (no CPU fallback)")


def to_nhwc(x_nchw: torch.Tensor, dtype: str = "fp32") -> torch.Tensor:
    """fp32 NCHW -> NHWC in the compute dtype."""
    _cuda(x_nchw)
    B, C, H, W = x_nchw.shape
    x = x_nchw.to(torch.float32).contiguous()
    y = torch.empty(B, H, W, C, dtype=_TORCH_DT[dtype], device=x.device)
    _lib.check(_lib.load().ldm_nchw_to_nhwc(x.data_ptr(), y.data_ptr(), B, C, H * W, _dt(y), _lib.stream_ptr()))
    return y


def to_nchw(x_nhwc: torch.Tensor, channels: Optional[int] = None, ld: Optional[int] = None) -> torch.Tensor:
    """NHWC (pixel stride ld) -> fp32 NCHW."""
    _cuda(x_nhwc)
    B, H, W, Cfull = x_nhwc.shape
    C = channels or Cfull
    y = torch.empty(B, C, H, W, dtype=torch.float32, device=x_nhwc.device)
    _lib.check(_lib.load().ldm_nhwc_to_nchw(x_nhwc.data_ptr(), ld or Cfull, y.data_ptr(), B, C, H * W, _dt(x_nhwc),
                                            _lib.stream_ptr()))
    return y


def group_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float = 1e-5,
               silu: bool = False, res: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
               channels: Optional[int] = None) -> torch.Tensor:
    """y = [silu](GroupNorm(x)) [+ res] on NHWC x (src/UNet.py:52-58,106,147).  `x`, `res`, `out` may be
    channel-slices of wider NHWC buffers (their last-dim stride-1 views): pixel strides are taken from .stride(2)."""
    _cuda(x, gamma, beta, res)
    B, H, W, _ = x.shape
    C = channels or x.shape[3]
    if out is None:
        out = torch.empty(B, H, W, C, dtype=x.dtype, device=x.device)
    lib = _lib.load()
    ws = torch.empty(lib.ldm_group_norm_workspace_bytes(B, groups), dtype=torch.uint8, device=x.device)
    _lib.check(lib.ldm_group_norm(x.data_ptr(), x.stride(2), out.data_ptr(), out.stride(2), _lib.ptr(res),
                                  res.stride(2) if res is not None else 0, gamma.data_ptr(), beta.data_ptr(), B, H * W,
                                  C, groups, eps, int(silu), _dt(x), ws.data_ptr(), _lib.stream_ptr()))
    return out


def pack_conv_weight(w_oihw: torch.Tensor, dtype: str, w2_oi11: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(w_oihw, w2_oi11)
    cout, cin, k, _ = w_oihw.shape
    cin2 = w2_oi11.shape[1] if w2_oi11 is not None else 0
    out = torch.empty(cout, k * k * cin + cin2, dtype=_TORCH_DT[dtype], device=w_oihw.device)
    w = w_oihw.to(torch.float32).contiguous()
    w2 = w2_oi11.to(torch.float32).contiguous() if w2_oi11 is not None else None
    _lib.check(_lib.load().ldm_pack_conv_weight(w.data_ptr(), cout, cin, k, _lib.ptr(w2), cin2, out.data_ptr(),
                                                _dt(out), _lib.stream_ptr()))
    return out


def conv2d(x: torch.Tensor, w_packed: torch.Tensor, ksize: int, bias: Optional[torch.Tensor] = None,
           x2: Optional[torch.Tensor] = None, rowvec: Optional[torch.Tensor] = None,
           res: Optional[torch.Tensor] = None, impl: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Implicit-GEMM conv (3x3 pad 1 / 1x1) on NHWC x; optional K-concatenated 1x1 source x2, per-sample row
    vector [B, Cout] and residual (src/UNet.py:54,82,88-99,119-120)."""
    _cuda(x, w_packed, bias, x2, rowvec, res)
    B, H, W, cin = x.shape
    cout = w_packed.shape[0]
    cin2 = x2.shape[3] if x2 is not None else 0
    assert w_packed.shape[1] == ksize * ksize * cin + cin2
    if out is None:
        out = torch.empty(B, H, W, cout, dtype=x.dtype, device=x.device)
    _lib.check(_lib.load().ldm_conv2d(
        x.data_ptr(), x.stride(2), cin, _lib.ptr(x2), x2.stride(2) if x2 is not None else 0, cin2,
        w_packed.data_ptr(), _lib.ptr(bias), _lib.ptr(rowvec), rowvec.stride(0) if rowvec is not None else 0,
        _lib.ptr(res), res.stride(2) if res is not None else 0, out.data_ptr(), out.stride(2), cout, B, H, W, ksize,
        _dt(x), impl, _lib.stream_ptr()))
    return out


def conv2d_gn(x: torch.Tensor, w_packed: torch.Tensor, ksize: int, bias: Optional[torch.Tensor], *, mode: int, groups: int,
              gamma: Optional[torch.Tensor] = None, beta: Optional[torch.Tensor] = None, silu: bool = False,
              eps: float = 1e-5, x2: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None,
              gn_rowvec: Optional[torch.Tensor] = None, gn_res: Optional[torch.Tensor] = None, nvar: int = 1,
              out: Optional[torch.Tensor] = None, scratch: Optional[torch.Tensor] = None, tag: int = 1):
    """Convolution with the GroupNorm that follows it fused into the tcgen05 epilogue (bf16 NHWC; csrc/conv_epilogue.cuh).
    mode 1 -> (y, stats) with stats float32 [B, groups, nslots, 2] partial sums {S, Q} of the stored y;
    mode 2 -> y = [silu](GroupNorm(conv(x) + gn_rowvec)) [+ gn_res]; with nvar = 2 the output holds 2B images
    (image n normalised with row vectors n and n + B)."""
    _cuda(x, w_packed, bias, x2, res, gamma, beta, gn_rowvec, gn_res)
    B, H, W, cin = x.shape
    cout = w_packed.shape[0]
    cin2 = x2.shape[3] if x2 is not None else 0
    assert w_packed.shape[1] == ksize * ksize * cin + cin2 and x.dtype == torch.bfloat16
    lib = _lib.load()
    if out is None:
        out = torch.empty(B * nvar, H, W, cout, dtype=x.dtype, device=x.device)
    if scratch is None:
        scratch = torch.zeros(lib.ldm_conv2d_gn_scratch_bytes(B * nvar), dtype=torch.uint8, device=x.device)
    import ctypes
    nslots = ctypes.c_int(0)
    with torch.cuda.device(x.device):
        _lib.check(lib.ldm_conv2d_gn(
            x.data_ptr(), x.stride(2), cin, _lib.ptr(x2), x2.stride(2) if x2 is not None else 0, cin2, w_packed.data_ptr(),
            _lib.ptr(bias), _lib.ptr(res), res.stride(2) if res is not None else 0, out.data_ptr(), out.stride(2), cout, B, H, W,
            ksize, mode, groups, eps, int(silu), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(gn_rowvec),
            gn_rowvec.stride(0) if gn_rowvec is not None else 0, _lib.ptr(gn_res), gn_res.stride(2) if gn_res is not None else 0,
            nvar, B if nvar > 1 else 0, scratch.data_ptr(), scratch.numel(), tag, ctypes.byref(nslots), _lib.stream_ptr()))
    if mode == 1:
        stats = scratch[: B * groups * nslots.value * 8].view(torch.float32).view(B, groups, nslots.value, 2)
        return out, stats
    return out


def conv_transpose2x2(x: torch.Tensor, w_iohw: torch.Tensor, bias: torch.Tensor, impl: int = 0,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ConvTranspose2d(k=2, s=2) on NHWC x (src/UNet.py:231-233)."""
    _cuda(x, w_iohw, bias)
    B, H, W, cin = x.shape
    cout = w_iohw.shape[1]
    lib = _lib.load()
    wp = torch.empty(4 * cout, cin, dtype=x.dtype, device=x.device)
    w = w_iohw.to(torch.float32).contiguous()
    _lib.check(lib.ldm_pack_conv_transpose_weight(w.data_ptr(), cin, cout, wp.data_ptr(), _dt(x), _lib.stream_ptr()))
    if out is None:
        out = torch.empty(B, 2 * H, 2 * W, cout, dtype=x.dtype, device=x.device)
    _lib.check(lib.ldm_conv_transpose2x2(x.data_ptr(), x.stride(2), cin, wp.data_ptr(), bias.data_ptr(),
                                         out.data_ptr(), out.stride(2), cout, B, H, W, _dt(x), impl,
                                         _lib.stream_ptr()))
    return out


def linear_attention(qkv: torch.Tensor) -> torch.Tensor:
    """qkv [B,H,W,384] -> [B,H,W,128] (src/UNet.py:149-163)."""
    _cuda(qkv)
    B, H, W, _ = qkv.shape
    out = torch.empty(B, H, W, 128, dtype=qkv.dtype, device=qkv.device)
    _lib.check(_lib.load().ldm_linear_attention(qkv.data_ptr(), out.data_ptr(), B, H * W, _dt(qkv), _lib.stream_ptr()))
    return out


def linear_attention_prenorm(x: torch.Tensor, w_qkv: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
                             impl: int = 0) -> torch.Tensor:
    """PreNorm GroupNorm(1, C) + to_qkv + LinearAttention core on the raw bf16 NHWC input x [B,H,W,64] -> [B,H,W,128]
    (src/UNet.py:106-110,145,149-163).  impl 0: tcgen05 kernel, 1: mma.sync kernel."""
    _cuda(x, w_qkv, gamma, beta)
    B, H, W, cin = x.shape
    assert x.dtype == torch.bfloat16 and w_qkv.shape[0] == 384
    lib = _lib.load()
    w = w_qkv.reshape(384, cin).to(torch.float32).contiguous()
    out = torch.empty(B, H, W, 128, dtype=x.dtype, device=x.device)
    nbytes = lib.ldm_linear_attention_prenorm_scratch_bytes(B)
    scratch = torch.empty(nbytes + 256, dtype=torch.uint8, device=x.device)
    off = (-scratch.data_ptr()) % 256
    with torch.cuda.device(x.device):
        _lib.check(lib.ldm_linear_attention_prenorm(x.data_ptr(), x.stride(2), cin, w.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                                    eps, out.data_ptr(), B, H * W, impl, scratch.data_ptr() + off, nbytes,
                                                    _lib.stream_ptr()))
    return out


def linear_attention_prenorm_to_out(x: torch.Tensor, w_qkv: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                                    w_out: torch.Tensor, b_out: torch.Tensor, eps: float = 1e-5):
    """PreNorm + to_qkv + LinearAttention + to_out.0 (1x1 conv incl. bias) in the tcgen05 kernel -> (y [B,H,W,64] bf16,
    stats fp32 [B, H*W/16, 2] partial sums {S, Q} of y for the GroupNorm(1, C) that follows)  (src/UNet.py:106-110,145-163)."""
    _cuda(x, w_qkv, gamma, beta, w_out, b_out)
    B, H, W, cin = x.shape
    lib = _lib.load()
    w = w_qkv.reshape(384, cin).to(torch.float32).contiguous()
    wo = w_out.reshape(64, 128).to(torch.float32).contiguous()
    bo = b_out.to(torch.float32).contiguous()
    y = torch.empty(B, H, W, 64, dtype=x.dtype, device=x.device)
    stats = torch.zeros(B, H * W // 16, 2, dtype=torch.float32, device=x.device)
    nbytes = lib.ldm_linear_attention_prenorm_scratch_bytes(B)
    scratch = torch.empty(nbytes + 256, dtype=torch.uint8, device=x.device)
    off = (-scratch.data_ptr()) % 256
    with torch.cuda.device(x.device):
        _lib.check(lib.ldm_linear_attention_prenorm_to_out(x.data_ptr(), x.stride(2), cin, w.data_ptr(), gamma.data_ptr(),
                                                           beta.data_ptr(), eps, wo.data_ptr(), bo.data_ptr(), y.data_ptr(),
                                                           y.stride(2), stats.data_ptr(), B, H * W, scratch.data_ptr() + off, nbytes,
                                                           _lib.stream_ptr()))
    return y, stats


def attention(qkv: torch.Tensor) -> torch.Tensor:
    """qkv [B,H,W,384] -> [B,H,W,128] (src/UNet.py:122-135)."""
    _cuda(qkv)
    B, H, W, _ = qkv.shape
    out = torch.empty(B, H, W, 128, dtype=qkv.dtype, device=qkv.device)
    _lib.check(_lib.load().ldm_attention(qkv.data_ptr(), out.data_ptr(), B, H * W, _dt(qkv), _lib.stream_ptr()))
    return out


def max_pool2x2(x: torch.Tensor) -> torch.Tensor:
    _cuda(x)
    B, H, W, C = x.shape
    out = torch.empty(B, H // 2, W // 2, C, dtype=x.dtype, device=x.device)
    _lib.check(_lib.load().ldm_max_pool2x2(x.data_ptr(), x.stride(2), out.data_ptr(), C, B, H, W, C, _dt(x),
                                           _lib.stream_ptr()))
    return out


_U8_CONVENTIONS = {"save_image": 0, "reverse_transform": 1}


def images_to_uint8(x_nchw: torch.Tensor, convention: str = "save_image") -> torch.Tensor:
    """fp32 [B,C,H,W] -> uint8 [B,H,W,C] on the device.  ``save_image``: what torchvision.utils.save_image writes for the
    raw tensor (src/utils.py:121-130); ``reverse_transform``: the reference's PIL path (src/transforms.py:22-35)."""
    _cuda(x_nchw)
    if convention not in _U8_CONVENTIONS:
        raise ValueError(f"convention must be one of {sorted(_U8_CONVENTIONS)}")
    x = x_nchw.detach().to(torch.float32).contiguous()
    B, Cc, H, W = x.shape
    out = torch.empty(B, H, W, Cc, dtype=torch.uint8, device=x.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ldm_images_to_uint8(x.data_ptr(), out.data_ptr(), B, Cc, H * W, _U8_CONVENTIONS[convention],
                                                   _lib.stream_ptr()))
    return out


def _on_tensor_device(fn):
    """Run ``fn`` with its first CUDA tensor argument's device current: the kernels launch on ``_lib.stream_ptr()``, the
    CURRENT device's stream, so an op on a cuda:1 tensor must not run under cuda:0 (multi-GPU-in-one-process)."""
    import functools

    @functools.wraps(fn)
    def guarded(*args, **kwargs):
        for v in args + tuple(kwargs.values()):
            if isinstance(v, torch.Tensor) and v.is_cuda:
                if v.device.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(v.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return guarded


for _name, _fn in list(globals().items()):
    if isinstance(_fn, type(_on_tensor_device)) and _fn.__module__ == __name__ and not _name.startswith("_"):
        globals()[_name] = _on_tensor_device(_fn)
del _name, _fn
