"""Training-step path: ``UNet.forward`` with gradients (src/DiffusionModelTrainer.py:36-67).

The reference gets its backward pass from autograd over ATen/cuDNN.  Here every forward kernel of the C ABI has a
hand-written backward kernel (csrc/backward.cu) and the two are tied together by ``torch.autograd.Function`` -- autograd
is used for what it is (the graph, saved tensors, gradient fan-in at the residual / concat joins); every tensor
operation on the path is a kernel of libldm_b200.so.  Activations are NHWC in the model's compute dtype, parameters stay
the fp32 ``nn.Parameter``s of the 200-key ``state_dict`` and receive fp32 ``.grad`` (so Adam, GradScaler-free bf16
training, ``wandb.watch`` and checkpointing behave as in the reference).  The four dead ``bottleneck.res{1,2}.mlp_t``
tensors get no gradient, exactly as in the reference.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch.autograd import Function

from . import _lib, ops

_HIDDEN = 128
_EPS = 1e-5


def _st():
    return _lib.stream_ptr()


def _lb():
    return _lib.load()


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


class _GradArena:
    """One zero-filled fp32 buffer per training forward from which the backward kernels take the accumulation targets that have
    no slot in an optimizer's gradient bucket (see ``_grad_target``): one memset instead of ~120 small fills per step.  Slices
    are 256-byte aligned (the weight-gradient kernel uses 16-byte vector reductions).  Falls back to torch.zeros when no arena
    is open or it is used up (several backward passes through one forward)."""
    buf: Optional[torch.Tensor] = None
    off: int = 0

    @classmethod
    def open(cls, model, device, bucketed: bool) -> None:
        key = "_grad_arena_elems_bucketed" if bucketed else "_grad_arena_elems"
        n = getattr(model, key, None)
        if n is None:
            if bucketed:   # only what never lands in a slot: the concatenated time-projection weights and biases, the
                # [4*Cout][Cin] staging form of the ConvTranspose weight gradients, the final conv's viewed weight + slack
                n = sum((p.numel() + 63) // 64 * 64 + 64 for name, p in model.named_parameters() if ".mlp_t." in name)
                n += sum((m.weight.numel() + 63) // 64 * 64 + 64 for m in model.modules() if isinstance(m, torch.nn.ConvTranspose2d))
                n += 262144
            else:
                n = sum((p.numel() + 63) // 64 * 64 + 64 for p in model.parameters()) + 4096
            setattr(model, key, n)
        cls.buf = torch.zeros(n, dtype=torch.float32, device=device)
        cls.off = 0

    @classmethod
    def zeros(cls, n: int, device) -> torch.Tensor:
        n_al = (n + 63) // 64 * 64
        b = cls.buf
        if b is not None and b.device == device and cls.off + n_al <= b.numel():
            v = b[cls.off:cls.off + n]
            cls.off += n_al
            return v
        return torch.zeros(n, dtype=torch.float32, device=device)


def _zeros(n: int, device) -> torch.Tensor:
    return _GradArena.zeros(int(n), device)


def _grad_target(p: torch.Tensor) -> torch.Tensor:
    """Zero-filled fp32 accumulation target for the gradient of parameter ``p`` (same shape).  When ``p`` belongs to a
    ``trainer.FlatAdam`` this is its slice of the optimizer's flat gradient bucket: the backward kernel writes where the
    all-reduce and the Adam kernel read, and autograd adopts the slice as ``p.grad`` (no per-step packing).  Otherwise a
    slice of the step's arena."""
    bucket = getattr(p, "_ldm_grad_bucket", None)
    if bucket is not None:
        slot = bucket.take_slot(p)
        if slot is not None:
            return slot
    return _zeros(p.numel(), p.device).view_as(p)


class _Conv(Function):
    """F.conv2d (3x3 pad 1 / 1x1) on NHWC; dgrad = the same implicit-GEMM kernel with the flipped, transposed filter."""

    @staticmethod
    def forward(ctx, x, w, b, impl):
        k = w.shape[2]
        dt = "bf16" if x.dtype == torch.bfloat16 else "fp32"
        cout, cin = w.shape[0], w.shape[1]
        # both packed forms of the filter in one launch: [cout][tap*cin] for this conv, [cin][tap'*cout] (flipped) for dgrad
        wp = torch.empty(cout, k * k * cin, dtype=x.dtype, device=x.device)
        wd = torch.empty(cin, k * k * cout, dtype=x.dtype, device=x.device)
        _lib.check(_lb().ldm_pack_conv_weight_pair(w.data_ptr(), cout, cin, k, wp.data_ptr(), wd.data_ptr(), ops._dt(x), _st()))
        y = ops.conv2d(x, wp, k, bias=b.detach() if b is not None else None, impl=impl)
        ctx.save_for_backward(x, w, wd)
        ctx.has_bias, ctx.impl, ctx.dt = b is not None, impl, dt
        ctx.bias = b                                # the Parameter itself (not saved: only its gradient slot is looked up)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, wd = ctx.saved_tensors
        dy = _c(dy)
        cout, cin, k, _ = w.shape
        B, H, W, _ = x.shape
        lib = _lb()
        dx = ops.conv2d(dy, wd, k, impl=ctx.impl)
        dw = _grad_target(w)                        # zero-filled: the optimizer's bucket slice, or the step's arena
        db = _grad_target(ctx.bias) if ctx.has_bias else None
        nscr = lib.ldm_conv2d_wgrad_scratch_bytes(cin, cout, B, H, W, k, ops._dt(x)) if ctx.impl == 0 else 0
        if nscr > 0:   # tcgen05: contraction over pixels on channel-major copies of x and dy
            scr = torch.empty(nscr, dtype=torch.uint8, device=x.device)
            _lib.check(lib.ldm_conv2d_wgrad_tc(x.data_ptr(), x.stride(2), cin, dy.data_ptr(), dy.stride(2), cout, dw.data_ptr(),
                                               _lib.ptr(db), B, H, W, k, scr.data_ptr(), _st()))
        else:
            _lib.check(lib.ldm_conv2d_wgrad(x.data_ptr(), x.stride(2), cin, dy.data_ptr(), dy.stride(2), cout, dw.data_ptr(),
                                            _lib.ptr(db), B, H, W, k, ops._dt(x), _st()))
        return dx, dw, db, None


class _GroupNorm(Function):
    """[SiLU](GroupNorm(x + rowvec)); rowvec is the ResNetBlock's time-embedding projection (src/UNet.py:88-96)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, silu, rowvec):
        B, H, W, Cc = x.shape
        y = torch.empty_like(x)
        lib = _lb()
        ws = torch.empty(lib.ldm_group_norm_workspace_bytes(B, groups), dtype=torch.uint8, device=x.device)
        rv = rowvec.detach() if rowvec is not None else None
        _lib.check(lib.ldm_group_norm_rowvec(x.data_ptr(), x.stride(2), y.data_ptr(), y.stride(2), None, 0, gamma.data_ptr(),
                                             beta.data_ptr(), _lib.ptr(rv), rv.stride(0) if rv is not None else 0, B, H * W,
                                             Cc, groups, _EPS, int(silu), ops._dt(x), ws.data_ptr(), _st()))
        ctx.save_for_backward(x, gamma, beta, rowvec if rowvec is not None else torch.empty(0), ws)   # ws: the statistics
        ctx.groups, ctx.silu, ctx.has_rv = groups, silu, rowvec is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, rowvec, fwd_ws = ctx.saved_tensors
        dy = _c(dy)
        B, H, W, Cc = x.shape
        dx = torch.empty_like(x)
        dg, db = _grad_target(gamma), _grad_target(beta)   # zero-filled
        rv = rowvec if ctx.has_rv else None
        drv = torch.empty(B, Cc, dtype=torch.float32, device=x.device) if ctx.has_rv else None
        lib = _lb()
        ws = torch.empty(lib.ldm_group_norm_backward_workspace_bytes(B, H * W, Cc, ctx.groups), dtype=torch.uint8, device=x.device)
        _lib.check(lib.ldm_group_norm_backward(x.data_ptr(), x.stride(2), dy.data_ptr(), dy.stride(2), gamma.data_ptr(),
                                               beta.data_ptr(), _lib.ptr(rv), rv.stride(0) if rv is not None else 0,
                                               dx.data_ptr(), dx.stride(2), dg.data_ptr(), db.data_ptr(), _lib.ptr(drv),
                                               Cc, B, H * W, Cc, ctx.groups, _EPS, int(ctx.silu), ops._dt(x),
                                               fwd_ws.data_ptr(), ws.data_ptr(), _st()))
        return dx, dg, db, None, None, drv


class _MaxPool(Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.max_pool2x2(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = _c(dy)
        B, H, W, Cc = x.shape
        dx = torch.empty_like(x)
        _lib.check(_lb().ldm_max_pool2x2_backward(x.data_ptr(), x.stride(2), dy.data_ptr(), dy.stride(2), dx.data_ptr(),
                                                  dx.stride(2), B, H, W, Cc, ops._dt(x), _st()))
        return dx


class _ConvT(Function):
    """ConvTranspose2d(k2, s2): forward = [M, 4*Cout] GEMM + pixel shuffle; backward = gather + 1x1 dgrad / wgrad."""

    @staticmethod
    def forward(ctx, x, w, b, impl):
        ctx.save_for_backward(x, w)
        ctx.impl, ctx.bias = impl, b
        return ops.conv_transpose2x2(x, w.detach(), b.detach(), impl=impl)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _c(dy)
        B, H, W, cin = x.shape
        cout = w.shape[1]
        lib = _lb()
        dyq = torch.empty(B, H, W, 4 * cout, dtype=x.dtype, device=x.device)
        _lib.check(lib.ldm_pixel_unshuffle2x2(dy.data_ptr(), dy.stride(2), dyq.data_ptr(), B, H, W, cout, ops._dt(x), _st()))
        # dx[m][ci] = sum_{q,co} dyq[m][q*Cout+co] w[ci][co][q]: a 1x1 conv whose packed filter is [ci][(q,co)]
        wd = w.detach().permute(0, 2, 3, 1).reshape(cin, 4 * cout).to(x.dtype).contiguous()   # layout only
        dx = ops.conv2d(dyq, wd, 1, impl=ctx.impl)
        # dW'[(q,co)][ci] = sum_m dyq[m][(q,co)] x[m][ci]  -> back to the IOHW parameter layout
        dwq = _zeros(4 * cout * cin, x.device).view(4 * cout, cin)
        nscr = lib.ldm_conv2d_wgrad_scratch_bytes(cin, 4 * cout, B, H, W, 1, ops._dt(x)) if ctx.impl == 0 else 0
        if nscr > 0:
            scr = torch.empty(nscr, dtype=torch.uint8, device=x.device)
            _lib.check(lib.ldm_conv2d_wgrad_tc(x.data_ptr(), x.stride(2), cin, dyq.data_ptr(), 4 * cout, 4 * cout, dwq.data_ptr(),
                                               None, B, H, W, 1, scr.data_ptr(), _st()))
        else:
            _lib.check(lib.ldm_conv2d_wgrad(x.data_ptr(), x.stride(2), cin, dyq.data_ptr(), 4 * cout, 4 * cout, dwq.data_ptr(),
                                            None, B, H, W, 1, ops._dt(x), _st()))
        dw = _grad_target(w)
        dw.copy_(dwq.view(2, 2, cout, cin).permute(3, 2, 0, 1))                                # layout only
        db = _grad_target(ctx.bias)
        _lib.check(lib.ldm_column_sum(dy.data_ptr(), dy.stride(2), db.data_ptr(), B * 4 * H * W, cout, ops._dt(x), _st()))
        return dx, dw, db, None


class _LinAttn(Function):
    @staticmethod
    def forward(ctx, qkv):
        ctx.save_for_backward(qkv)
        return ops.linear_attention(qkv)

    @staticmethod
    def backward(ctx, dout):
        (qkv,) = ctx.saved_tensors
        dout = _c(dout)
        B, H, W, _ = qkv.shape
        dqkv = torch.empty_like(qkv)
        lib = _lb()
        ws = torch.empty(lib.ldm_linear_attention_backward_workspace_bytes(B), dtype=torch.uint8, device=qkv.device)
        _lib.check(lib.ldm_linear_attention_backward(qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), B, H * W,
                                                     ops._dt(qkv), ws.data_ptr(), _st()))
        return dqkv


class _Attn(Function):
    @staticmethod
    def forward(ctx, qkv):
        ctx.save_for_backward(qkv)
        return ops.attention(qkv)

    @staticmethod
    def backward(ctx, dout):
        (qkv,) = ctx.saved_tensors
        dout = _c(dout)
        B, H, W, _ = qkv.shape
        dqkv = torch.empty_like(qkv)
        _lib.check(_lb().ldm_attention_backward(qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), B, H * W, ops._dt(qkv), _st()))
        return dqkv


class _InitialConv(Function):
    @staticmethod
    def forward(ctx, x, w, b, dtype):
        B, Cin, H, W = x.shape
        cout = w.shape[0]
        y = torch.empty(B, H, W, cout, dtype=ops._TORCH_DT[dtype], device=x.device)
        scratch = torch.empty(9 * Cin * cout, dtype=torch.float32, device=x.device)
        _lib.check(_lb().ldm_initial_conv(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), B, Cin, cout, H, W,
                                          ops._dt(y), scratch.data_ptr(), _st()))
        ctx.save_for_backward(x, w)
        ctx.bias = b
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _c(dy)
        B, Cin, H, W = x.shape
        cout = w.shape[0]
        dw = _grad_target(w)
        db = _grad_target(ctx.bias)
        lib = _lb()
        scr = torch.empty(lib.ldm_initial_conv_wgrad_scratch_bytes(B, Cin, cout, H, W), dtype=torch.uint8, device=x.device)
        _lib.check(lib.ldm_initial_conv_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), B, Cin, cout, H, W,
                                              ops._dt(dy), scr.data_ptr(), _st()))
        return None, dw, db, None     # the noised image x_t needs no gradient (src/DDPM.py:133-149)


class _FinalConv(Function):
    @staticmethod
    def forward(ctx, x, w, b):
        B, H, W, cin = x.shape
        cout = w.shape[0]                      # w: the 1x1 Conv2d weight [out, ch, 1, 1] (contiguous: read as [out][ch])
        y = torch.empty(B, cout, H, W, dtype=torch.float32, device=x.device)
        _lib.check(_lb().ldm_final_conv(x.data_ptr(), x.stride(2), w.data_ptr(), b.data_ptr(), y.data_ptr(), B, cin, cout,
                                        H * W, ops._dt(x), _st()))
        ctx.save_for_backward(x, w)
        ctx.bias = b
        return y

    @staticmethod
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        dout = _c(dout.to(torch.float32))
        B, H, W, cin = x.shape
        cout = w.shape[0]
        dx = torch.empty_like(x)
        dw = _grad_target(w)
        db = _grad_target(ctx.bias)
        _lib.check(_lb().ldm_final_conv_backward(dout.data_ptr(), x.data_ptr(), x.stride(2), w.data_ptr(), dx.data_ptr(),
                                                 dw.data_ptr(), db.data_ptr(), B, cin, cout, H * W, ops._dt(x), _st()))
        return dx, dw, db


def _time_ws(B, D, total, dev):
    return torch.empty(_lb().ldm_time_workspace_bytes(B, D, total), dtype=torch.uint8, device=dev)


class _TimeEmbed(Function):
    @staticmethod
    def forward(ctx, t, y, w1, b1, w3, b3, label):
        B, D = t.numel(), w3.shape[0]
        temb = torch.empty(B, D, dtype=torch.float32, device=w1.device)
        ws = _time_ws(B, D, 0, w1.device)
        _lib.check(_lb().ldm_time_embed(t.data_ptr(), _lib.ptr(y), y.numel() if y is not None else 0, w1.data_ptr(),
                                        b1.data_ptr(), w3.data_ptr(), b3.data_ptr(), _lib.ptr(label), temb.data_ptr(), B, D,
                                        ws.data_ptr(), _st()))
        ctx.save_for_backward(t, y if y is not None else torch.empty(0), w1, b1, w3, label if label is not None else torch.empty(0))
        ctx.has_y = y is not None
        ctx.b3 = b3
        return temb

    @staticmethod
    def backward(ctx, dtemb):
        t, y, w1, b1, w3, label = ctx.saved_tensors
        dtemb = _c(dtemb)
        B, D = dtemb.shape
        dw1, db1, dw3, db3 = (_grad_target(v) for v in (w1, b1, w3, ctx.b3))
        dlabel = _grad_target(label) if ctx.has_y else None
        ws = _time_ws(B, D, 0, w1.device)
        yy = y if ctx.has_y else None
        _lib.check(_lb().ldm_time_embed_backward(t.data_ptr(), _lib.ptr(yy), yy.numel() if yy is not None else 0,
                                                 w1.data_ptr(), b1.data_ptr(), w3.data_ptr(), dtemb.data_ptr(),
                                                 dw1.data_ptr(), db1.data_ptr(), dw3.data_ptr(), db3.data_ptr(),
                                                 _lib.ptr(dlabel), B, D, ws.data_ptr(), _st()))
        return None, None, dw1, db1, dw3, db3, dlabel


class _TimeProj(Function):
    """The ResNetBlocks' time-embedding projections ``mlp_t`` = Linear(SiLU(temb)) (src/UNet.py:70-73,88-93) of several blocks
    as ONE GEMM over their concatenated weights.  The blocks' (weight, bias) Parameters are the inputs themselves (not a
    torch.cat of them), so that backward can drop every block's rows of dW / db into that parameter's gradient slot."""

    @staticmethod
    def forward(ctx, temb, *wb):
        ws_, bs_ = wb[0::2], wb[1::2]
        w = torch.cat([v.detach() for v in ws_], 0)
        b = torch.cat([v.detach() for v in bs_], 0)
        B, D = temb.shape
        total = w.shape[0]
        out = torch.empty(B, total, dtype=torch.float32, device=temb.device)
        ws = _time_ws(B, D, total, temb.device)
        _lib.check(_lb().ldm_time_proj(temb.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), B, D, total,
                                       ws.data_ptr(), _st()))
        ctx.save_for_backward(temb, w)
        ctx.params = wb
        return out

    @staticmethod
    def backward(ctx, dout):
        temb, w = ctx.saved_tensors
        dout = _c(dout)
        B, D = temb.shape
        total = w.shape[0]
        dw = _zeros(w.numel(), w.device).view_as(w)
        db = _zeros(total, temb.device)
        dtemb = torch.empty_like(temb)
        ws = _time_ws(B, D, total, temb.device)
        _lib.check(_lb().ldm_time_proj_backward(temb.data_ptr(), w.data_ptr(), dout.data_ptr(), dw.data_ptr(), db.data_ptr(),
                                                dtemb.data_ptr(), B, D, total, ws.data_ptr(), _st()))
        grads, srcs, off = [], [], 0
        for wi, bi in zip(ctx.params[0::2], ctx.params[1::2]):
            n = wi.shape[0]
            grads += [_grad_target(wi), _grad_target(bi)]
            srcs += [dw[off:off + n], db[off:off + n]]
            off += n
        torch._foreach_copy_(grads, srcs)          # one multi-tensor launch: every block's rows into that parameter's slot
        return (dtemb, *grads)


class _Add(Function):
    @staticmethod
    def forward(ctx, a, b):
        out = torch.empty_like(a)
        _lib.check(_lb().ldm_add(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), ops._dt(a), _st()))
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


class _Cat(Function):
    """torch.cat((a, b), channel) on NHWC (src/UNet.py:245)."""

    @staticmethod
    def forward(ctx, a, b):
        B, H, W, ca = a.shape
        cb = b.shape[3]
        out = torch.empty(B, H, W, ca + cb, dtype=a.dtype, device=a.device)
        lib, rows = _lb(), B * H * W
        _lib.check(lib.ldm_copy_channels(a.data_ptr(), ca, out.data_ptr(), ca + cb, ca, rows, ops._dt(a), _st()))
        _lib.check(lib.ldm_copy_channels(b.data_ptr(), cb, out.data_ptr() + ca * a.element_size(), ca + cb, cb, rows,
                                         ops._dt(a), _st()))
        ctx.ca, ctx.cb = ca, cb
        return out

    @staticmethod
    def backward(ctx, g):
        g = _c(g)
        B, H, W, ct = g.shape
        ca, cb = ctx.ca, ctx.cb
        da = torch.empty(B, H, W, ca, dtype=g.dtype, device=g.device)
        db = torch.empty(B, H, W, cb, dtype=g.dtype, device=g.device)
        lib, rows = _lb(), B * H * W
        _lib.check(lib.ldm_copy_channels(g.data_ptr(), ct, da.data_ptr(), ca, ca, rows, ops._dt(g), _st()))
        _lib.check(lib.ldm_copy_channels(g.data_ptr() + ca * g.element_size(), ct, db.data_ptr(), cb, cb, rows, ops._dt(g), _st()))
        return da, db


# ------------------------------------------------------------------------------------------------ the network
def _resblock(p, x, tproj_slice, impl):
    """ResNetBlock.forward, src/UNet.py:85-99."""
    h = _GroupNorm.apply(x, p.block1.norm.weight, p.block1.norm.bias, 8, True, None)
    h = _Conv.apply(h, p.block1.conv2d.weight, p.block1.conv2d.bias, impl)
    h = _GroupNorm.apply(h, p.block2.norm.weight, p.block2.norm.bias, 8, True, tproj_slice)   # h + mlp_t(t), then block2
    h = _Conv.apply(h, p.block2.conv2d.weight, p.block2.conv2d.bias, impl)
    sc = _Conv.apply(x, p.shortcut.weight, p.shortcut.bias, impl) if hasattr(p, "shortcut") else x
    return _Add.apply(h, sc)


def _attn_site(site, x, linear, impl):
    """Residual(PreNorm(dim, LinearAttention | Attention)), src/UNet.py:14-20,102-164."""
    pre, att = site.fn, site.fn.fn
    xn = _GroupNorm.apply(x, pre.norm.weight, pre.norm.bias, 1, False, None)
    qkv = _Conv.apply(xn, att.to_qkv.weight, None, impl)
    if linear:
        o = _LinAttn.apply(qkv)
        o = _Conv.apply(o, att.to_out[0].weight, att.to_out[0].bias, impl)
        o = _GroupNorm.apply(o, att.to_out[1].weight, att.to_out[1].bias, 1, False, None)
    else:
        o = _Attn.apply(qkv)
        o = _Conv.apply(o, att.to_out.weight, att.to_out.bias, impl)
    return _Add.apply(o, x)


def _time_slices(blocks, temb):
    """{id(block): [B, cout] slice} of one concatenated time projection over `blocks` (None without a time embedding)."""
    if temb is None or not blocks:
        return {}
    params = []
    for bl in blocks:
        params += [bl.mlp_t[1].weight, bl.mlp_t[1].bias]
    tproj = _TimeProj.apply(temb, *params)
    out, off = {}, 0
    for bl in blocks:
        n = bl.mlp_t[1].weight.shape[0]
        out[id(bl)] = tproj[:, off:off + n]
        off += n
    return out


def _check_inputs(model, x_noisy, t, y):
    dev = x_noisy.device
    for prm in model.parameters():
        if prm.device != dev or prm.dtype != torch.float32:
            raise _lib.LdmError("UNet parameters must be fp32 tensors on the input's CUDA device (call model.to(device))")
    x = _c(x_noisy.detach().to(torch.float32))
    t = _c(t.detach().to(device=dev, dtype=torch.int64))
    if y is not None:
        if model.num_classes is None:
            raise ValueError("labels given but num_classes is None")
        y = _c(y.detach().to(device=dev, dtype=torch.int64))
        if y.numel() not in (1, x.shape[0]):
            raise ValueError("labels must have 1 or batch entries (src/UNet.py:375-376)")
    L = len(model.channel_multipliers)
    if x.shape[2] % (1 << L) != 0 or x.shape[2] != x.shape[3]:
        raise _lib.LdmError(f"image size {tuple(x.shape[2:])} is not divisible by 2^{L} (the reference fails at torch.cat)")
    return x, t, y


def _split_levels(model) -> int:
    """Encoder levels that belong to the FIRST half.  The cut sits after level 1 (of 4): the first half's backward (the two
    full-resolution levels: most of the encoder's time, 3 % of the parameters) then covers the all-reduce of everything
    else, and what is left to reduce after backward is small.  LDM_TRAIN_SPLIT overrides (0 .. number of levels)."""
    import os
    L = len(model.encoder.downs)
    v = os.environ.get("LDM_TRAIN_SPLIT")
    return max(0, min(L, int(v))) if v is not None else min(2, L)


def _first_tail_param(model) -> torch.Tensor:
    k = _split_levels(model)
    if k < len(model.encoder.downs):
        return next(model.encoder.downs[k][0].parameters())
    return next(model.bottleneck.parameters())


def encoder_part(model, x_noisy: torch.Tensor, t: torch.Tensor, y: Optional[torch.Tensor]):
    """First half of UNet.forward (src/UNet.py:361-381): time / label embedding, initial conv, the first encoder levels.
    Returns ``(h, skip_0 .. skip_{k-1}, temb)`` -- everything the second half needs, as tensors, so that the two halves can be
    captured as separate CUDA graphs and the gradient all-reduce of the second half can run under the backward of this one.
    Opens the step: the optimizer's gradient bucket (if any) and the arena are zeroed here."""
    with torch.cuda.device(x_noisy.device):
        x, t, y = _check_inputs(model, x_noisy, t, y)
        dev = x.device
        dt, impl = model.compute_dtype, model.conv_impl
        dt = "bf16" if _lib.DTYPES[dt] == _lib.BF16 else "fp32"
        bucket = getattr(next(model.parameters()), "_ldm_grad_bucket", None)   # trainer.FlatAdam's flat gradient buffer, if any
        if bucket is not None:
            bucket.begin_step()                # one memset of the bucket; the backward kernels accumulate into its slices
        _GradArena.open(model, dev, bucket is not None)
        temb = None
        if model.with_time_emb:
            tm = model.time_emb.time_mlp
            temb = _TimeEmbed.apply(t, y, tm[1].weight, tm[1].bias, tm[3].weight, tm[3].bias,
                                    model.label_emb.weight if (y is not None) else None)
        levels = list(model.encoder.downs)[:_split_levels(model)]
        tsl = _time_slices([lvl[0] for lvl in levels], temb)
        h = _InitialConv.apply(x, model.initial_conv.weight, model.initial_conv.bias, dt)
        skips = []
        for res, attn in levels:
            h = _resblock(res, h, tsl.get(id(res)), impl)
            h = _attn_site(attn, h, True, impl)
            skips.append(h)
            h = _MaxPool.apply(h)
        if temb is None:
            temb = torch.zeros(1, dtype=torch.float32, device=dev)   # placeholder: graph callables exchange tensors only
        return (h, *skips, temb)


def decoder_part(model, h: torch.Tensor, *rest: torch.Tensor):
    """Second half of UNet.forward (src/UNet.py:376-389): the remaining encoder levels, bottleneck, decoder, final conv, from
    ``encoder_part``'s outputs."""
    with torch.cuda.device(h.device):
        impl = model.conv_impl
        skips, temb = list(rest[:-1]), rest[-1]
        levels = list(model.encoder.downs)[_split_levels(model):]
        tsl = _time_slices([lvl[0] for lvl in levels] + [lvl[0] for lvl in model.decoder.ups],
                           temb if model.with_time_emb else None)   # BottleNeck never gets t
        for res, attn in levels:
            h = _resblock(res, h, tsl.get(id(res)), impl)
            h = _attn_site(attn, h, True, impl)
            skips.append(h)
            h = _MaxPool.apply(h)
        h = _resblock(model.bottleneck.res1, h, None, impl)
        h = _attn_site(model.bottleneck.attn, h, False, impl)
        h = _resblock(model.bottleneck.res2, h, None, impl)
        for res, attn, up in model.decoder.ups:
            h = _ConvT.apply(h, up.weight, up.bias, impl)
            h = _Cat.apply(h, skips.pop())
            h = _resblock(res, h, tsl.get(id(res)), impl)
            h = _attn_site(attn, h, True, impl)
        h = _resblock(model.final_conv[0], h, None, impl)
        return _FinalConv.apply(h, model.final_conv[1].weight, model.final_conv[1].bias)


def _overlap_hook(model, boundary: torch.Tensor) -> None:
    """Data parallel: once the gradient of the boundary tensor exists, every gradient of the second half (encoder levels 2+,
    bottleneck, decoder, final conv: 97 % of the bytes, the tail of the flat bucket) is final -- start its all-reduce now, under
    the first half's backward (trainer.FlatAdam.reduce_tail_async; NCCL's stream waits for the work queued so far)."""
    bucket = getattr(next(model.parameters()), "_ldm_grad_bucket", None)
    if bucket is None or not boundary.requires_grad or not bucket.wants_overlap():
        return
    first_tail = _first_tail_param(model)
    boundary.register_hook(lambda g: bucket.reduce_tail_async(first_tail))


def unet_autograd_forward(model, x_noisy: torch.Tensor, t: torch.Tensor, y: Optional[torch.Tensor]):
    """UNet.forward (src/UNet.py:361-389) with gradients to every live parameter.  Runs with the input's device current (the
    kernels launch on that device's stream whatever torch.cuda.current_device() was; autograd does the same for backward)."""
    outs = encoder_part(model, x_noisy, t, y)
    _overlap_hook(model, outs[0])
    return decoder_part(model, *outs)


class _GraphedEncoder(torch.nn.Module):
    """What gets captured: one half of the autograd forward of ``model`` for one call signature.  Wrapper modules of their
    own, because torch.cuda.make_graphed_callables patches ``forward`` of the module it is given -- capturing a second
    signature (with / without labels) on the UNet itself would capture the first one's replay stub."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, x_noisy, t, y=None):
        return encoder_part(self.model, x_noisy, t, y)


class _GraphedDecoder(torch.nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, h, *rest):
        return decoder_part(self.model, h, *rest)


def make_graphed(model, x_noisy: torch.Tensor, t: torch.Tensor, y: Optional[torch.Tensor]):
    """Capture the UNet's training forward and backward as CUDA graphs (torch.cuda.make_graphed_callables): the ~560
    kernel launches of a step are replayed instead of re-issued from Python, which is what bounds small batches.  The two
    halves (encoder | bottleneck + decoder + final conv) are captured as separate graph pairs: between the backward of the
    second half and the backward of the first, a data-parallel job starts the all-reduce of the second half's gradients.
    Returns a callable with the model's signature for inputs of exactly these shapes (labels given or not, as captured);
    gradients land in the same fp32 ``.grad`` tensors.  The four dead bottleneck mlp_t parameters are unused inputs.
    ``model`` itself is left untouched, so several signatures can be captured side by side."""
    args_a = (x_noisy.detach().clone(), t.detach().clone()) + ((y.detach().clone(),) if y is not None else ())
    enc, dec = _GraphedEncoder(model), _GraphedDecoder(model)
    enc.train(model.training)
    dec.train(model.training)
    outs = enc(*args_a)                               # eager, once: the shapes / dtypes of the boundary tensors
    args_b = tuple(torch.zeros_like(o).requires_grad_(o.requires_grad) for o in outs)
    del outs
    # No cyclic garbage collection while a stream is capturing: graph objects, samplers and native UNet handles of models that
    # died earlier are freed with cudaFree / cudaGraphExecDestroy, which invalidates a capture in progress (the collector runs
    # whenever its allocation counters say so, e.g. in the middle of the second half's forward capture).
    import gc
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        g_enc, g_dec = torch.cuda.make_graphed_callables((enc, dec), (args_a, args_b), allow_unused_input=True)
    finally:
        if was_enabled:
            gc.enable()

    def forward(x_noisy, t, y=None):
        outs = g_enc(x_noisy, t, y) if y is not None else g_enc(x_noisy, t)
        _overlap_hook(model, outs[0])
        return g_dec(*outs)

    return forward
