"""GPU parity of the individual kernels (through the C ABI) against the reference's PyTorch ops in fp32
on the same device (TF32 disabled), the CPU oracle and the golden vectors from the unmodified reference.

Tolerances: fp32 kernels 1e-5..1e-4 relative L2 (north-star fp32 bar: 1e-4); bf16 kernels 2e-2
(north-star bf16 bar), in practice ~3e-3 = bf16 output rounding.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden, rel_l2

pytestmark = pytest.mark.gpu
T = torch.from_numpy
TOL = {"fp32": 2e-5, "bf16": 6e-3}


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def dev():
    return torch.device("cuda:0")


def nhwc_ref(x_nchw, dtype):
    """what the kernel sees: NHWC values rounded to the compute dtype, as fp32 NCHW"""
    return x_nchw.to(torch.bfloat16).float() if dtype == "bf16" else x_nchw


# ------------------------------------------------------------------ diffusion element-wise kernels
@pytest.mark.parametrize("n_steps", [1000, 400])
def test_p_sample_golden(n_steps):
    import ldm_b200
    g = golden(f"g2_p_sample_T{n_steps}.npz")
    d = ldm_b200.Diffusion(n_steps, dev())
    xt, eps = T(g["xt"]).to(dev()), T(g["eps"]).to(dev())
    for i, step in enumerate(g["steps"]):
        t = torch.full((4,), int(step), dtype=torch.long, device=dev())
        out = d.p_sample(xt, t, eps, noise=T(g["noise"][i]).to(dev()))
        assert rel_l2(out, T(g["out"][i])) < 1e-6, f"t={step}"
        out1 = d.p_sample(xt, t[:1], eps, noise=T(g["noise"][i]).to(dev()))
        assert torch.equal(out, out1)


def test_p_sample_cfg_and_philox():
    import ldm_b200
    d = ldm_b200.Diffusion(1000, dev())
    g = torch.Generator().manual_seed(0)
    x, ec, eu = (torch.randn(8, 3, 32, 32, generator=g).to(dev()) for _ in range(3))
    t = torch.tensor([500], device=dev())
    z = torch.randn(8, 3, 32, 32, generator=g).to(dev())
    fused = d.p_sample(x, t, ec, noise=z, eps_uncond=eu, cfg_scale=3.0)
    plain = d.p_sample(x, t, torch.lerp(eu, ec, 3.0), noise=z)
    assert rel_l2(fused, plain) < 1e-6
    # in-kernel Philox: deterministic per (seed, sample), N(0,1) moments, invariant to batch sharding
    a = d.p_sample(x, t, ec, seed=123)
    b = d.p_sample(x, t, ec, seed=123)
    assert torch.equal(a, b)
    zhat = (a - d.p_sample(x, t, ec, noise=torch.zeros_like(x))) / float(d.beta[500].sqrt())
    assert abs(float(zhat.mean())) < 0.02 and abs(float(zhat.std()) - 1.0) < 0.02
    # t == 0: no noise
    t0 = torch.tensor([0], device=dev())
    assert torch.equal(d.p_sample(x, t0, ec, seed=1), d.p_sample(x, t0, ec, seed=2))


def test_q_sample_golden():
    import ldm_b200
    g = golden("g3_q_sample.npz")
    d = ldm_b200.Diffusion(1000, dev())
    xt = d.q_sample(T(g["x0"]).to(dev()), T(g["t"]).to(dev()), eps=T(g["noise"]).to(dev()))
    assert rel_l2(xt, T(g["xt"])) < 1e-6
    noise, xt2, t = d(T(g["x0"]).to(dev()))
    assert noise.shape == xt2.shape and t.shape == (8,) and t.dtype == torch.int64
    assert abs(float(noise.std()) - 1.0) < 0.03
    ab = d.alpha_bar[t].reshape(-1, 1, 1, 1)
    assert rel_l2(xt2, ab.sqrt() * T(g["x0"]).to(dev()) + (1 - ab).sqrt() * noise) < 1e-6


# ------------------------------------------------------------------ GroupNorm (+SiLU, +residual)
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("C,HW,G,silu", [(64, 32, 8, True), (128, 16, 8, True), (768, 4, 8, True), (192, 16, 8, True),
                                        (384, 8, 8, True), (512, 2, 8, True), (64, 32, 1, False), (512, 2, 1, False),
                                        (256, 8, 1, False)])
def test_group_norm(dtype, C, HW, G, silu):
    from ldm_b200 import ops
    g = torch.Generator().manual_seed(C + HW)
    B = 3
    x = (torch.randn(B, C, HW, HW, generator=g) * 2 + 0.5).to(dev())
    gamma = torch.randn(C, generator=g).to(dev())
    beta = torch.randn(C, generator=g).to(dev())
    res = torch.randn(B, C, HW, HW, generator=g).to(dev())
    xh = ops.to_nhwc(x, dtype)
    ref = F.group_norm(nhwc_ref(x, dtype), G, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    out = ops.to_nchw(ops.group_norm(xh, gamma, beta, G, silu=silu))
    assert rel_l2(out, ref) < TOL[dtype]
    # residual + strided output slice
    wide = torch.zeros(B, HW, HW, C + 64, dtype=xh.dtype, device=dev())
    rh = ops.to_nhwc(res, dtype)
    ops.group_norm(xh, gamma, beta, G, silu=silu, res=rh, out=wide[..., 64:])
    out2 = ops.to_nchw(wide[..., 64:], channels=C, ld=C + 64)
    assert rel_l2(out2, ref + nhwc_ref(res, dtype)) < TOL[dtype]
    assert float(wide[..., :64].float().abs().max()) == 0.0


# ------------------------------------------------------------------ convolutions
def _conv_case(dtype, impl, B, cin, cout, R, k, second=0, rowvec=False, res=False, seed=0):
    from ldm_b200 import ops
    g = torch.Generator().manual_seed(seed + cin + cout + R)
    x = torch.randn(B, cin, R, R, generator=g).to(dev())
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev())
    b = torch.randn(cout, generator=g).to(dev())
    xq, wq = nhwc_ref(x, dtype), nhwc_ref(w, dtype)
    ref = F.conv2d(xq, wq, b, padding=k // 2)
    x2h = w2 = None
    if second:
        x2 = torch.randn(B, second, R, R, generator=g).to(dev())
        w2 = (torch.randn(cout, second, 1, 1, generator=g) / second ** 0.5).to(dev())
        ref = ref + F.conv2d(nhwc_ref(x2, dtype), nhwc_ref(w2, dtype))
        x2h = ops.to_nhwc(x2, dtype)
    rv = None
    if rowvec:
        rv = torch.randn(B, cout + 64, generator=g).to(dev())[:, 64:]
        ref = ref + rv[:, :, None, None]
    rh = None
    if res:
        r = torch.randn(B, cout, R, R, generator=g).to(dev())
        ref = ref + nhwc_ref(r, dtype)
        rh = ops.to_nhwc(r, dtype)
    wp = ops.pack_conv_weight(w, dtype, w2)
    out = ops.conv2d(ops.to_nhwc(x, dtype), wp, k, bias=b, x2=x2h, rowvec=rv, res=rh, impl=impl)
    return ops.to_nchw(out), ref


CONV_CASES = [
    # B, cin, cout, R, k, second, rowvec, res
    (2, 64, 64, 32, 3, 0, True, False),
    (2, 64, 64, 32, 3, 0, False, True),
    (3, 64, 128, 16, 3, 0, True, False),
    (3, 128, 128, 16, 3, 64, False, False),
    (2, 256, 512, 4, 3, 0, True, False),
    (5, 512, 512, 2, 3, 0, False, True),
    (2, 768, 256, 4, 3, 0, True, False),
    (2, 256, 256, 4, 3, 768, False, False),
    (2, 64, 384, 32, 1, 0, False, False),
    (2, 128, 64, 32, 1, 0, False, False),
    (3, 512, 384, 2, 1, 0, False, False),
    (1, 128, 512, 2, 1, 0, False, True),
    (1, 64, 64, 8, 3, 0, False, False),
    # full-resolution layers of the decoder / final block (halo-slab kernel: resident filter, 2-source K concat)
    (2, 128, 64, 32, 3, 0, True, False),
    (3, 64, 64, 32, 3, 128, False, False),
    (40, 64, 64, 32, 3, 0, True, False),     # 360 tiles: several tiles per persistent CTA, ring wrap-around
    (37, 128, 64, 32, 3, 0, False, True),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fp32_simt(case):
    out, ref = _conv_case("fp32", 0, *case)
    assert rel_l2(out, ref) < 2e-5


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_bf16_tcgen05(case):
    out, ref = _conv_case("bf16", 0, *case)      # tcgen05 / TMEM / TMA
    assert rel_l2(out, ref) < 4e-3, "tcgen05 conv vs F.conv2d on bf16-rounded operands"
    out_s, _ = _conv_case("bf16", 1, *case)      # same inputs through the FFMA kernel
    assert rel_l2(out, out_s) < 3e-3


# ------------------------------------------------------------------ GroupNorm fused into the conv epilogue
# (csrc/conv_epilogue.cuh): Block = GroupNorm(8) -> SiLU on conv1's output incl. the time-embedding row (src/UNet.py:52-58,
# 88-93), LinearAttention.to_out's GroupNorm(1, C) + Residual (:147,:20), and PreNorm statistics (:106)
GN_FUSE_CASES = [
    # B, cin, cout, R, k, groups, silu, rowvec, gn_res, second
    (3, 64, 64, 32, 3, 8, True, True, False, 0),       # halo kernel: 9 tiles per sample exchange packets
    (40, 64, 64, 32, 3, 8, True, True, False, 0),      # several tiles per persistent CTA, deferred second pass wraps the ring
    (2, 128, 64, 32, 3, 8, True, False, False, 0),
    (3, 64, 128, 16, 3, 8, True, True, False, 0),      # conv_tc, two M tiles per sample
    (300, 64, 128, 16, 3, 8, True, True, False, 0),    # paired M tiles (MT = 2): a work unit is one whole sample
    (3, 192, 64, 16, 3, 8, True, True, False, 0),
    (5, 128, 256, 8, 3, 8, True, True, False, 0),      # two samples per tile
    (5, 256, 512, 4, 3, 8, True, True, False, 0),      # eight samples per tile, several N tiles
    (20, 768, 256, 4, 3, 8, True, True, False, 0),
    (3, 128, 64, 32, 1, 1, False, False, True, 0),     # to_out: GroupNorm(1, C) + residual, 8 tiles per sample
    (160, 128, 64, 32, 1, 1, False, False, True, 0),   # ... paired tiles
    (3, 128, 128, 16, 1, 1, False, False, True, 0),
    (5, 128, 256, 8, 1, 1, False, False, True, 0),
    (9, 128, 512, 4, 1, 1, False, False, True, 0),     # GroupNorm(1, 512) over four N tiles of eight samples
    (80, 128, 512, 4, 1, 1, False, False, True, 0),
]


def _gn_fuse_inputs(B, cin, cout, R, k, second, seed=0):
    from ldm_b200 import ops
    g = torch.Generator().manual_seed(seed + cin + cout + R + B)
    x = torch.randn(B, cin, R, R, generator=g).to(dev())
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev())
    b = torch.randn(cout, generator=g).to(dev())
    conv = F.conv2d(nhwc_ref(x, "bf16"), nhwc_ref(w, "bf16"), b, padding=k // 2)
    x2h = w2 = None
    if second:
        x2 = torch.randn(B, second, R, R, generator=g).to(dev())
        w2 = (torch.randn(cout, second, 1, 1, generator=g) / second ** 0.5).to(dev())
        conv = conv + F.conv2d(nhwc_ref(x2, "bf16"), nhwc_ref(w2, "bf16"))
        x2h = ops.to_nhwc(x2, "bf16")
    return g, ops.to_nhwc(x, "bf16"), ops.pack_conv_weight(w, "bf16", w2), b, x2h, conv


@pytest.mark.parametrize("case", GN_FUSE_CASES)
def test_conv_gn_fused_normalise(case):
    from ldm_b200 import ops
    B, cin, cout, R, k, G, silu, rowvec, gn_res, second = case
    g, xh, wp, b, x2h, conv = _gn_fuse_inputs(B, cin, cout, R, k, second)
    gamma = (1 + 0.3 * torch.randn(cout, generator=g)).to(dev())
    beta = (0.3 * torch.randn(cout, generator=g)).to(dev())
    rv = torch.randn(B, cout + 64, generator=g).to(dev())[:, 64:] if rowvec else None
    pre = conv + (rv[:, :, None, None] if rowvec else 0)
    ref = F.group_norm(pre, G, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    rh = None
    if gn_res:
        r = torch.randn(B, cout, R, R, generator=g).to(dev())
        ref = ref + nhwc_ref(r, "bf16")
        rh = ops.to_nhwc(r, "bf16")
    # strided output slice, as the skip half of the decoder's concat buffer
    wide = torch.zeros(B, R, R, cout + 64, dtype=torch.bfloat16, device=dev())
    ops.conv2d_gn(xh, wp, k, b, mode=2, groups=G, gamma=gamma, beta=beta, silu=silu, x2=x2h, gn_rowvec=rv, gn_res=rh,
                  out=wide[..., 64:])
    out = ops.to_nchw(wide[..., 64:], channels=cout, ld=cout + 64)
    assert rel_l2(out, ref) < 6e-3
    assert float(wide[..., :64].float().abs().max()) == 0.0
    # same bits whatever the batch the sample sits in (fixed-order sums): the first sample alone
    one = ops.conv2d_gn(xh[:1].contiguous(), wp, k, b, mode=2, groups=G, gamma=gamma, beta=beta, silu=silu,
                        x2=x2h[:1].contiguous() if x2h is not None else None, gn_rowvec=rv[:1] if rowvec else None,
                        gn_res=rh[:1].contiguous() if rh is not None else None)
    assert torch.equal(one[0], wide[0, ..., 64:])


def test_conv_gn_fused_two_variants():
    """conv1 of the first ResNetBlock runs once for the cond / uncond halves; its epilogue normalises twice."""
    from ldm_b200 import ops
    B, cin, cout, R, k = 5, 64, 64, 32, 3
    g, xh, wp, b, _, conv = _gn_fuse_inputs(B, cin, cout, R, k, 0)
    gamma = (1 + 0.3 * torch.randn(cout, generator=g)).to(dev())
    beta = (0.3 * torch.randn(cout, generator=g)).to(dev())
    rv = torch.randn(2 * B, cout, generator=g).to(dev())
    out = ops.conv2d_gn(xh, wp, k, b, mode=2, groups=8, gamma=gamma, beta=beta, silu=True, gn_rowvec=rv, nvar=2)
    ref = F.silu(F.group_norm(torch.cat([conv, conv]) + rv[:, :, None, None], 8, gamma, beta, 1e-5))
    assert rel_l2(ops.to_nchw(out), ref) < 6e-3


@pytest.mark.parametrize("case", [(3, 64, 64, 32, 3, 0, True), (3, 64, 64, 32, 3, 128, False), (300, 128, 128, 16, 3, 64, False),
                                  (5, 256, 256, 8, 3, 128, False), (9, 512, 512, 4, 3, 256, False), (2, 256, 256, 4, 3, 768, False)])
def test_conv_gn_fused_statistics(case):
    """conv2 (+ shortcut / residual) leaves GroupNorm(1, C) partial sums of its output for the PreNorm that follows."""
    from ldm_b200 import ops
    B, cin, cout, R, k, second, res = case
    g, xh, wp, b, x2h, conv = _gn_fuse_inputs(B, cin, cout, R, k, second)
    rh = None
    if res:
        r = torch.randn(B, cout, R, R, generator=g).to(dev())
        conv = conv + nhwc_ref(r, "bf16")
        rh = ops.to_nhwc(r, "bf16")
    y, stats = ops.conv2d_gn(xh, wp, k, b, mode=1, groups=1, x2=x2h, res=rh)
    assert rel_l2(ops.to_nchw(y), conv) < 4e-3
    s = stats.double().sum(dim=2)[:, 0]                       # [B, 2]
    n = cout * R * R
    mean = s[:, 0] / n
    var = s[:, 1] / n - mean ** 2
    ref_mean = conv.double().mean(dim=(1, 2, 3))
    ref_var = conv.double().var(dim=(1, 2, 3), unbiased=False)
    assert float((mean - ref_mean).abs().max()) < 2e-3 * float(ref_var.sqrt().max())
    assert float((var / ref_var - 1).abs().max()) < 2e-3


@pytest.mark.parametrize("dtype,impl", [("fp32", 0), ("bf16", 0), ("bf16", 1)])
@pytest.mark.parametrize("B,cin,cout,R", [(2, 512, 256, 2), (3, 64, 64, 16), (2, 128, 64, 8)])
def test_conv_transpose(dtype, impl, B, cin, cout, R):
    from ldm_b200 import ops
    g = torch.Generator().manual_seed(cin + R)
    x = torch.randn(B, cin, R, R, generator=g).to(dev())
    w = (torch.randn(cin, cout, 2, 2, generator=g) / cin ** 0.5).to(dev())
    b = torch.randn(cout, generator=g).to(dev())
    ref = F.conv_transpose2d(nhwc_ref(x, dtype), nhwc_ref(w, dtype), b, stride=2)
    # write into the first channels of a wider concat buffer, as the decoder does
    wide = torch.zeros(B, 2 * R, 2 * R, cout + 128, dtype=torch.float32 if dtype == "fp32" else torch.bfloat16, device=dev())
    ops.conv_transpose2x2(ops.to_nhwc(x, dtype), w, b, impl=impl, out=wide[..., :cout])
    out = ops.to_nchw(wide[..., :cout], channels=cout, ld=cout + 128)
    assert rel_l2(out, ref) < TOL[dtype]
    assert float(wide[..., cout:].float().abs().max()) == 0.0


# ------------------------------------------------------------------ attention cores, pooling
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["linattn_64_n1024", "linattn_128_n16", "attn_512_n4", "attn_64_n64"])
def test_attention_cores(dtype, name):
    from ldm_b200 import ops
    from oracle import unet_oracle as U
    g = golden(f"g4_{name}.npz")
    x = T(g["in0"]).to(dev())
    wqkv = T(g["w::to_qkv.weight"]).to(dev())
    qkv = F.conv2d(x, wqkv)                                   # [B,384,H,W] fp32, library op only to make inputs
    qh = ops.to_nhwc(qkv, dtype)
    qkv_q = nhwc_ref(qkv, dtype)
    b, _, hh, ww = qkv.shape
    q, k, v = U._split_heads(qkv_q)
    if name.startswith("lin"):
        out = ops.to_nchw(ops.linear_attention(qh))
        q = q.softmax(dim=-2) * 32 ** -0.5
        k = k.softmax(dim=-1)
        ctx = torch.einsum("bhdn,bhen->bhde", k, v)
        ref = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b, 128, hh, ww)
    else:
        out = ops.to_nchw(ops.attention(qh))
        sim = torch.einsum("bhdi,bhdj->bhij", q * 32 ** -0.5, k)
        attn = (sim - sim.amax(dim=-1, keepdim=True)).softmax(dim=-1)
        ref = torch.einsum("bhij,bhdj->bhid", attn, v).permute(0, 1, 3, 2).reshape(b, 128, hh, ww)
    assert rel_l2(out, ref) < TOL[dtype]


def _linattn_prenorm_reference(x, w, gamma, beta):
    """src/UNet.py:106-110 (PreNorm GroupNorm(1, C)), :145 (to_qkv), :149-163 (LinearAttention up to `out`), fp32."""
    B, C, H, W = x.shape
    xn = F.group_norm(x, 1, gamma, beta, 1e-5)
    qkv = F.conv2d(xn, w).reshape(B, 3, 4, 32, H * W)          # "b (h c) x y -> b h c (x y)", chunks q | k | v
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    q = q.softmax(dim=-2) * 32 ** -0.5
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q)
    return out.reshape(B, 128, H, W)


@pytest.mark.parametrize("B,R", [(3, 32), (160, 32), (5, 16)])
def test_linear_attention_prenorm_to_out_fused(B, R):
    """... and with to_out.0 folded in (y = conv1x1(attention) + bias, applied to the per-sample context matrix) plus the partial
    sums the following GroupNorm(1, C) needs."""
    from ldm_b200 import ops
    g = torch.Generator().manual_seed(B * 7 + R)
    x = (torch.randn(B, 64, R, R, generator=g) * 1.7 + 0.4).to(dev())
    w = (torch.randn(384, 64, 1, 1, generator=g) * 0.25).to(dev())
    gamma = (1 + 0.3 * torch.randn(64, generator=g)).to(dev())
    beta = (0.3 * torch.randn(64, generator=g)).to(dev())
    wo = (torch.randn(64, 128, 1, 1, generator=g) / 128 ** 0.5).to(dev())
    bo = torch.randn(64, generator=g).to(dev())
    ref = F.conv2d(_linattn_prenorm_reference(nhwc_ref(x, "bf16"), w, gamma, beta), wo, bo)
    y, stats = ops.linear_attention_prenorm_to_out(ops.to_nhwc(x, "bf16"), w, gamma, beta, wo, bo)
    got = ops.to_nchw(y)
    assert rel_l2(got, ref) < 1.2e-2
    s = stats.double().sum(dim=1)                               # [B, 2]
    n = 64 * R * R
    mean, var = s[:, 0] / n, s[:, 1] / n - (s[:, 0] / n) ** 2
    assert float((mean - got.double().mean(dim=(1, 2, 3))).abs().max()) < 2e-3 * float(got.double().std())
    assert float((var / got.double().var(dim=(1, 2, 3), unbiased=False) - 1).abs().max()) < 5e-3


def test_linear_attention_to_out_exact_fallback():
    """linattn_tc2_kernel shifts the softmax over tokens by the first token and evaluates the per-head softmax unshifted; a logit
    88 away from its reference overflows, the sample is flagged and linattn_tc_kernel (two-pass softmax) redoes it in the same
    call.  Sample 1 is scaled so that its logits span hundreds: it must come out finite and equal to what the exact kernel alone
    produces (LDM_LINATTN_V1 is read once per process, so 'alone' = the unfused op + a 1x1 conv in torch), the others untouched."""
    from ldm_b200 import ops
    g = torch.Generator().manual_seed(11)
    B, R = 4, 16
    x = torch.randn(B, 64, R, R, generator=g)
    x[1, :, 0, 0] *= -3.0                                        # first token far from the rest of the sample
    x = x.to(dev())
    w = (torch.randn(384, 64, 1, 1, generator=g) * 6.0).to(dev())   # |q|, |k| ~ 50 sigma: k range > 88 within a channel
    gamma = torch.ones(64, device=dev())
    beta = torch.zeros(64, device=dev())
    wo = (torch.randn(64, 128, 1, 1, generator=g) / 128 ** 0.5).to(dev())
    bo = torch.randn(64, generator=g).to(dev())
    y, stats = ops.linear_attention_prenorm_to_out(ops.to_nhwc(x, "bf16"), w, gamma, beta, wo, bo)
    got = ops.to_nchw(y)
    assert bool(torch.isfinite(got).all()) and bool(torch.isfinite(stats).all())
    att = ops.to_nchw(ops.linear_attention_prenorm(ops.to_nhwc(x, "bf16"), w, gamma, beta, impl=0))   # exact tcgen05 kernel
    ref = F.conv2d(att, wo, bo)
    assert rel_l2(got, ref) < 2e-2


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("B,R", [(3, 32), (160, 32), (5, 16)])
def test_linear_attention_prenorm_fused(impl, B, R):
    """PreNorm + to_qkv + LinearAttention in one kernel on the raw block input: tcgen05 / TMEM (impl 0) and mma.sync (impl 1)."""
    from ldm_b200 import ops
    g = torch.Generator().manual_seed(B + R)
    x = (torch.randn(B, 64, R, R, generator=g) * 1.7 + 0.4).to(dev())
    w = (torch.randn(384, 64, 1, 1, generator=g) * 0.25).to(dev())
    gamma = (1 + 0.3 * torch.randn(64, generator=g)).to(dev())
    beta = (0.3 * torch.randn(64, generator=g)).to(dev())
    ref = _linattn_prenorm_reference(nhwc_ref(x, "bf16"), w, gamma, beta)
    out = ops.to_nchw(ops.linear_attention_prenorm(ops.to_nhwc(x, "bf16"), w, gamma, beta, impl=impl))
    assert rel_l2(out, ref) < 1.2e-2, "bf16 operands (x, folded weights, P, V, ctx, softmax(q)) against the fp32 reference"
    if impl == 0:   # the two kernels round at the same places
        other = ops.to_nchw(ops.linear_attention_prenorm(ops.to_nhwc(x, "bf16"), w, gamma, beta, impl=1))
        assert rel_l2(out, other) < 6e-3


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_max_pool(dtype):
    from ldm_b200 import ops
    x = torch.randn(3, 128, 16, 16, generator=torch.Generator().manual_seed(3)).to(dev())
    out = ops.to_nchw(ops.max_pool2x2(ops.to_nhwc(x, dtype)))
    assert torch.equal(out, F.max_pool2d(nhwc_ref(x, dtype), 2, 2))


def test_errors_are_reported_not_fatal():
    from ldm_b200 import ops, _lib
    x = torch.zeros(1, 8, 8, 24, device=dev())
    with pytest.raises(_lib.LdmError):
        ops.conv2d(x, torch.zeros(64, 24, device=dev()), 1)          # Cin not a multiple of 16
    with pytest.raises(_lib.LdmError):
        ops.group_norm(x, torch.ones(24, device=dev()), torch.zeros(24, device=dev()), 5)
    with pytest.raises(_lib.LdmError):
        ops.to_nhwc(torch.zeros(1, 3, 4, 4))                         # host tensor: no CPU fallback
