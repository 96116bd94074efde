"""GPU parity of the backward kernels, one at a time, against torch.autograd over the reference's PyTorch ops
(fp32, TF32 off) on the same bf16-/fp32-rounded inputs.  They are reached through the autograd.Function wrappers of
ldm_b200/train.py, i.e. through the C ABI entry points of the "training step" section of include/ldm_b200.h."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = {"fp32": 5e-5, "bf16": 2e-2}


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def dev():
    return torch.device("cuda:0")


def rnd(dtype, x):
    return x.to(torch.bfloat16).float() if dtype == "bf16" else x


def nhwc(x, dtype):
    from ldm_b200 import ops
    return ops.to_nhwc(x, dtype)


def nchw(x):
    from ldm_b200 import ops
    return ops.to_nchw(x.contiguous())


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("B,cin,cout,R,k,bias", [(2, 64, 128, 16, 3, True), (3, 128, 64, 32, 3, True), (2, 64, 384, 8, 1, False),
                                                 (5, 512, 512, 2, 3, True), (2, 768, 256, 4, 3, True),
                                                 # below 8x8 with B*H*W % 8 == 0: the batch-flattened tcgen05 wgrad
                                                 (16, 512, 512, 2, 3, True), (24, 768, 256, 4, 3, True),
                                                 (10, 256, 1024, 4, 1, False), (6, 128, 192, 4, 3, True)])
def test_conv_backward(dtype, B, cin, cout, R, k, bias):
    from ldm_b200 import train
    g = torch.Generator().manual_seed(cin + cout + R)
    x = rnd(dtype, torch.randn(B, cin, R, R, generator=g).to(dev()))
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev())
    b = torch.randn(cout, generator=g).to(dev()) if bias else None
    dy = rnd(dtype, torch.randn(B, cout, R, R, generator=g).to(dev()))
    xr = x.clone().requires_grad_(True)
    wr = rnd(dtype, w).clone().requires_grad_(True)
    br = b.clone().requires_grad_(True) if bias else None
    F.conv2d(xr, wr, br, padding=k // 2).backward(dy)
    xh = nhwc(x, dtype).requires_grad_(True)
    wp = w.clone().requires_grad_(True)
    bp = b.clone().requires_grad_(True) if bias else None
    y = train._Conv.apply(xh, wp, bp, 0)
    y.backward(nhwc(dy, dtype))
    assert rel_l2(nchw(xh.grad), xr.grad) < TOL[dtype]
    assert rel_l2(wp.grad, wr.grad) < TOL[dtype]
    if bias:
        assert rel_l2(bp.grad, br.grad) < TOL[dtype]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("C,R,G,silu,rv", [(64, 32, 8, True, True), (192, 16, 8, True, False), (768, 4, 8, True, True),
                                           (512, 2, 1, False, False), (128, 16, 1, False, False)])
def test_group_norm_backward(dtype, C, R, G, silu, rv):
    from ldm_b200 import train
    g = torch.Generator().manual_seed(C + R)
    B = 3
    x = rnd(dtype, (torch.randn(B, C, R, R, generator=g) * 1.5 + 0.3).to(dev()))
    gamma = (torch.randn(C, generator=g) * 0.5 + 1).to(dev())
    beta = torch.randn(C, generator=g).to(dev())
    rowvec = torch.randn(B, C + 32, generator=g).to(dev())[:, 32:] if rv else None
    dy = rnd(dtype, torch.randn(B, C, R, R, generator=g).to(dev()))
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rr = rowvec.clone().requires_grad_(True) if rv else None
    z = F.group_norm(xr + rr[:, :, None, None] if rv else xr, G, gr, br, 1e-5)
    (F.silu(z) if silu else z).backward(dy)
    xh = nhwc(x, dtype).requires_grad_(True)
    gp, bp = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rp = rowvec.detach().clone().requires_grad_(True) if rv else None
    train._GroupNorm.apply(xh, gp, bp, G, silu, rp).backward(nhwc(dy, dtype))
    assert rel_l2(nchw(xh.grad), xr.grad) < TOL[dtype] * 2
    assert rel_l2(gp.grad, gr.grad) < TOL[dtype] * 2
    assert rel_l2(bp.grad, br.grad) < TOL[dtype] * 2
    if rv:
        assert rel_l2(rp.grad, rr.grad) < TOL[dtype] * 4


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("N,linear", [(1024, True), (16, True), (64, True), (144, True), (256, True), (4, False), (64, False)])
def test_attention_backward(dtype, N, linear):
    from ldm_b200 import train
    from oracle import unet_oracle as U
    g = torch.Generator().manual_seed(N)
    R = int(N ** 0.5)
    B = 2
    qkv = rnd(dtype, torch.randn(B, 384, R, R, generator=g).to(dev()))
    dout = rnd(dtype, torch.randn(B, 128, R, R, generator=g).to(dev()))
    qr = qkv.clone().requires_grad_(True)
    q, k, v = U._split_heads(qr)
    if linear:
        q = q.softmax(dim=-2) * 32 ** -0.5
        k = k.softmax(dim=-1)
        ctx = torch.einsum("bhdn,bhen->bhde", k, v)
        ref = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(B, 128, R, R)
    else:
        sim = torch.einsum("bhdi,bhdj->bhij", q * 32 ** -0.5, k)
        attn = (sim - sim.amax(dim=-1, keepdim=True)).softmax(dim=-1)
        ref = torch.einsum("bhij,bhdj->bhid", attn, v).permute(0, 1, 3, 2).reshape(B, 128, R, R)
    ref.backward(dout)
    qh = nhwc(qkv, dtype).requires_grad_(True)
    (train._LinAttn if linear else train._Attn).apply(qh).backward(nhwc(dout, dtype))
    assert rel_l2(nchw(qh.grad), qr.grad) < (1e-4 if dtype == "fp32" else 3e-2)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_pool_convT_cat_add_backward(dtype):
    from ldm_b200 import train
    g = torch.Generator().manual_seed(8)
    x = rnd(dtype, torch.randn(2, 64, 16, 16, generator=g).to(dev()))
    dy = rnd(dtype, torch.randn(2, 64, 8, 8, generator=g).to(dev()))
    xr = x.clone().requires_grad_(True)
    F.max_pool2d(xr, 2, 2).backward(dy)
    xh = nhwc(x, dtype).requires_grad_(True)
    train._MaxPool.apply(xh).backward(nhwc(dy, dtype))
    assert torch.equal(nchw(xh.grad), xr.grad)
    # ConvTranspose2d(k2, s2)
    w = (torch.randn(64, 128, 2, 2, generator=g) / 8).to(dev())
    b = torch.randn(128, generator=g).to(dev())
    dyt = rnd(dtype, torch.randn(2, 128, 32, 32, generator=g).to(dev()))
    xr = x.clone().requires_grad_(True)
    wr, br = rnd(dtype, w).clone().requires_grad_(True), b.clone().requires_grad_(True)
    F.conv_transpose2d(xr, wr, br, stride=2).backward(dyt)
    xh = nhwc(x, dtype).requires_grad_(True)
    wp, bp = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    train._ConvT.apply(xh, wp, bp, 0).backward(nhwc(dyt, dtype))
    assert rel_l2(nchw(xh.grad), xr.grad) < TOL[dtype]
    assert rel_l2(wp.grad, wr.grad) < TOL[dtype]
    assert rel_l2(bp.grad, br.grad) < TOL[dtype]
    # cat + add
    a = nhwc(x, dtype).requires_grad_(True)
    c = nhwc(x * 2, dtype).requires_grad_(True)
    out = train._Cat.apply(train._Add.apply(a, c), c)
    assert rel_l2(nchw(out), torch.cat((x + rnd(dtype, x * 2), rnd(dtype, x * 2)), 1)) < TOL[dtype]
    out.backward(torch.ones_like(out))
    assert float((a.grad.float() - 1).abs().max()) == 0 and float((c.grad.float() - 2).abs().max()) == 0


def test_time_embedding_backward():
    from ldm_b200 import train
    from oracle import unet_oracle as U
    g = torch.Generator().manual_seed(1)
    D, B = 256, 5
    w1 = (torch.randn(D, D // 4, generator=g) / 8).to(dev()); b1 = torch.randn(D, generator=g).to(dev())
    w3 = (torch.randn(D, D, generator=g) / 16).to(dev()); b3 = torch.randn(D, generator=g).to(dev())
    label = torch.randn(10, D, generator=g).to(dev())
    wc = (torch.randn(320, D, generator=g) / 16).to(dev()); bc = torch.randn(320, generator=g).to(dev())
    t = torch.tensor([0, 3, 250, 999, 77], device=dev()); y = torch.tensor([1, 1, 9, 0, 4], device=dev())
    dout = torch.randn(B, 320, generator=g).to(dev())
    ps = [p.clone().requires_grad_(True) for p in (w1, b1, w3, b3, label, wc, bc)]
    if True:
        half = D // 8
        f = torch.exp(torch.arange(half, device=dev()) * -(torch.log(torch.tensor(10000.0)) / (half - 1)))
        e = t[:, None].float() * f[None, :]
        emb = torch.cat((e.sin(), e.cos()), -1)
    temb = F.linear(F.gelu(F.linear(emb, ps[0], ps[1])), ps[2], ps[3]) + ps[4][y]
    F.linear(F.silu(temb), ps[5], ps[6]).backward(dout)
    qs = [p.clone().requires_grad_(True) for p in (w1, b1, w3, b3, label, wc, bc)]
    te = train._TimeEmbed.apply(t, y, qs[0], qs[1], qs[2], qs[3], qs[4])
    assert rel_l2(te, temb) < 1e-5
    train._TimeProj.apply(te, qs[5], qs[6]).backward(dout)
    for a, b_ in zip(qs, ps):
        assert rel_l2(a.grad, b_.grad) < 1e-4
