"""GPU: the first-stage autoencoder and the latent diffusion model (BASELINE config 4) through the C ABI, against golden
vectors of the unmodified reference (oracle/make_golden_next.py: G10, G11) and torch fp32 restatements of the new kernels."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden, rel_l2
from test_unet_gpu import dev

pytestmark = pytest.mark.gpu
T = torch.from_numpy
TOL = {"fp32": 1e-4, "bf16": 2e-2}


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _ae(g, dtype):
    import ldm_b200
    cfg = [int(v) for v in g["config"]]
    torch.manual_seed(int(g["weight_seed"]))
    return ldm_b200.Autoencoder(cfg[0], cfg[1], cfg[2], cfg[3], cfg[5:], cfg[4], dtype=dtype).to(dev())


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("tag", ["ldm", "deep"])
def test_autoencoder_matches_reference_golden(tag, dtype):
    g = golden(f"g10_autoencoder_{tag}.npz")
    ae = _ae(g, dtype)
    img = T(g["img"]).to(dev())
    dist = ae.encode(img, epsilon=T(g["epsilon"]).to(dev()))
    tol = TOL[dtype]
    assert rel_l2(dist.mu.cpu(), T(g["mu"])) < tol
    assert rel_l2(dist.log_var.cpu(), T(g["log_var"])) < tol
    assert rel_l2(dist.sigma.cpu(), torch.exp(T(g["log_var"]) / 2)) < tol
    assert rel_l2(dist.sample().cpu(), T(g["z"])) < tol
    assert ae.last_launches > 20
    rec = ae.decode(T(g["z"]).to(dev()))                    # teacher-forced: the reference's own latent
    assert rec.shape == img.shape and rel_l2(rec.cpu(), T(g["recon"])) < tol
    with pytest.raises(Exception, match="inference-only"):   # backward does not exist: the module says so
        ae(img, epsilon=T(g["epsilon"]).to(dev()))
    with torch.no_grad():
        out, mu, lv = ae(img, epsilon=T(g["epsilon"]).to(dev()))
    assert rel_l2(out.cpu(), T(g["forward_img"])) < 2 * tol
    assert torch.equal(mu, ae.distribution.mu)
    d2 = ae.encode(img)                                       # own noise: a different sample, same moments
    assert torch.equal(d2.mu, dist.mu) and not torch.equal(d2.sample(), dist.sample())


def test_autoencoder_has_no_cpu_path_and_keeps_the_state_dict_contract():
    import ldm_b200
    from ldm_b200 import _lib
    torch.manual_seed(0)
    ae = ldm_b200.Autoencoder(3, 4, 3, 64, [1, 2], 2)
    keys = list(ae.state_dict().keys())
    assert keys[0] == "encoder.conv_in.weight" and "encoder.down.0.downsample.conv.weight" in keys
    assert "encoder.down.1.downsample.conv.weight" not in keys            # Identity at the last level
    assert "decoder.up.1.upsample.conv.weight" in keys and "decoder.up.0.upsample.conv.weight" not in keys
    assert {"quant_conv.weight", "post_quant_conv.bias", "encoder.mid.attn_1.proj_out.weight"} <= set(keys)
    with pytest.raises(_lib.LdmError):
        ae.encode(torch.zeros(1, 3, 32, 32))


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("C,R", [(64, 32), (128, 16), (256, 8), (64, 3)])
def test_group_norm_32_groups(dtype, C, R):
    from ldm_b200 import ops
    gen = torch.Generator().manual_seed(C + R)
    x = (torch.randn(3, C, R, R, generator=gen) * 1.7 + 0.4).to(dev())
    gamma, beta = (torch.randn(C, generator=gen) * 0.5 + 1).to(dev()), torch.randn(C, generator=gen).to(dev())
    xh = ops.to_nhwc(x, dtype)
    xr = ops.to_nchw(xh)                                      # the rounded input both sides see
    for silu in (False, True):
        want = F.group_norm(xr, 32, gamma, beta, eps=1e-6)
        want = want * torch.sigmoid(want) if silu else want
        got = ops.to_nchw(ops.group_norm(xh, gamma, beta, 32, eps=1e-6, silu=silu))
        assert rel_l2(got, want) < (1e-5 if dtype == "fp32" else 6e-3)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_upsample_and_stride2_pick(dtype):
    from ldm_b200 import ops, _lib
    lib = _lib.load()
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 64, 8, 8, generator=gen).to(dev())
    xh = ops.to_nhwc(x, dtype)
    xr = ops.to_nchw(xh)
    up = torch.empty(2, 16, 16, 64, dtype=xh.dtype, device=dev())
    _lib.check(lib.ldm_upsample_nearest2x(xh.data_ptr(), 64, up.data_ptr(), 64, 2, 8, 8, 64, ops._dt(xh), _lib.stream_ptr()))
    assert torch.equal(ops.to_nchw(up), F.interpolate(xr, scale_factor=2.0, mode="nearest"))
    # DownSample: pad (0,1,0,1) + stride-2 conv == odd positions of the pad-1 stride-1 conv
    w = (torch.randn(64, 64, 3, 3, generator=gen) / 24).to(dev())
    b = torch.randn(64, generator=gen).to(dev())
    full = ops.conv2d(xh, ops.pack_conv_weight(w, dtype), 3, bias=b, impl=0 if dtype == "bf16" else 1)
    pick = torch.empty(2, 4, 4, 64, dtype=xh.dtype, device=dev())
    _lib.check(lib.ldm_downsample_pick(full.data_ptr(), 64, pick.data_ptr(), 64, 2, 8, 8, 64, ops._dt(xh), _lib.stream_ptr()))
    wr = w.to(torch.bfloat16).float() if dtype == "bf16" else w
    want = F.conv2d(F.pad(xr, (0, 1, 0, 1)), wr, b, stride=2)
    assert rel_l2(ops.to_nchw(pick), want) < (1e-5 if dtype == "fp32" else 6e-3)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("N,C", [(256, 128), (64, 256), (16, 512), (1024, 64), (9, 64)])
def test_attention_single_head(dtype, N, C):
    from ldm_b200 import ops, _lib
    gen = torch.Generator().manual_seed(N + C)
    B = 2
    qkv = torch.randn(B, N, 3 * C, generator=gen).to(dev())
    qkv = qkv.to(torch.bfloat16) if dtype == "bf16" else qkv
    out = torch.empty(B, N, C, dtype=qkv.dtype, device=dev())
    _lib.check(_lib.load().ldm_attention_single_head(qkv.data_ptr(), out.data_ptr(), B, N, C, ops._dt(qkv), _lib.stream_ptr()))
    q, k, v = qkv.float().split(C, dim=2)
    attn = torch.softmax(torch.einsum("bic,bjc->bij", q, k) * C ** -0.5, dim=2)
    want = torch.einsum("bij,bjc->bic", attn, v)
    assert rel_l2(out.float(), want) < (1e-5 if dtype == "fp32" else 6e-3)


@pytest.mark.parametrize("dtype,tol_eps,tol_x", [("fp32", 1e-4, 2e-4), ("bf16", 2e-2, 2e-2)])
def test_latent_diffusion_model_matches_reference_golden(dtype, tol_eps, tol_x):
    """BASELINE config 4: encode -> eps-prediction / reverse steps on the latent with the model's sqrt-linear schedule ->
    decode, each stage teacher-forced with the reference's tensors and once chained end to end."""
    import ldm_b200
    import oracle
    g = golden("g11_ldm_latent.npz")
    unet = ldm_b200.UNet(4, 4, 64, (1, 2, 4, 8), True, 10, dtype=dtype).to(dev())
    unet.load_state_dict(oracle.init_state_dict(int(g["unet_seed"]), 4, 4, 64, (1, 2, 4, 8), True, 10))
    torch.manual_seed(int(g["ae_seed"]))
    ae = ldm_b200.Autoencoder(3, 4, 3, 64, [1, 2], 2, dtype=dtype).to(dev())
    ldm = ldm_b200.LatentDiffusionModel(unet, ae, 0.18215, 1000, 0.00085, 0.012).to(dev())
    assert torch.equal(ldm.beta.data.cpu(), T(g["beta"])) and torch.equal(ldm.alpha_bar.data.cpu(), T(g["alpha_bar"]))
    img, y = T(g["img"]).to(dev()), T(g["y"]).to(dev())
    z0 = ldm.autoencoder_encode(img, epsilon=T(g["encode_epsilon"]).to(dev()))
    assert rel_l2(z0.cpu(), T(g["z0"])) < tol_eps
    z0_ref = T(g["z0"]).to(dev())
    with torch.no_grad():
        assert rel_l2(ldm(z0_ref, T(g["t"]).to(dev()), y).cpu(), T(g["eps_pred"])) < tol_eps
    d = ldm.make_diffusion(dev())
    noise = torch.zeros(1000, *z0_ref.shape, device=dev())
    for i, step in enumerate((999, 998, 997)):
        noise[step] = T(g["step_noise"][i]).to(dev())
    x = d.sample(ldm, y, tuple(z0_ref.shape), dev(), cfg_scale=3, x_T=z0_ref, noise=noise, first_step=999, num_steps=3,
                 return_device=True)
    assert rel_l2(x.cpu(), T(g["x_after3"])) < tol_x
    dec = ldm.autoencoder_decode(T(g["x_after3"]).to(dev()))
    assert rel_l2(dec.cpu(), T(g["decoded"])) < tol_eps
    chained = ldm.autoencoder_decode(x)                      # our latent through our decoder
    assert rel_l2(chained.cpu(), T(g["decoded"])) < 3 * tol_x
