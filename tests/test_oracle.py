"""CPU: the oracle restatement against the golden vectors produced by the unmodified reference
(oracle/make_golden.py), and against the live reference when /root/reference is mounted."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import REFERENCE, golden, rel_l2
import oracle
from oracle import unet_oracle as U
from oracle import ddpm_oracle as D

T = torch.from_numpy


def _sha(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.digest()


@pytest.mark.parametrize("tag,cin", [("cifar", 3), ("mnist", 1)])
def test_unet_matches_reference_golden(tag, cin):
    g = golden(f"g1_unet_{tag}.npz")
    sd = oracle.init_state_dict(int(g["weight_seed"]), cin, cin, 64, (1, 2, 4, 8), True, 10)
    assert len(sd) == 200
    assert _sha(sd) == bytes(g["weight_sha256"]), "default init no longer reproduces the reference's weights"
    x, t, y = T(g["x"]), T(g["t"]), T(g["y"])
    with torch.no_grad():
        assert rel_l2(U.unet_forward(sd, x, t, y), T(g["eps_cond"])) < 2e-6
        assert rel_l2(U.unet_forward(sd, x, t, None), T(g["eps_uncond"])) < 2e-6
        assert rel_l2(U.unet_forward(sd, x, t, torch.tensor([3])), T(g["eps_bcast3"])) < 2e-6


def test_unet_fp64_vs_fp32_floor():
    g = golden("g1_unet_cifar.npz")
    sd = oracle.init_state_dict(0, 3, 3, 64, (1, 2, 4, 8), True, 10)
    sd64 = U.cast_state_dict(sd, torch.float64)
    with torch.no_grad():
        out = U.unet_forward(sd64, T(g["x"]).double(), T(g["t"]), T(g["y"]))
    assert rel_l2(out, T(g["eps_cond"])) < 5e-6  # fp32 vs fp64 floor 2.4e-7


@pytest.mark.parametrize("name,fn,prefix", [
    ("linattn_64_n1024", U.linear_attention, ""),
    ("linattn_128_n16", U.linear_attention, ""),
    ("attn_512_n4", U.attention, ""),
    ("attn_64_n64", U.attention, ""),
])
def test_attention_submodules(name, fn, prefix):
    g = golden(f"g4_{name}.npz")
    sd = {"p." + k[3:]: T(g[k]) for k in g.files if k.startswith("w::")}
    with torch.no_grad():
        out = fn(sd, "p", T(g["in0"]))
    assert rel_l2(out, T(g["out"])) < 2e-6


def test_block_and_resblocks():
    g = golden("g4_block_64_64_r8.npz")
    sd = {"p." + k[3:]: T(g[k]) for k in g.files if k.startswith("w::")}
    with torch.no_grad():
        assert rel_l2(U.block(sd, "p", T(g["in0"])), T(g["out"])) < 2e-6
    g = golden("g4_resblock_64_128_t_r8.npz")
    sd = {"p." + k[3:]: T(g[k]) for k in g.files if k.startswith("w::")}
    with torch.no_grad():
        assert rel_l2(U.resnet_block(sd, "p", T(g["in0"]), T(g["in1"])), T(g["out"])) < 2e-6
    g = golden("g4_resblock_64_64_not_r16.npz")
    sd = {"p." + k[3:]: T(g[k]) for k in g.files if k.startswith("w::")}
    with torch.no_grad():
        assert rel_l2(U.resnet_block(sd, "p", T(g["in0"]), None), T(g["out"])) < 2e-6


def test_time_embedding():
    g = golden("g4_time_embedding_256.npz")
    sd = {"time_emb." + k[3:]: T(g[k]) for k in g.files if k.startswith("w::")}
    with torch.no_grad():
        out = U.time_embedding(sd, T(g["in0"]), torch.float32)
    assert rel_l2(out, T(g["out"])) < 2e-6


@pytest.mark.parametrize("n_steps", [1000, 400])
def test_schedule_and_p_sample(n_steps):
    g = golden(f"g2_p_sample_T{n_steps}.npz")
    s = D.make_schedule(n_steps)
    assert torch.equal(s["beta"], T(g["beta"])) and torch.equal(s["alpha_bar"], T(g["alpha_bar"]))
    xt, eps = T(g["xt"]), T(g["eps"])
    for i, step in enumerate(g["steps"]):
        t = torch.full((4,), int(step), dtype=torch.long)
        out = D.p_sample(s, xt, t, eps, T(g["noise"][i]))
        assert torch.allclose(out, T(g["out"][i]), rtol=1e-6, atol=1e-6), f"t={step}"
    if n_steps == 1000:
        assert abs(float(s["alpha_bar"][-1]) - 4.036e-5) < 1e-7


def test_q_sample_and_forward():
    g = golden("g3_q_sample.npz")
    s = D.make_schedule(1000)
    xt = D.q_sample(s, T(g["x0"]), T(g["t"]), T(g["noise"]))
    assert torch.allclose(xt, T(g["xt"]), rtol=1e-6, atol=1e-6)


def test_trajectory_statistics_short():
    """Teacher-forced check of the oracle loop on the reference's trajectory checkpoints (3 steps only:
    a full 1000-step CPU trajectory is minutes; the GPU test runs all of it)."""
    g = golden("g6_trajectory_T1000.npz")
    sd = oracle.init_state_dict(int(g["weight_seed"]), 3, 3, 64, (1, 2, 4, 8), True, 10)
    s = D.make_schedule(1000)
    torch.manual_seed(int(g["noise_seed"]))
    x_T = torch.randn(2, 3, 32, 32)
    assert torch.equal(x_T, T(g["x_at_999"])), "x_T draw no longer matches the reference's RNG order"
    model = lambda x, t, y: U.unet_forward(sd, x, t, y)
    x = x_T
    with torch.no_grad():
        for step in (999, 998):
            t = torch.full((2,), step, dtype=torch.long)
            eps = D.cfg_combine(model(x, t, torch.tensor([3])), model(x, t, None), 3.0)
            z = torch.randn(2, 3, 32, 32)
            x = D.p_sample(s, x, t, eps, z)
            st = g["stats"][step - 1]
            assert abs(float(x.mean()) - st[0]) < 1e-4 * max(1.0, abs(st[1]))
            assert abs(float(x.std()) - st[1]) < 1e-4 * st[1]


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not mounted")
def test_against_live_reference():
    sys.path.insert(0, REFERENCE)
    try:
        from src.UNet import UNet as RefUNet
    finally:
        sys.path.remove(REFERENCE)
    torch.manual_seed(11)
    ref = RefUNet(3, 3, 64, [1, 2], True, 10).eval()   # 2-level variant, cheap
    sd = {k: v.detach() for k, v in ref.state_dict().items()}
    mine = oracle.init_state_dict(11, 3, 3, 64, (1, 2), True, 10)
    assert all(torch.equal(mine[k], sd[k]) for k in sd)
    x = torch.randn(2, 3, 16, 16)
    t = torch.tensor([5, 700])
    y = torch.tensor([1, 9])
    with torch.no_grad():
        assert rel_l2(U.unet_forward(sd, x, t, y), ref(x, t, y)) < 2e-6
    for m in [k for k in list(sys.modules) if k == "src" or k.startswith("src.")]:
        del sys.modules[m]


# ---------------------------------------------------------------- steps either side of the hot path
def test_adam_matches_torch_optim_golden():
    from oracle import trainer_oracle as TO
    g = golden("g7_adam.npz")
    for j in range(3):
        p = T(g[f"p0_{j}"]).clone()
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        for s in range(int(g["steps"])):
            TO.adam_step(p, T(g[f"grads_{j}"][s]), m, v, s + 1, float(g["lr"]))
        assert rel_l2(p, T(g[f"p_final_{j}"])) < 1e-7
        assert rel_l2(m, T(g[f"exp_avg_{j}"])) < 1e-6
        assert rel_l2(v, T(g[f"exp_avg_sq_{j}"])) < 1e-6


def test_output_stage_matches_reference_golden():
    from oracle import trainer_oracle as TO
    g = golden("g8_output.npz")
    assert np.array_equal(TO.images_to_uint8(g["x"], "reverse_transform"), g["reverse_transform"])
    assert np.array_equal(TO.images_to_uint8(g["x"], "save_image"), g["save_image"])
    assert np.array_equal(TO.images_to_uint8(g["x_gray"], "reverse_transform")[..., 0], g["reverse_transform_gray"])


def test_val_loss_matches_reference_golden():
    from oracle import trainer_oracle as TO
    g = golden("g9_val_loss.npz")
    sd = oracle.init_state_dict(int(g["weight_seed"]), 3, 3, 64, (1, 2, 4, 8), True, 10)
    sched = D.make_schedule(1000)
    with torch.no_grad():
        l3 = TO.val_loss(sd, sched, T(g["x0"]), T(g["noise"]), T(g["t"]), T(g["y"]), 3.0)
        l0 = TO.val_loss(sd, sched, T(g["x0"]), T(g["noise"]), T(g["t"]), T(g["y"]), 0.0)
    assert abs(float(l3) - float(g["loss_cfg3"])) < 1e-5 * float(g["loss_cfg3"])
    assert abs(float(l0) - float(g["loss_cfg0"])) < 1e-5 * float(g["loss_cfg0"])


# ---------------------------------------------------------------- first-stage autoencoder and the latent model
def _ae_from_golden(g, **kw):
    import ldm_b200
    cfg = [int(v) for v in g["config"]]
    torch.manual_seed(int(g["weight_seed"]))
    ae = ldm_b200.Autoencoder(cfg[0], cfg[1], cfg[2], cfg[3], cfg[5:], cfg[4], **kw)
    return ae


@pytest.mark.parametrize("tag", ["ldm", "deep"])
def test_autoencoder_oracle_matches_reference_golden(tag):
    from oracle import autoencoder_oracle as A
    g = golden(f"g10_autoencoder_{tag}.npz")
    ae = _ae_from_golden(g)                       # parameter holders only: built on the CPU, never called
    sd = {k: v.detach() for k, v in ae.state_dict().items()}
    assert _sha(sd) == bytes(g["weight_sha256"]), "holder construction order no longer reproduces the reference's init"
    img = T(g["img"])
    with torch.no_grad():
        mu, lv = A.encode_moments(sd, img)
        z = A.gaussian_sample(mu, lv, T(g["epsilon"]))
        rec = A.decode(sd, z)
    assert rel_l2(mu, T(g["mu"])) < 2e-6 and rel_l2(lv, T(g["log_var"])) < 2e-6
    assert rel_l2(z, T(g["z"])) < 2e-6
    assert rel_l2(rec, T(g["recon"])) < 5e-6
    assert rel_l2(rec, T(g["forward_img"])) < 5e-6


def test_latent_diffusion_oracle_matches_reference_golden():
    from oracle import autoencoder_oracle as A
    import ldm_b200
    g = golden("g11_ldm_latent.npz")
    sched = D.make_ldm_schedule(1000, 0.00085, 0.012)
    assert torch.equal(sched["beta"], T(g["beta"])) and torch.equal(sched["alpha_bar"], T(g["alpha_bar"]))
    sched = dict(sched, alpha=1.0 - sched["beta"], sigma2=sched["beta"])   # what a Diffusion built on this schedule derives
    torch.manual_seed(int(g["ae_seed"]))
    ae_sd = {k: v.detach() for k, v in ldm_b200.Autoencoder(3, 4, 3, 64, [1, 2], 2).state_dict().items()}
    unet_sd = oracle.init_state_dict(int(g["unet_seed"]), 4, 4, 64, (1, 2, 4, 8), True, 10)
    y = T(g["y"])
    with torch.no_grad():
        mu, lv = A.encode_moments(ae_sd, T(g["img"]))
        z0 = 0.18215 * A.gaussian_sample(mu, lv, T(g["encode_epsilon"]))
        assert rel_l2(z0, T(g["z0"])) < 2e-6
        assert rel_l2(U.unet_forward(unet_sd, z0, T(g["t"]), y), T(g["eps_pred"])) < 5e-6
        x = z0
        for i, step in enumerate((999, 998, 997)):
            tt = torch.full((2,), step, dtype=torch.long)
            eps = D.cfg_combine(U.unet_forward(unet_sd, x, tt, y), U.unet_forward(unet_sd, x, tt, None), 3.0)
            x = D.p_sample(sched, x, tt, eps, T(g["step_noise"][i]))
        assert rel_l2(x, T(g["x_after3"])) < 1e-5
        assert rel_l2(A.decode(ae_sd, x / 0.18215), T(g["decoded"])) < 1e-5
