"""GPU: the graph sampler (Diffusion.sample) against the oracle loop, the reference's own trajectory
(golden G6), and its invariances (graph == eager launches, independent of batch sharding)."""
import numpy as np
import pytest
import torch

from conftest import golden, rel_l2
from test_unet_gpu import make_model, dev

pytestmark = pytest.mark.gpu
T = torch.from_numpy


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("cfg", [3.0, 0.0])
def test_short_trajectory_vs_oracle(dtype, tol, cfg):
    """12-step schedule, injected x_T and noise, classes broadcast from length 1 (main.py:318)."""
    import ldm_b200
    from oracle import unet_oracle as U, ddpm_oracle as D
    m, sd = make_model(dtype)
    n = 12
    g = torch.Generator().manual_seed(9)
    x_T = torch.randn(3, 3, 32, 32, generator=g)
    noise = torch.randn(n, 3, 3, 32, 32, generator=g)
    cls = torch.tensor([3])
    sched = D.make_schedule(n)
    with torch.no_grad():
        want = D.sample_loop(sched, lambda x, t, y: U.unet_forward(sd, x, t, y), cls, x_T, noise, cfg_scale=cfg)
    d = ldm_b200.Diffusion(n, dev())
    got = d.sample(m, cls, (3, 3, 32, 32), dev(), cfg_scale=cfg, x_T=x_T, noise=noise)
    assert got.device.type == "cpu" and got.dtype == torch.float32      # reference returns xt.detach().cpu()
    assert rel_l2(got, want) < tol * 3   # 12 accumulated steps
    # per-class labels of length B (DiffusionModelTrainer.sample, src/DiffusionModelTrainer.py:161-174)
    cls_b = torch.tensor([1, 5, 9])
    with torch.no_grad():
        want_b = D.sample_loop(sched, lambda x, t, y: U.unet_forward(sd, x, t, y), cls_b, x_T, noise, cfg_scale=cfg)
    got_b = d.sample(m, cls_b, (3, 3, 32, 32), dev(), cfg_scale=cfg, x_T=x_T, noise=noise)
    assert rel_l2(got_b, want_b) < tol * 3


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_graph_equals_eager_and_sharding_invariance(dtype):
    import ldm_b200
    m, _ = make_model(dtype)
    d = ldm_b200.Diffusion(10, dev())
    cls = torch.tensor([7])
    a = d.sample(m, cls, (4, 3, 32, 32), dev(), cfg_scale=3, seed=1234, use_graph=True)
    b = d.sample(m, cls, (4, 3, 32, 32), dev(), cfg_scale=3, seed=1234, use_graph=False)
    assert torch.equal(a, b), "graph replay must be bit-identical to eager launches"
    a2 = d.sample(m, cls, (4, 3, 32, 32), dev(), cfg_scale=3, seed=1234, use_graph=True)
    assert torch.equal(a, a2), "replaying the cached graph must be deterministic"
    lo = d.sample(m, cls, (2, 3, 32, 32), dev(), cfg_scale=3, seed=1234, sample_offset=0)
    hi = d.sample(m, cls, (2, 3, 32, 32), dev(), cfg_scale=3, seed=1234, sample_offset=2)
    assert torch.equal(torch.cat([lo, hi]), a), "a 2+2 sharded batch must reproduce the 4-image batch bit for bit"
    c = d.sample(m, cls, (4, 3, 32, 32), dev(), cfg_scale=3, seed=99)
    assert not torch.equal(a, c)
    assert torch.isfinite(a).all()
    assert d.last_launches > 10 * 50


def test_large_batch_is_invariant_to_batch_size():
    """BASELINE config 5 (generation sweep, batch 64-8192 per GPU): a sample's trajectory depends on its global index
    only, so the last 8 images of a 2048-image batch equal the same 8 global samples generated on their own."""
    import ldm_b200
    m, _ = make_model("bf16")
    d = ldm_b200.Diffusion(1000, dev())
    cls = torch.tensor([3])
    B = 2048
    big = d.sample(m, cls, (B, 3, 32, 32), dev(), cfg_scale=3, seed=5, first_step=999, num_steps=3, return_device=True)
    head = d.sample(m, cls, (8, 3, 32, 32), dev(), cfg_scale=3, seed=5, first_step=999, num_steps=3, return_device=True)
    tail = d.sample(m, cls, (8, 3, 32, 32), dev(), cfg_scale=3, seed=5, sample_offset=B - 8, first_step=999, num_steps=3,
                    return_device=True)
    assert torch.isfinite(big).all()
    assert torch.equal(big[:8], head) and torch.equal(big[-8:], tail)


def _reference_noise(g, T_, shape):
    torch.manual_seed(int(g["noise_seed"]))
    x_T = torch.randn(shape)
    noise = torch.zeros((T_,) + shape)
    for step in reversed(range(1, T_)):
        noise[step] = torch.randn(shape)
    return x_T, noise


@pytest.mark.parametrize("dtype,std_tol,mean_tol,x_tol", [("fp32", 1e-3, 1e-3, 2e-3), ("bf16", 5e-3, 1e-2, 2e-2)])
def test_full_trajectory_vs_reference(dtype, std_tol, mean_tol, x_tol):
    """The reference's own 1000-step, cfg-3 trajectory (golden G6): checkpoints and per-step drift statistics.
    Gates follow SURVEY.md 8c (reference bf16-autocast vs fp32: x_t <= 3.6e-3, |dstd|/std <= 6.5e-4, |dmean|/std <= 1.7e-3)."""
    import ldm_b200
    g = golden("g6_trajectory_T1000.npz")
    Tn, shape = 1000, (2, 3, 32, 32)
    m, _ = make_model(dtype, seed=int(g["weight_seed"]))
    x_T, noise = _reference_noise(g, Tn, shape)
    assert torch.equal(x_T, T(g["x_at_999"]))
    d = ldm_b200.Diffusion(Tn, dev())
    cls = torch.tensor([3])
    noise_d = noise.to(dev())
    x = x_T.to(dev())
    stats = g["stats"]
    worst_std = worst_mean = 0.0
    for step in reversed(range(Tn)):
        # x is the input of timestep `step`
        if step % 10 == 0 or step > 990:
            mean, std = float(x.mean()), float(x.std())
            worst_std = max(worst_std, abs(std - stats[step][1]) / stats[step][1])
            worst_mean = max(worst_mean, abs(mean - stats[step][0]) / stats[step][1])
        if f"x_at_{step}" in g.files:
            assert rel_l2(x, T(g[f"x_at_{step}"])) < x_tol, f"x_t at step {step}"
        x = d.sample(m, cls, shape, dev(), cfg_scale=3, x_T=x, noise=noise_d, first_step=step, num_steps=1,
                     return_device=True, seed=0)
    assert rel_l2(x, T(g["out"])) < x_tol
    assert worst_std < std_tol and worst_mean < mean_tol, (worst_std, worst_mean)


@pytest.mark.parametrize("dtype,x_tol", [("fp32", 2e-3), ("bf16", 2e-2)])
def test_free_running_sample_call_vs_reference(dtype, x_tol):
    """ONE Diffusion.sample() call over all 1000 timesteps (the graph-replayed loop, not 1000 single-step calls) with the
    reference's x_T and per-step noise lands on the reference's final image (golden G6 `out`)."""
    import ldm_b200
    g = golden("g6_trajectory_T1000.npz")
    Tn, shape = 1000, (2, 3, 32, 32)
    m, _ = make_model(dtype, seed=int(g["weight_seed"]))
    x_T, noise = _reference_noise(g, Tn, shape)
    d = ldm_b200.Diffusion(Tn, dev())
    out = d.sample(m, torch.tensor([3]), shape, dev(), cfg_scale=3, x_T=x_T, noise=noise.to(dev()), seed=0)
    assert not out.is_cuda and out.shape == shape            # the reference returns a CPU tensor (src/DDPM.py:128)
    assert rel_l2(out, T(g["out"])) < x_tol
