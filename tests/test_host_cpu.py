"""CPU: host-side logic -- the C-ABI library loads and exports every symbol include/ldm_b200.h declares,
the drop-in classes keep the reference's surface, errors are loud (no fallback), and the multi-rank helpers
work under a world_size-2 gloo group."""
import importlib.util
import os
import re
import sys

import pytest
import torch
import torch.multiprocessing as mp

from conftest import REFERENCE, ROOT


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "ldm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(ldm_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    from ldm_b200 import _lib
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.ldm_abi_version() == 1


def test_no_cpu_fallback():
    import ldm_b200
    from ldm_b200 import _lib
    m = ldm_b200.UNet(3, 3, 64, [1, 2], True, 10)
    with pytest.raises(_lib.LdmError):
        with torch.no_grad():
            m(torch.zeros(1, 3, 16, 16), torch.zeros(1, dtype=torch.long))
    d = ldm_b200.Diffusion(10, "cpu")
    with pytest.raises(_lib.LdmError):
        d.sample(m, torch.tensor([1]), (1, 3, 16, 16), "cpu")
    with pytest.raises(_lib.LdmError):
        d.p_sample(torch.zeros(1, 3, 16, 16), torch.zeros(1, dtype=torch.long), torch.zeros(1, 3, 16, 16))


def test_error_strings_cross_the_abi(lib):
    import ctypes as C
    from ldm_b200 import _lib
    d = _lib.UNetDesc()
    d.in_channels = d.out_channels = 3
    d.channels = 48          # not a multiple of 64
    d.n_levels = 4
    d.image_size = 32
    out = C.c_void_p()
    assert lib.ldm_unet_create(C.byref(d), C.byref(out)) != 0
    assert b"multiple of 64" in lib.ldm_last_error()
    d.channels = 64
    d.image_size = 28        # the reference's 4-level UNet cannot take 28x28
    for i in range(4):
        d.channel_multipliers[i] = 1
    assert lib.ldm_unet_create(C.byref(d), C.byref(out)) != 0
    assert b"not divisible" in lib.ldm_last_error()


def test_dropin_surface_matches_reference():
    import ldm_b200
    import oracle
    torch.manual_seed(0)
    m = ldm_b200.UNet(3, 3, 64, [1, 2, 4, 8], True, 10)
    sd = oracle.init_state_dict(0, 3, 3, 64, (1, 2, 4, 8), True, 10)
    assert list(m.state_dict().keys()) == list(sd.keys())
    assert all(torch.equal(m.state_dict()[k], sd[k]) for k in sd)
    assert m.num_classes == 10 and len(list(m.buffers())) == 0
    d = ldm_b200.Diffusion(1000, "cpu", n_samples=100)
    s = oracle.make_schedule(1000)
    for k in ("beta", "alpha", "alpha_bar", "sigma2"):
        assert torch.equal(getattr(d, k), s[k])
    assert d.n_steps == 1000 and d.n_samples == 100 and len(d.state_dict()) == 0   # schedule is not in state_dict
    mean, var = d.q_xt_x0(torch.ones(2, 1, 2, 2), torch.tensor([0, 999]))
    assert mean.shape == (2, 1, 2, 2) and var.shape == (2, 1, 1, 1)
    ldm = ldm_b200.LatentDiffusionModel(m, torch.nn.Identity(), 0.18215, 1000, 0.00085, 0.012)
    ref = oracle.ddpm_oracle.make_ldm_schedule(1000, 0.00085, 0.012)
    assert torch.equal(ldm.beta.data, ref["beta"]) and torch.equal(ldm.alpha_bar.data, ref["alpha_bar"])
    keys = list(ldm.state_dict().keys())
    assert "beta" in keys and "alpha_bar" in keys and "model.diffusion_model.initial_conv.weight" in keys


def test_yaml_target_resolves_to_native_classes():
    """config_files/*.yaml say `target: src.UNet.UNet` / `src.DDPM.Diffusion`; with our package first on sys.path
    the reference's own factory (src/utils.py:48-88 semantics) instantiates the native classes."""
    import importlib
    import ldm_b200
    pkg = os.path.join(ROOT, "latent-diffusion-models_b200")
    saved = [k for k in sys.modules if k == "src" or k.startswith("src.")]
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, pkg)
    try:
        for target, want in (("src.UNet.UNet", ldm_b200.UNet), ("src.DDPM.Diffusion", ldm_b200.Diffusion)):
            mod, cls = target.rsplit(".", 1)
            assert getattr(importlib.import_module(mod), cls) is want
        cfg = {"target": "src.UNet.UNet", "params": dict(in_channels=3, out_channels=3, channels=64,
               channel_multipliers=[1, 2, 4, 8], with_time_emb=True, num_classes=10)}
        mod, cls = cfg["target"].rsplit(".", 1)
        assert isinstance(getattr(importlib.import_module(mod), cls)(**cfg["params"]), ldm_b200.UNet)
    finally:
        sys.path.remove(pkg)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from ldm_b200 import dist as D
    r, lr, w = D.init_from_env("gloo")
    off = D.shard_offset(256, r)
    so, sc = D.split_batch(10, r, w)
    t = D.max_over_ranks(1.0 + r)
    total = D.sum_over_ranks(256.0)
    p = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2))]
    p[0].grad = torch.full((3,), float(r + 1))        # p[1] has no grad: zero-filled, all ranks reduce equal lengths
    flat = D.allreduce_mean_(D.flatten_grads(p))
    D.unflatten_grads(p, flat)
    D.barrier()
    q.put((r, off, so, sc, t, total, flat.tolist(), p[0].grad.tolist(), p[1].grad is None))
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_sharding_and_gradient_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [0, 256]
    assert [(r[2], r[3]) for r in res] == [(0, 5), (5, 5)]
    assert all(r[4] == 2.0 and r[5] == 512.0 for r in res)
    assert all(r[6] == [1.5, 1.5, 1.5, 0.0, 0.0] and r[7] == [1.5, 1.5, 1.5] and r[8] for r in res)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with the contract's keys."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-batch", "1", "--cpu-timesteps", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cifar10_ddpm_sampling_images_per_sec" and d["unit"] == "images/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data"):
        assert k in d, k
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference is only mounted in the build container")
def test_unmodified_generate_images_drives_the_native_classes(tmp_path, monkeypatch):
    """The reference's own caller (generate_images.py:18-41, unmodified, imported from /root/reference) and its own factory
    (src/utils.py:48-88) and YAML run over the `src.UNet` / `src.DDPM` shadow: every class the config names resolves to the
    native implementation and `Diffusion.sample` is driven with the reference's call protocol -- ten batch-1 calls,
    classes = tensor([i]), shape (1, C, S, S), cfg_scale 3 -- and the PNGs land where the reference writes them.
    No GPU here: the CUDA entry point is replaced by a recorder (the numerics of that call are the -m gpu suites' job)."""
    import importlib
    import yaml
    import ldm_b200
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.") or k == "generate_images"}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE)
    try:
        importlib.import_module("src")                                   # the reference's package ...
        shadow = os.path.join(ROOT, "latent-diffusion-models_b200", "src")
        for name in ("UNet", "DDPM"):                                    # ... with the two hot-path modules shadowed
            spec = importlib.util.spec_from_file_location(f"src.{name}", os.path.join(shadow, f"{name}.py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[f"src.{name}"] = mod
            spec.loader.exec_module(mod)
        gen = importlib.import_module("generate_images")                 # the unmodified caller
        from src.utils import get_model_from_config                      # the unmodified factory
        cfg = yaml.safe_load(open(os.path.join(REFERENCE, "config_files", "pixel_diffusion_model_cifar10.yaml")))
        cfg["diffusion"]["params"]["device"] = "cpu"                     # the YAML says cuda (SURVEY 8c shim)
        diffusion = get_model_from_config(cfg["diffusion"])
        model = get_model_from_config(cfg["model"])
        assert type(model) is ldm_b200.UNet and type(diffusion) is ldm_b200.Diffusion
        assert diffusion.n_steps == cfg["diffusion"]["params"]["n_steps"]
        calls = []

        def recorder(self, eps_model, classes, shape, device, cfg_scale=3, **kw):
            calls.append((eps_model, classes.clone(), tuple(shape), str(device), cfg_scale, kw))
            return torch.zeros(shape)                                    # a CPU tensor, as src/DDPM.py:128 returns

        monkeypatch.setattr(ldm_b200.Diffusion, "sample", recorder)
        monkeypatch.chdir(tmp_path)
        gen.main(model, diffusion, cfg, "cpu")
        assert len(calls) == model.num_classes == 10
        for i, (m, classes, shape, device, scale, kw) in enumerate(calls):
            assert m is model and classes.tolist() == [i] and classes.dtype == torch.int64
            assert shape == (1, cfg["data"]["image_channels"], cfg["data"]["image_size"], cfg["data"]["image_size"])
            assert scale == 3 and not kw
        folder = tmp_path / cfg["diffusion"]["type"] / cfg["project_name"] / "results"
        assert sorted(p.name for p in folder.iterdir()) == [str(i) for i in range(10)]
        assert all(any(f.suffix == ".png" for f in (folder / str(i)).iterdir()) for i in range(10))
    finally:
        sys.path.remove(REFERENCE)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.") or k == "generate_images"]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_native_state_is_not_shared_by_copies():
    """copy.deepcopy / pickle of a UNet or Diffusion (EMA copies, best-model snapshots, torch.save(model)) must not carry
    native handles: two Python owners of one ldm_unet* repack each other's weights and double-free."""
    import copy
    import pickle
    import ldm_b200
    m = ldm_b200.UNet(3, 3, 64, [1, 2], True, 10)
    m._handles[(32, 0)] = 0xDEAD                      # as if a forward had created a native handle
    m._loaded[0xDEAD] = ("fingerprint",)
    for clone in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert clone._handles == {} and clone._loaded == {} and clone._ws == {} and clone._uid != m._uid
        assert all(torch.equal(a, b) for a, b in zip(clone.state_dict().values(), m.state_dict().values()))
    m._handles.clear()                                # nothing real to destroy
    d = ldm_b200.Diffusion(10, "cpu")
    d._samplers[("key",)] = {"s": 0xBEEF, "unet": lambda: None}
    c = copy.deepcopy(d)
    assert c._samplers == {} and torch.equal(c.beta, d.beta)
    d._samplers.clear()
