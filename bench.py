#!/usr/bin/env python
"""Headline benchmark: CIFAR-10-shaped DDPM sampling throughput (images/sec), 1000-step reverse process with
classifier-free guidance, batch 256 per GPU (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full pass of the hot path over one batch: ``Diffusion.sample(model, classes, (B,3,32,32), device,
cfg_scale=3)`` = T x {one 2B-row UNet eps-prediction (cond + uncond), fused CFG + p_sample update}.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  value        images/sec, inputs resident in HBM (x_T drawn on the device, result left on the device)
  e2e          the same call through the public API with HOST inputs: x_T from pinned host memory, host labels,
               result copied back to the host (the reference's ``xt.detach().cpu()``), inside the timed region
  roofline     dominant kernel family (tcgen05 implicit-GEMM convolutions): algorithmic FLOPs / CUDA-event time of
               every conv launch of one 2B-row UNet pass, against MEASURED_PEAKS.json
  cpu_baseline the CPU oracle (a restatement of the reference's PyTorch code, oracle/) on the host cores, bounded sample
  clocks       nvidia-smi samples taken during the timed region

``--impl reference`` times the reference algorithm on the host CPU (the oracle port; the reference itself is a
Python checkout that does not exist on the GPU box) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNET_GFLOP_PER_IMAGE = 1.5132  # one UNet forward, CIFAR config (2xMAC, FlopCounterMode on the reference)
METRIC = "cifar10_ddpm_sampling_images_per_sec"
UNIT = "images/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_sampling_rate(batch: int, timesteps: int, n_steps: int, cfg_scale: float, warmup: int = 1):
    """images/sec of the CPU oracle: `timesteps` reverse steps at `batch` images are timed and extrapolated to the
    full n_steps-step trajectory (every step costs the same: 2 UNet passes + the p_sample update)."""
    import torch
    import oracle
    from oracle import unet_oracle as U, ddpm_oracle as D
    torch.manual_seed(42)
    sd = oracle.init_state_dict(42, 3, 3, 64, (1, 2, 4, 8), True, 10)
    sched = D.make_schedule(n_steps)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(batch, 3, 32, 32, generator=g)
    cls = torch.tensor([3])

    def one_step(x, step):
        t = torch.full((batch,), step, dtype=torch.long)
        eps = U.unet_forward(sd, x, t, cls)
        if cfg_scale > 0:
            eps = D.cfg_combine(eps, U.unet_forward(sd, x, t, None), cfg_scale)
        return D.p_sample(sched, x, t, eps, torch.randn(x.shape, generator=g))

    with torch.no_grad():
        for i in range(warmup):
            x = one_step(x, n_steps - 1 - i)
        t0 = time.perf_counter()
        for i in range(timesteps):
            x = one_step(x, n_steps - 1 - warmup - i)
        dt = time.perf_counter() - t0
    per_step = dt / timesteps
    return batch / (per_step * n_steps), per_step, torch.get_num_threads()


def gpu_library_baseline(dev, batch, n_steps, cfg_scale, stream, timesteps=10):
    """images/sec of the reference algorithm in eager PyTorch on the GPU: fp32 with cuDNN's TF32 convolutions (torch's
    default) and bf16 autocast.  `timesteps` reverse steps are timed after 2 warm-up steps and extrapolated x T."""
    import torch
    import oracle
    from oracle import unet_oracle as U, ddpm_oracle as D
    out = {"sample": f"{timesteps} of {n_steps} timesteps at batch {batch} (2 UNet passes + p_sample each), extrapolated; eager "
                     f"PyTorch {torch.__version__}, cuDNN {torch.backends.cudnn.version()}"}
    try:
        sd = {k: v.to(dev) for k, v in oracle.init_state_dict(42, 3, 3, 64, (1, 2, 4, 8), True, 10).items()}
        sched = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in D.make_schedule(n_steps).items()}
        cls = torch.tensor([3], device=dev)

        def one_step(x, step):
            t = torch.full((batch,), step, dtype=torch.long, device=dev)
            eps = U.unet_forward(sd, x, t, cls)
            if cfg_scale > 0:
                eps = D.cfg_combine(eps, U.unet_forward(sd, x, t, None), cfg_scale)
            return D.p_sample(sched, x.float(), t, eps.float(), torch.randn(x.shape, device=dev))

        for name, ctx in (("fp32_tf32", torch.autocast("cuda", enabled=False)),
                          ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
            with torch.cuda.stream(stream), torch.no_grad(), ctx:
                x = torch.randn(batch, 3, 32, 32, device=dev)
                for i in range(2):
                    x = one_step(x, n_steps - 1 - i)
                torch.cuda.synchronize(dev)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for i in range(timesteps):
                    x = one_step(x, n_steps - 3 - i)
                b.record(stream)
                torch.cuda.synchronize(dev)
            per_step = a.elapsed_time(b) / timesteps
            out[name] = {"ms_per_timestep": per_step, "images_per_sec": batch / (per_step / 1e3 * n_steps)}
    except Exception as exc:   # noqa: BLE001 -- context only: never let the library baseline break the bench line
        out["error"] = f"{type(exc).__name__}: {exc}"
    return out


def run_reference(args):
    """The reference algorithm on the host CPU (oracle port), bounded sample; rank 0 only."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    b, ts = args.cpu_batch, args.cpu_timesteps
    vals = []
    for i in range(args.warmup + args.steps):
        v, per_step, cores = cpu_sampling_rate(b, ts, args.n_steps, args.cfg_scale, warmup=1 if i == 0 else 0)
        if i >= args.warmup:
            vals.append((v, per_step))
    v = statistics.mean(x[0] for x in vals)
    per_step = statistics.mean(x[1] for x in vals)
    sample = (f"{ts} of {args.n_steps} reverse timesteps at batch {b} (cfg_scale {args.cfg_scale}: 2 UNet passes + p_sample "
              f"per timestep), extrapolated x{args.n_steps}/{ts}; fp32, torch CPU kernels, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * b / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args, args.batch, 1), bounded_sample=sample),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, batch, world):
    return {"workload": f"pixel_diffusion_model_cifar10: Diffusion.sample, 3x32x32, T={args.n_steps}, cfg_scale={args.cfg_scale}, "
                        f"UNet(64,[1,2,4,8],num_classes=10), batch {batch} per GPU",
            "batch_per_gpu": batch, "global_batch": batch * world, "n_steps": args.n_steps, "cfg_scale": args.cfg_scale,
            "image": [3, 32, 32], "parallelism": f"batch-sharded x{world}, no collectives",
            "weights": "random init, torch.manual_seed(42)"}


def train_leg(args, dev, rank, world, stream, batch=None):
    """images/sec of DiffusionModelTrainer._train_epoch's body (src/DiffusionModelTrainer.py:36-67) on synthetic data."""
    import torch
    import ldm_b200
    from ldm_b200 import dist as ldist
    import torch.distributed as tdist
    B = batch or args.train_batch
    torch.manual_seed(42)
    model = ldm_b200.UNet(3, 3, 64, (1, 2, 4, 8), True, 10, dtype=args.dtype).to(dev)
    diffusion = ldm_b200.Diffusion(args.n_steps, dev)
    from ldm_b200 import trainer as ltrainer
    # Adam with the reference's settings (src/Trainer.py:68-71) over flat buffers: packs the gradients, all-reduces them
    # (N > 1) and updates all 20.35 M parameters in one launch.  Built before the graph capture: it re-homes p.data.
    opt = ltrainer.FlatAdam(model.parameters(), lr=5e-4)
    g = torch.Generator().manual_seed(rank)
    x0 = (torch.rand(B, 3, 32, 32, generator=g) * 2 - 1).pin_memory()
    y = torch.randint(0, 10, (B,), generator=g).pin_memory()
    fwd = model
    graphed = False
    torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)   # the leg runs on the bench's side stream
    if not args.train_eager:
        try:   # forward + backward of the UNet replayed as CUDA graphs (ldm_b200.train.make_graphed)
            from ldm_b200 import train as ltrain
            n0, xt0, t0 = diffusion(x0.to(dev))
            fwd = ltrain.make_graphed(model, xt0, t0, y.to(dev))
            graphed = True
        except Exception as exc:   # noqa: BLE001 -- the eager autograd path is always available
            sys.stderr.write(f"bench: graphed training unavailable ({exc}); using eager autograd\n")
            fwd = model

    def step():
        data, targets = x0.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
        return ltrainer.train_step(model, diffusion, opt, data, targets, forward=fwd)

    with torch.cuda.stream(stream):
        for _ in range(5):
            step()
        torch.cuda.synchronize(dev)
        if world > 1:
            tdist.barrier()
        n = 20
        windows = []
        for _ in range(3):   # three windows of n steps; the median window is reported
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n):
                loss = step()
            lv = float(loss.detach())  # the reference's loss.item() (:67), once per window
            e1.record(stream)
            torch.cuda.synchronize(dev)
            if world > 1:
                tdist.barrier()
            windows.append(ldist.max_over_ranks(e0.elapsed_time(e1), dev))
    ms = sorted(windows)[1]
    return {"metric": "cifar10_ddpm_train_images_per_sec", "value": B * world * n / (ms / 1e3), "unit": "images/s",
            "batch_per_gpu": B, "steps": n, "ms_per_step": ms / n, "ms_per_step_windows": [w / n for w in windows], "loss": lv,
            "cuda_graphs": graphed,
            "optimizer": "FlatAdam: backward kernels accumulate into the flat gradient bucket, ldm_adam_step per bucket piece",
            "note": "q_sample + UNet fwd + MSE + bwd (tcgen05 fwd / dgrad / wgrad) + bucket all-reduce (tail under the encoder backward) + Adam; "
                    "4.536 GFLOP/image"}


def run_ours(args):
    import torch
    import ldm_b200
    from ldm_b200 import _lib, dist as ldist
    import torch.distributed as tdist

    rank, local_rank, world = ldist.init_from_env("nccl")
    if world != args.gpus and rank == 0:
        sys.stderr.write(f"bench: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE\n")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B, T, cfg = args.batch, args.n_steps, args.cfg_scale
    shape = (B, 3, 32, 32)
    torch.manual_seed(42)
    model = ldm_b200.UNet(3, 3, 64, (1, 2, 4, 8), True, 10, dtype=args.dtype).to(dev)
    model.requires_grad_(False)
    diffusion = ldm_b200.Diffusion(T, dev)
    classes_host = torch.tensor([3])                       # length-1 label, broadcast (main.py:318)
    classes_dev = classes_host.to(dev)
    g = torch.Generator().manual_seed(7 + rank)
    x_T_host = torch.randn(shape, generator=g).pin_memory()
    offset = ldist.shard_offset(B, rank)
    stream = torch.cuda.Stream(dev)

    def device_step(i):
        return diffusion.sample(model, classes_dev, shape, dev, cfg_scale=cfg, seed=1234 + i, sample_offset=offset,
                                return_device=True)

    def e2e_step():
        return diffusion.sample(model, classes_host, shape, dev, cfg_scale=cfg, x_T=x_T_host, seed=99,
                                sample_offset=offset)

    def barrier():
        if world > 1:
            tdist.barrier()

    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            device_step(-1 - i)
        torch.cuda.synchronize(dev)
        clocks = ClockSampler(local_rank).start() if rank == 0 else None
        barrier()
        torch.cuda.synchronize(dev)
        launches0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            out = device_step(i)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        barrier()
        launches = _lib.launch_count() - launches0
        ms_total = ldist.max_over_ranks(e0.elapsed_time(e1), dev)
        clock_rec = clocks.stop() if clocks else None
        assert bool(torch.isfinite(out).all()) or os.environ.get("LDM_BENCH_ALLOW_NONFINITE"), "sampler produced non-finite images"

        # ---- end to end through the public API with host buffers
        e2e_steps = max(1, args.e2e_steps)
        e2e_step()
        torch.cuda.synchronize(dev)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s0.record(stream)
        for _ in range(e2e_steps):
            res = e2e_step()
        s1.record(stream)
        torch.cuda.synchronize(dev)
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        assert res.device.type == "cpu" and tuple(res.shape) == shape
        e2e_ms = ldist.max_over_ranks(max(s0.elapsed_time(s1), wall_ms), dev)

        # ---- the other two settings SURVEY.md 8(d) asks to be reported beside the headline: T = 400 (the reference YAML's
        # n_steps) and cfg_scale = 0 (one UNet pass per timestep); plus the e2e call returning uint8 images
        variants = {}
        if not args.no_variants:
            d400 = ldm_b200.Diffusion(400, dev)

            def timed(fn, reps=1):
                fn()
                torch.cuda.synchronize(dev)
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(reps):
                    fn()
                b.record(stream)
                torch.cuda.synchronize(dev)
                barrier()
                return B * world * reps / (ldist.max_over_ranks(a.elapsed_time(b), dev) / 1e3)

            variants["T400_cfg3_images_per_sec"] = timed(lambda: d400.sample(
                model, classes_dev, shape, dev, cfg_scale=cfg, seed=5, sample_offset=offset, return_device=True))
            variants["T1000_cfg0_images_per_sec"] = timed(lambda: diffusion.sample(
                model, classes_dev, shape, dev, cfg_scale=0, seed=6, sample_offset=offset, return_device=True))
            # generate_images.py (:18-41): one image per call, batch 1 -- latency, not throughput
            one = timed(lambda: diffusion.sample(model, classes_dev, (1, 3, 32, 32), dev, cfg_scale=cfg, seed=7, return_device=True))
            variants["generate_images_batch1_seconds_per_image"] = B * world / one
            # the same with every GroupNorm riding in its producing convolution (opt-in at small batch: ~20 launches fewer per
            # timestep, but a sample's bits then depend on the batch it is generated in)
            os.environ["LDM_GN_NORM_SMALL_BATCH"] = "1"
            d_small = ldm_b200.Diffusion(T, dev)
            one_f = timed(lambda: d_small.sample(model, classes_dev, (1, 3, 32, 32), dev, cfg_scale=cfg, seed=7, return_device=True))
            variants["generate_images_batch1_seconds_per_image_fused_groupnorm"] = B * world / one_f
            del os.environ["LDM_GN_NORM_SMALL_BATCH"]
            del d_small
            # BASELINE config 4: Autoencoder(3,4,3,64,[1,2],2) encode -> 1000-step CFG sampling of the [B,4,16,16] latent
            # with the LatentDiffusionModel's sqrt-linear schedule -> decode (scale 0.18215, 0.00085/0.012)
            torch.manual_seed(43)
            lat_unet = ldm_b200.UNet(4, 4, 64, (1, 2, 4, 8), True, 10, dtype=args.dtype).to(dev)
            lat_unet.requires_grad_(False)
            ae = ldm_b200.Autoencoder(3, 4, 3, 64, [1, 2], 2, dtype=args.dtype).to(dev)
            ldm = ldm_b200.LatentDiffusionModel(lat_unet, ae, 0.18215, T, 0.00085, 0.012).to(dev)
            dl = ldm.make_diffusion(dev)
            imgs = torch.rand(shape, device=dev) * 2 - 1

            def ldm_round():
                z0 = ldm.autoencoder_encode(imgs)                       # encode leg (its latent seeds nothing: timing only)
                z = dl.sample(ldm, classes_dev, tuple(z0.shape), dev, cfg_scale=cfg, seed=8, sample_offset=offset,
                              return_device=True)
                return ldm.autoencoder_decode(z)

            variants["ldm_encode_sample_decode_images_per_sec"] = timed(ldm_round)
            from ldm_b200 import ops as lops
            variants["e2e_uint8_output_images_per_sec"] = timed(lambda: lops.images_to_uint8(diffusion.sample(
                model, classes_host, shape, dev, cfg_scale=cfg, x_T=x_T_host, seed=99, sample_offset=offset,
                return_device=True), "save_image").cpu())

            # ---- bounded-sample variants: K of the T timesteps are timed (every timestep costs the same: one 2B-row UNet
            # pass + the fused update) and the rate is extrapolated to the full trajectory, x T / K
            def rate_bounded(diff, mdl, shp, k, scale=cfg, reps=1):
                def go():
                    return diff.sample(mdl, classes_dev, shp, dev, cfg_scale=scale, seed=11, sample_offset=offset,
                                       first_step=diff.n_steps - 1, num_steps=k, return_device=True)
                go()
                torch.cuda.synchronize(dev)
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(reps):
                    go()
                b.record(stream)
                torch.cuda.synchronize(dev)
                barrier()
                ms = ldist.max_over_ranks(a.elapsed_time(b), dev) / reps
                return shp[0] * world / (ms / 1e3 * diff.n_steps / k)

            # BASELINE config 5 (augmentation-scale generation sweep): batch 64 ... 8192 per GPU
            K = 20
            sweep = {}
            for bb in (64, 256, 1024, 4096, 8192):
                try:
                    sweep[str(bb)] = rate_bounded(diffusion, model, (bb, 3, 32, 32), K)
                except Exception as exc:   # noqa: BLE001 -- e.g. not enough memory on a shared device: say so, keep going
                    sweep[str(bb)] = f"failed: {exc}"
                torch.cuda.empty_cache()
            variants["batch_sweep_images_per_sec"] = dict(sweep, sample=f"{K} of {T} timesteps timed, extrapolated x{T}/{K}")
            # BASELINE config 1: the MNIST-shaped model (1 channel; 32x32 because 28 is not divisible by 2^4)
            torch.manual_seed(44)
            mn = ldm_b200.UNet(1, 1, 64, (1, 2, 4, 8), True, 10, dtype=args.dtype).to(dev)
            mn.requires_grad_(False)
            variants["mnist_shaped_1x32x32_images_per_sec"] = rate_bounded(diffusion, mn, (B, 1, 32, 32), K)
            del mn
            # the fp32 parity path (FFMA implicit GEMM, <= 1e-4 against the reference): reported once, not optimised
            torch.manual_seed(42)
            m32 = ldm_b200.UNet(3, 3, 64, (1, 2, 4, 8), True, 10, dtype="fp32").to(dev)
            m32.requires_grad_(False)
            variants["fp32_parity_path_images_per_sec"] = rate_bounded(diffusion, m32, (64, 3, 32, 32), 4)
            del m32
            torch.cuda.empty_cache()

        # ---- roofline leg: every launch of one 2B-row UNet pass timed with CUDA events on this stream
        prof = None
        if rank == 0:
            rows = 2 * B if cfg > 0 else B
            xx = torch.randn(rows, 3, 32, 32, device=dev)
            tt = torch.full((rows,), T // 2, dtype=torch.long, device=dev)
            model.profile(xx, tt, classes_dev, y_rows=B)
            acc = {}
            reps = 3
            for _ in range(reps):
                for k, v in model.profile(xx, tt, classes_dev, y_rows=B).items():
                    a = acc.setdefault(k, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
                    for kk in a:
                        a[kk] += v[kk]
            prof = {k: {kk: vv / reps for kk, vv in v.items()} for k, v in acc.items()}

    # ---- context: the same algorithm in eager PyTorch on this GPU (cuDNN / cuBLAS / ATen: "the Blackwell library stack",
    # SURVEY.md 8(d)) -- the oracle's functional restatement of src/UNet.py + src/DDPM.py on CUDA tensors, outside every
    # timed region of ours, on a bounded sample; a baseline, never a fallback
    lib_base = None
    if rank == 0 and not args.no_variants:
        lib_base = gpu_library_baseline(dev, B, T, cfg, stream)

    # ---- secondary: the training step of the reference config (batch 64 per GPU, q_sample + fwd + bwd + grad
    # all-reduce + Adam), reported beside the headline, never instead of it
    train = None
    train256 = None
    if not args.no_train:
        train = train_leg(args, dev, rank, world, stream)
        if args.train_batch != 256:   # the reference's batch 64 and 256 per GPU
            train256 = train_leg(args, dev, rank, world, stream, batch=256)
    if rank != 0:
        return
    peaks = load_peaks()
    images = B * world * args.steps
    value = images / (ms_total / 1e3)
    e2e_value = B * world * e2e_steps / (e2e_ms / 1e3)
    passes = 2 if cfg > 0 else 1
    unet_tflops = images * T * passes * UNET_GFLOP_PER_IMAGE / 1e3 / (ms_total / 1e3)
    # the dominant kernel: conv_halo_kernel (the six full-resolution 3x3 convolutions of a pass: the largest single entry of the
    # ncu launch list); the other 3x3 convolutions (conv_tc_kernel) are listed beside it and summed into `family_3x3`
    tc_path = "conv_halo" in prof or "conv_tc" in prof
    conv_name = "conv_halo" if "conv_halo" in prof else ("conv_tc" if "conv_tc" in prof else "conv_ffma")
    conv = prof.get(conv_name) or {"ms": 1e-9, "flops": 0.0, "bytes": 0.0, "launches": 1}
    total_ms = sum(v["ms"] for v in prof.values())
    achieved = conv["flops"] / (conv["ms"] / 1e3) / 1e12
    fam3 = [prof[k] for k in ("conv_halo", "conv_tc") if k in prof]
    f3_ms, f3_fl, f3_n = sum(v["ms"] for v in fam3), sum(v["flops"] for v in fam3), sum(v["launches"] for v in fam3)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_conv_halo_traffic.json")   # tools/ncu_traffic.py over the committed ncu capture
    if conv_name == "conv_halo" and os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        except Exception:
            traffic = None
    roofline = {
        "kernel": {"conv_halo": "conv_halo_kernel: the 3x3 convolutions at full resolution on tcgen05 / TMEM / TMA (slab + halo loaded "
                                "once per tile, 9 taps = descriptor shifts); 45 % of the UNet's FLOPs",
                   "conv_tc": "conv_tc_kernel: 3x3 implicit-GEMM convolutions on tcgen05"}.get(conv_name, "conv_simt"),
        "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
        "frac": achieved / peaks["bf16_tflops_sustained"], "frac_of_burst_peak": achieved / peaks["bf16_tflops"],
        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
        "traffic": traffic, "traffic_source": traffic_src,
        "launches_per_unet_pass": conv["launches"], "flops_per_launch_avg": conv["flops"] / conv["launches"],
        "avg_launch_ms": conv["ms"] / conv["launches"], "share_of_unet_pass": conv["ms"] / total_ms,
        "family_3x3": ({"kernels": "conv_halo_kernel + conv_tc_kernel (every 3x3 convolution; 80 % of the UNet's FLOPs)", "ms": f3_ms,
                        "launches": f3_n, "tflops": f3_fl / (f3_ms / 1e3) / 1e12,
                        "frac": f3_fl / (f3_ms / 1e3) / 1e12 / peaks["bf16_tflops_sustained"]} if tc_path and f3_ms > 0 else None),
        "families": {k: {"ms": round(v["ms"], 4), "launches": v["launches"],
                         "tflops": v["flops"] / (v["ms"] / 1e3) / 1e12 if v["ms"] > 0 else 0.0,
                         "gbs": v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["ms"] > 0 else 0.0,
                         "share": v["ms"] / total_ms} for k, v in prof.items()},
        "whole_step": {"unet_tflops": unet_tflops, "frac_of_sustained_peak": unet_tflops / peaks["bf16_tflops_sustained"],
                       "gflop_per_image_per_unet_pass": UNET_GFLOP_PER_IMAGE},
    }
    cpu = None
    if not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        v, per_step, cores = cpu_sampling_rate(args.cpu_batch, args.cpu_timesteps, T, cfg)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_timesteps} of {T} reverse timesteps at batch {args.cpu_batch} after 1 warm-up timestep "
                         f"({per_step:.2f} s per timestep), extrapolated to the full trajectory; oracle/ (torch CPU fp32)"}
    act_mb = (2 * B if cfg > 0 else B) * 64 * 32 * 32 * 2 / 1e6
    config = workload_config(args, B, world)
    config["l2"] = (f"no flush needed: every timestep streams ~40.7 MB of bf16 weights plus {act_mb:.0f} MB per full-resolution "
                    f"activation tensor (working set >> 126 MB L2)")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic", "config": config,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": x_T_host.numel() * 4 + classes_host.numel() * 8,
                "d2h_bytes_per_step": x_T_host.numel() * 4, "steps": e2e_steps},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clock_rec,
        "unet_tflops": unet_tflops, "variants": variants, "train": train, "train_batch256": train256,
        "gpu_library_baseline": lib_base,
    }
    if lib_base:
        for k in ("fp32_tf32", "bf16_autocast"):
            if isinstance(lib_base.get(k), dict):
                lib_base[k]["ours_over_it"] = value / world / lib_base[k]["images_per_sec"]
    emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's original stdout; everything else (NCCL banners, library chatter) was
    re-routed to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per sampling call")
    ap.add_argument("--n-steps", type=int, default=1000, help="diffusion timesteps T")
    ap.add_argument("--cfg-scale", type=float, default=3.0)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--cpu-batch", type=int, default=64)
    ap.add_argument("--cpu-timesteps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step measurement")
    ap.add_argument("--no-variants", action="store_true", help="skip the T=400 / cfg 0 / uint8-output variant timings")
    ap.add_argument("--train-batch", type=int, default=64)
    ap.add_argument("--train-eager", action="store_true", help="do not capture the training forward/backward as CUDA graphs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        try:
            import torch.distributed as tdist
            if tdist.is_initialized():
                tdist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
