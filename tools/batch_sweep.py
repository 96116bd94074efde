"""GPU box: the augmentation-scale generation sweep (BASELINE config 5): images/s of the 1000-step CFG sampler at batch
64 ... 8192 per GPU, plus the batch-invariance check (sample i of a large batch == the same global sample in a batch of 8).
Usage: python tools/batch_sweep.py [batches...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ldm_b200
dev = torch.device("cuda:0")
batches = [int(a) for a in sys.argv[1:]] or [64, 256, 1024, 4096, 8192]
torch.manual_seed(42)
model = ldm_b200.UNet(3, 3, 64, (1, 2, 4, 8), True, 10, dtype="bf16").to(dev)
model.requires_grad_(False)
d = ldm_b200.Diffusion(1000, dev)
y = torch.tensor([3], device=dev)
small = d.sample(model, y, (8, 3, 32, 32), dev, cfg_scale=3, seed=11, return_device=True, first_step=999, num_steps=6)
out = []
for B in batches:
    shape = (B, 3, 32, 32)
    try:
        big = d.sample(model, y, shape, dev, cfg_scale=3, seed=11, return_device=True, first_step=999, num_steps=6)
        inv = float((big[:8] - small).norm() / small.norm())
        tail = d.sample(model, y, (8, 3, 32, 32), dev, cfg_scale=3, seed=11, sample_offset=B - 8, return_device=True,
                        first_step=999, num_steps=6)
        inv_tail = float((big[-8:] - tail).norm() / tail.norm())
        steps = 1000 if B <= 1024 else 200          # big batches: time 200 timesteps (steps are cost-identical)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x = d.sample(model, y, shape, dev, cfg_scale=3, seed=12, return_device=True, first_step=999, num_steps=steps)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) * (1000 / steps)
        rec = {"batch": B, "images_per_sec": B / (ms / 1e3), "ms_per_1000_steps": ms, "timed_steps": steps,
               "invariance_rel_l2_first8": inv, "invariance_rel_l2_last8": inv_tail, "finite": bool(torch.isfinite(x).all()),
               "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    except Exception as exc:  # noqa: BLE001
        rec = {"batch": B, "error": str(exc)[:300]}
    print(json.dumps(rec), flush=True)
    out.append(rec)
    d._samplers.clear() if hasattr(d, "_samplers") else None
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/batch_sweep.json", "w"), indent=1)
