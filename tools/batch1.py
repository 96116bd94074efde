"""GPU box: latency of generate_images.py's call (batch 1, T=1000, cfg 3)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ldm_b200
dev = torch.device("cuda:0")
torch.manual_seed(42)
m = ldm_b200.UNet(3, 3, 64, (1, 2, 4, 8), True, 10, dtype="bf16").to(dev); m.requires_grad_(False)
d = ldm_b200.Diffusion(1000, dev)
for B in (1, 4, 16):
    y = torch.tensor([3], device=dev)
    d.sample(m, y, (B, 3, 32, 32), dev, cfg_scale=3, seed=1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(2): d.sample(m, y, (B, 3, 32, 32), dev, cfg_scale=3, seed=2 + i)
    torch.cuda.synchronize(); print(f"PDL={os.environ.get('LDM_PDL','0')} B={B}: {(time.perf_counter()-t0)/2*1e3:.1f} ms per call")
