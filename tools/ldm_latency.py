"""GPU box: per-timestep time of the latent UNet sampler ([B,4,16,16], cfg 3)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ldm_b200
dev = torch.device("cuda:0")
torch.manual_seed(43)
m = ldm_b200.UNet(4, 4, 64, (1, 2, 4, 8), True, 10, dtype="bf16").to(dev); m.requires_grad_(False)
d = ldm_b200.Diffusion(1000, dev)
y = torch.tensor([3], device=dev)
d.sample(m, y, (256, 4, 16, 16), dev, cfg_scale=3, seed=1, return_device=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
d.sample(m, y, (256, 4, 16, 16), dev, cfg_scale=3, seed=2, return_device=True)
torch.cuda.synchronize(); print(f"dense_off={os.environ.get('LDM_NO_DENSE2X2','0')}: {(time.perf_counter()-t0):.4f} s per 1000 steps at batch 256")
