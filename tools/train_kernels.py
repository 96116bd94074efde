"""GPU box: per-kernel time of one training step (torch.profiler / CUPTI), aggregated by kernel name."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ldm_b200
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
m = ldm_b200.UNet(3, 3, 64, (1, 2, 4, 8), True, 10, dtype="bf16").to(dev)
d = ldm_b200.Diffusion(1000, dev)
opt = torch.optim.Adam(m.parameters(), lr=5e-4)
x0 = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
y = torch.randint(0, 10, (B,), device=dev)
def step():
    noise, xt, t = d(x0)
    loss = torch.nn.functional.mse_loss(noise, m(xt, t, y))
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
tot = sum(r.device_time_total for r in rows)
print(f"B={B}: total kernel time {tot/1e3:.2f} ms")
for r in rows[:28]:
    print(f"{r.device_time_total/1e3:8.3f} ms {r.count:5d}x  {r.key[:110]}")
