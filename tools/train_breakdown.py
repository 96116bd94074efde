"""GPU box: where the training step's time goes (CUDA events) + per-kernel launch shares via LDM launch counter."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ldm_b200
from ldm_b200 import _lib
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
torch.manual_seed(0)
m = ldm_b200.UNet(3, 3, 64, (1, 2, 4, 8), True, 10, dtype=dtype).to(dev)
d = ldm_b200.Diffusion(1000, dev)
opt = torch.optim.Adam(m.parameters(), lr=5e-4)
x0 = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
y = torch.randint(0, 10, (B,), device=dev)
def ev(): return torch.cuda.Event(enable_timing=True)
tot = {"q_sample": 0, "forward": 0, "backward": 0, "adam": 0}
wall = 0
for it in range(8):
    e = [ev() for _ in range(5)]
    t0 = time.perf_counter()
    e[0].record(); noise, xt, t = d(x0)
    e[1].record(); loss = torch.nn.functional.mse_loss(noise, m(xt, t, y))
    e[2].record(); opt.zero_grad(set_to_none=True); loss.backward()
    e[3].record(); opt.step()
    e[4].record(); torch.cuda.synchronize()
    if it >= 3:
        wall += time.perf_counter() - t0
        for k, (a, b) in zip(tot, zip(e[:-1], e[1:])): tot[k] += a.elapsed_time(b)
n = 5
print(f"B={B} {dtype}: " + "  ".join(f"{k} {v/n:.2f} ms" for k, v in tot.items()) + f"  | wall {wall/n*1e3:.2f} ms  launches/step {_lib.launch_count()//8}")
