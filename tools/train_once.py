import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ldm_b200
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
m = ldm_b200.UNet(3, 3, 64, (1, 2, 4, 8), True, 10, dtype="bf16").to(dev)
d = ldm_b200.Diffusion(1000, dev)
x0 = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
y = torch.randint(0, 10, (B,), device=dev)
for _ in range(2):
    noise, xt, t = d(x0)
    loss = torch.nn.functional.mse_loss(noise, m(xt, t, y))
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss.detach()))
