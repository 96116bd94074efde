"""GPU box: training step with the UNet forward+backward captured as CUDA graphs (torch.cuda.make_graphed_callables)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ldm_b200
from ldm_b200 import _lib
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
m = ldm_b200.UNet(3, 3, 64, (1, 2, 4, 8), True, 10, dtype="bf16").to(dev)
ref = ldm_b200.UNet(3, 3, 64, (1, 2, 4, 8), True, 10, dtype="bf16").to(dev)
ref.load_state_dict(m.state_dict())
d = ldm_b200.Diffusion(1000, dev)
x0 = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
y = torch.randint(0, 10, (B,), device=dev)
noise, xt, t = d(x0)
sx, st_, sy = xt.clone(), t.clone(), y.clone()
gm = torch.cuda.make_graphed_callables(m, (sx, st_, sy), allow_unused_input=True)
opt = torch.optim.Adam(m.parameters(), lr=5e-4, capturable=False)
# parity of one graphed step against the eager path
loss_g = torch.nn.functional.mse_loss(noise, gm(xt, t, y)); loss_g.backward()
loss_e = torch.nn.functional.mse_loss(noise, ref(xt, t, y)); loss_e.backward()
worst = max(float((a.grad - b.grad).norm() / b.grad.norm()) for a, b in zip(m.parameters(), ref.parameters()) if b.grad is not None)
print("graphed vs eager: loss", float(loss_g), float(loss_e), "worst grad rel diff", worst)
def step():
    noise, xt, t = d(x0)
    loss = torch.nn.functional.mse_loss(noise, gm(xt, t, y))
    opt.zero_grad(set_to_none=False)
    loss.backward()
    opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): l = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"B={B} graphed train step {ms:.2f} ms  {B/ms*1e3:.0f} img/s  loss {float(l):.4f}")
