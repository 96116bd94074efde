"""GPU box: N UNet eps-predictions at a given row count (ncu target: the last pass is the steady-state one)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ldm_b200
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dtype = sys.argv[3] if len(sys.argv) > 3 else "bf16"
dev = torch.device("cuda:0")
torch.manual_seed(42)
m = ldm_b200.UNet(3, 3, 64, (1, 2, 4, 8), True, 10, dtype=dtype).to(dev)
m.requires_grad_(False)
x = torch.randn(rows, 3, 32, 32, device=dev)
t = torch.full((rows,), 500, dtype=torch.long, device=dev)
y = torch.tensor([3], device=dev)
for _ in range(reps):
    out = m._forward_nograd(x, t, y, y_rows=rows // 2)
torch.cuda.synchronize()
print("launches per pass:", m.last_launches, "finite:", bool(torch.isfinite(out).all()))
