"""GPU box: per-launch CUDA-event times of one UNet pass (warm), printed by the library (LDM_PROFILE_DUMP=1)."""
import os, sys
os.environ["LDM_PROFILE_DUMP"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ldm_b200
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
torch.manual_seed(42)
m = ldm_b200.UNet(3, 3, 64, (1, 2, 4, 8), True, 10, dtype="bf16").to(dev)
m.requires_grad_(False)
x = torch.randn(rows, 3, 32, 32, device=dev)
t = torch.full((rows,), 500, dtype=torch.long, device=dev)
y = torch.tensor([3], device=dev)
for i in range(3):
    if i == 2: sys.stderr.write("=== pass\n")
    p = m.profile(x, t, y, y_rows=rows // 2)
tot = sum(v["ms"] for v in p.values())
for k, v in p.items(): print(f"{k:18s} {v['ms']:.3f} ms  {v['launches']} launches  {100*v['ms']/tot:.1f}%")
print("total", tot)
