"""GPU box: per-kernel time of Autoencoder encode + decode at batch 256 (torch.profiler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ldm_b200
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
ae = ldm_b200.Autoencoder(3, 4, 3, 64, [1, 2], 2, dtype="bf16").to(dev)
img = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
for _ in range(2):
    z = ae.encode(img).sample(); rec = ae.decode(z)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(); z = ae.encode(img).sample(); e1.record(); rec = ae.decode(z); e2.record(); torch.cuda.synchronize()
print(f"B={B}: encode {e0.elapsed_time(e1):.2f} ms ({ae.last_launches} launches in decode), decode {e1.elapsed_time(e2):.2f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    z = ae.encode(img).sample(); rec = ae.decode(z); torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
print(f"total kernel time {sum(r.device_time_total for r in rows)/1e3:.2f} ms")
for r in rows[:14]:
    print(f"{r.device_time_total/1e3:8.3f} ms {r.count:5d}x  {r.key[:100]}")
