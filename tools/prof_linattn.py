"""GPU box: per-kernel times (torch.profiler / CUPTI) of the fused PreNorm + LinearAttention + to_out call at one shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from ldm_b200 import ops
B, R = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(B, R, R, 64, device=dev).bfloat16()
w = torch.randn(384, 64, 1, 1, device=dev) * 0.25
gamma, beta = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1
wo, bo = torch.randn(64, 128, 1, 1, device=dev) / 11.3, torch.randn(64, device=dev)
for _ in range(3): ops.linear_attention_prenorm_to_out(x, w, gamma, beta, wo, bo)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10): ops.linear_attention_prenorm_to_out(x, w, gamma, beta, wo, bo)
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
    print(f"{e.key[:70]:70s} {e.device_time_total / e.count:9.1f} us x{e.count}")
