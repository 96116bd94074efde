"""GPU box: time the PreNorm + to_qkv + LinearAttention kernels (tcgen05 vs mma.sync) at one shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldm_b200 import ops
B, R = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(B, R, R, 64, device=dev).bfloat16()
w = torch.randn(384, 64, 1, 1, device=dev) * 0.25
gamma, beta = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1
for impl in (0, 1):
    for _ in range(3): ops.linear_attention_prenorm(x, w, gamma, beta, impl=impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.linear_attention_prenorm(x, w, gamma, beta, impl=impl)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} R={R} impl={impl} ({'tcgen05' if impl == 0 else 'mma.sync'}): {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call (fold + stats + attention)")
wo, bo = torch.randn(64, 128, 1, 1, device=dev) / 11.3, torch.randn(64, device=dev)
for _ in range(3): ops.linear_attention_prenorm_to_out(x, w, gamma, beta, wo, bo)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.linear_attention_prenorm_to_out(x, w, gamma, beta, wo, bo)
e1.record(); torch.cuda.synchronize()
print(f"B={B} R={R} tcgen05 + folded to_out: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call (fold + pack + stats + attention)")
