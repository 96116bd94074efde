"""GPU box: call the fused PreNorm + LinearAttention + to_out op a few times at one shape (target for ncu)."""
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
wo, bo = torch.randn(64, 128, 1, 1, device=dev) / 11.3, torch.randn(64, device=dev)
for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 4): ops.linear_attention_prenorm_to_out(x, w, gamma, beta, wo, bo)
torch.cuda.synchronize()
print("done")
