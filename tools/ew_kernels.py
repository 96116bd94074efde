"""GPU box: the update kernels at the headline batch (ncu target): p_sample with fused CFG + in-kernel Philox, q_sample."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ldm_b200
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
d = ldm_b200.Diffusion(1000, dev)
x, ec, eu = (torch.randn(B, 3, 32, 32, device=dev) for _ in range(3))
t = torch.tensor([500], device=dev)
tb = torch.randint(0, 1000, (B,), device=dev)
for _ in range(5):
    d.p_sample(x, t, ec, eps_uncond=eu, cfg_scale=3.0, seed=1)
    d.q_sample(x, tb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn, nbytes in (("p_sample (CFG, Philox)", lambda: d.p_sample(x, t, ec, eps_uncond=eu, cfg_scale=3.0, seed=1), 49152 * B),
                         ("q_sample (Philox)", lambda: d.q_sample(x, tb), 36864 * B)):
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print(f"{name}: {us:.2f} us per call incl. host wrapper; algorithmic {nbytes/1e6:.2f} MB")
