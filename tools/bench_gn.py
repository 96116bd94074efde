"""GPU box: GroupNorm apply / stats against a plain copy of the same bytes (rotating buffers larger than L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ldm_b200
from ldm_b200 import ops, _lib
dev = torch.device("cuda:0")
def timeit(fn, n=40):
    for _ in range(5): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (B, R, C, G) in [(512, 32, 64, 8), (512, 16, 128, 8), (512, 8, 256, 8), (512, 32, 64, 1), (512, 32, 128, 8)]:
    NB = 6
    xs = [torch.randn(B, R, R, C, device=dev).to(torch.bfloat16) for _ in range(NB)]
    ys = [torch.empty_like(xs[0]) for _ in range(NB)]
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    mb = xs[0].numel() * 2 / 1e6
    t_copy = timeit(lambda i: ys[i % NB].copy_(xs[i % NB]))
    t_gn = timeit(lambda i: ops.group_norm(xs[i % NB], gamma, beta, G, silu=True, out=ys[i % NB]))
    t_mul = timeit(lambda i: torch.mul(xs[i % NB], 1.5, out=ys[i % NB]))
    t_silu = timeit(lambda i: torch.nn.functional.silu(xs[i % NB]))
    print(f"B={B} {R}x{R}x{C} G={G}: {mb:.0f} MB  copy {t_copy:.1f} us ({2*mb/t_copy:.0f} GB/s)  torch.mul {t_mul:.1f} us  torch silu(alloc) {t_silu:.1f} us  "
          f"gn stats+apply {t_gn:.1f} us ({3*mb/t_gn:.0f} GB/s for 3 passes)")
from torch.profiler import profile, ProfilerActivity
B, R, C, G = 512, 32, 64, 8
xs = [torch.randn(B, R, R, C, device=dev).to(torch.bfloat16) for _ in range(6)]
ys = [torch.empty_like(xs[0]) for _ in range(6)]
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
for i in range(6): ops.group_norm(xs[i], gamma, beta, G, silu=True, out=ys[i])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(12): ops.group_norm(xs[i % 6], gamma, beta, G, silu=True, out=ys[i % 6])
    torch.cuda.synchronize()
for r in sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:4]:
    print(f"{r.device_time_total/r.count:8.2f} us avg {r.count:4d}x {r.key[:90]}")
