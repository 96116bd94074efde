"""GPU box: time the convolution + fused GroupNorm epilogue (ops.conv2d_gn) against conv2d followed by group_norm.
usage: bench_conv_gn.py B cin cout R k groups silu rowvec gn_res [nvar]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldm_b200 import ops, _lib
B, cin, cout, R, k, G, silu, rowvec, gn_res = [int(v) for v in sys.argv[1:10]]
nvar = int(sys.argv[10]) if len(sys.argv) > 10 else 1
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(B, R, R, cin, device=dev).bfloat16()
w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
b = torch.randn(cout, device=dev)
gamma, beta = torch.randn(cout, device=dev), torch.randn(cout, device=dev)
rv = torch.randn(B * nvar, cout, device=dev) if rowvec else None
rh = torch.randn(B, R, R, cout, device=dev).bfloat16() if gn_res else None
wp = ops.pack_conv_weight(w, "bf16")
out = torch.empty(B * nvar, R, R, cout, device=dev, dtype=torch.bfloat16)
tmp = torch.empty(B, R, R, cout, device=dev, dtype=torch.bfloat16)
scratch = torch.zeros(_lib.load().ldm_conv2d_gn_scratch_bytes(B * nvar), dtype=torch.uint8, device=dev)
tag = [0]


def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def fused(mode):
    def f():
        tag[0] += 1
        ops.conv2d_gn(x, wp, k, b, mode=mode, groups=G, gamma=gamma, beta=beta, silu=bool(silu), gn_rowvec=rv if mode == 2 else None,
                      gn_res=rh if mode == 2 else None, nvar=nvar if mode == 2 else 1, out=out if mode == 2 else tmp, scratch=scratch, tag=tag[0])
    return f


def separate():
    ops.conv2d(x, wp, k, bias=b, out=tmp)
    ops.group_norm(tmp, gamma, beta, G, silu=bool(silu), res=rh, out=out[:B])


tag_env = os.environ.get("LDM_EPI_DEBUG", "0")
cres = torch.randn(B, R, R, cout, device=dev).bfloat16() if os.environ.get("CONV_RES") else None   # conv-level residual (conv2)
t0 = timeit(lambda: ops.conv2d(x, wp, k, bias=b, res=cres, out=tmp))


def stats_only():
    ops.conv2d_gn(x, wp, k, b, mode=1, groups=G, res=cres, out=tmp, scratch=scratch)


t1 = timeit(stats_only)
t2 = timeit(fused(2))
ts = timeit(separate) if not rowvec and nvar == 1 else float("nan")
print(f"B={B} {cin}->{cout} @{R} k{k} G={G} silu={silu} rv={rowvec} res={gn_res} nvar={nvar} dbg={tag_env}: conv {t0:.1f} us | +stats {t1:.1f} | fused norm {t2:.1f} | conv+gn kernels {ts:.1f}")
