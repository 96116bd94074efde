"""GPU box: is the fused conv+GroupNorm output of image 0 bit-identical when it sits in a batch of 1 / 3 / 20 / 40?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldm_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
cin, cout, R, k = 64, 64, 32, 3
Bmax = 40
x = torch.randn(Bmax, R, R, cin, device=dev).bfloat16()
w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
b = torch.randn(cout, device=dev)
gamma, beta = torch.randn(cout, device=dev), torch.randn(cout, device=dev)
rv = torch.randn(Bmax, cout, device=dev)
wp = ops.pack_conv_weight(w, "bf16")
ref = {}
for mode in (0, 1, 2):
    for B in (1, 3, 20, 40, 40):
        if mode == 0:
            o = ops.conv2d(x[:B].contiguous(), wp, k, bias=b)
        elif mode == 1:
            o, st = ops.conv2d_gn(x[:B].contiguous(), wp, k, b, mode=1, groups=8)
            st = st[0].clone()
        else:
            o = ops.conv2d_gn(x[:B].contiguous(), wp, k, b, mode=2, groups=8, gamma=gamma, beta=beta, silu=True, gn_rowvec=rv[:B])
        o0 = o[0].float()
        if B == 1:
            ref[mode] = (o0, st if mode == 1 else None)
            continue
        d = (o0 - ref[mode][0]).abs()
        nz = d > 0
        msg = f"mode {mode} B={B}: differing {int(nz.sum())} / {d.numel()}, max abs {float(d.max()):.3e}"
        if int(nz.sum()):
            idx = nz.nonzero()
            msg += f" rows(h) {sorted(set(idx[:, 0].tolist()))[:12]} chans {sorted(set(idx[:, 2].tolist()))[:16]}"
        if mode == 1:
            msg += f" stats equal {bool(torch.equal(st, ref[mode][1]))}"
        print(msg)

# tile-shape invariance: conv_tc picks other tile shapes for other batch sizes
for (cin, cout, R, k, G, silu, res, Bs) in ((64, 128, 16, 3, 8, True, False, (1, 3, 300)), (128, 64, 32, 1, 1, False, True, (1, 3, 160)),
                                            (128, 512, 4, 1, 1, False, True, (1, 9, 80)), (128, 256, 8, 3, 8, True, False, (1, 5, 300))):
    Bm = max(Bs)
    x = torch.randn(Bm, R, R, cin, device=dev).bfloat16()
    w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device=dev)
    gamma, beta = torch.randn(cout, device=dev), torch.randn(cout, device=dev)
    rv = torch.randn(Bm, cout, device=dev) if not res else None
    rh = torch.randn(Bm, R, R, cout, device=dev).bfloat16() if res else None
    wp = ops.pack_conv_weight(w, "bf16")
    r0 = None
    for B in Bs:
        o = ops.conv2d_gn(x[:B].contiguous(), wp, k, b, mode=2, groups=G, gamma=gamma, beta=beta, silu=silu,
                          gn_rowvec=rv[:B] if rv is not None else None, gn_res=rh[:B].contiguous() if rh is not None else None)
        _, st = ops.conv2d_gn(x[:B].contiguous(), wp, k, b, mode=1, groups=G)
        if r0 is None:
            r0 = (o[0].clone(), st[0].clone())
        else:
            print(f"{cin}->{cout}@{R} k{k} G={G} B={B}: out equal {bool(torch.equal(o[0], r0[0]))} stats equal {bool(torch.equal(st[0], r0[1]))}")
