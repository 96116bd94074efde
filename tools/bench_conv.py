"""GPU box: time one convolution shape through the C ABI (CUDA events, warm)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldm_b200 import ops
B, cin, cout, R, k = [int(v) for v in sys.argv[1:6]]
second = int(sys.argv[6]) if len(sys.argv) > 6 else 0
dev = torch.device("cuda:0")
x = torch.randn(B, R, R, cin, device=dev).bfloat16()
w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
w2 = torch.randn(cout, second, 1, 1, device=dev) if second else None
x2 = torch.randn(B, R, R, second, device=dev).bfloat16() if second else None
b = torch.randn(cout, device=dev) if not os.environ.get("NOBIAS") else None
rv = torch.randn(B, cout, device=dev) if os.environ.get("ROWVEC") else None
wp = ops.pack_conv_weight(w, "bf16", w2)
out = torch.empty(B, R, R, cout, device=dev, dtype=torch.bfloat16)
for _ in range(3): ops.conv2d(x, wp, k, bias=b, x2=x2, rowvec=rv, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps): ops.conv2d(x, wp, k, bias=b, x2=x2, rowvec=rv, out=out)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
fl = 2.0 * B * R * R * cout * (k * k * cin + second)
print(f"conv B={B} {cin}->{cout} @{R} k{k} +{second}: {us:.1f} us  {fl/us/1e6:.1f} TFLOP/s  env={os.environ.get('LDM_HALO_DEBUG','-')} halo={os.environ.get('LDM_CONV_HALO','-')}")
if int(os.environ.get("LDM_HALO_DEBUG", "0")) & 4:
    import ctypes as C
    from ldm_b200 import _lib
    lib = _lib.load()
    buf = (C.c_ulonglong * 1024)()
    lib.ldm_debug_read_halo(buf, 1024)
    for it in (4, 8):
        dns = buf[it*16+3] - buf[it*16+2]; dcy = buf[512+it*16+3] - buf[512+it*16+2]
        print(f"tile {it}: issue span {dns} ns = {dcy} cycles -> {dcy/dns:.3f} GHz; per MMA {dcy/36:.1f} cycles")
    t0 = buf[0]
    names = {0: "mma:pre_tempty", 1: "mma:got_tempty", 2: "mma:got_afull", 3: "mma:committed", 4: "epi:pre_tfull", 5: "epi:got_tfull", 6: "epi:released", 7: "epi:tile_end", 8: "prod:pre_aempty", 9: "prod:got_aempty"}
    for it in range(8, 20):
        print(f"tile {it}: " + "  ".join(f"{names[k]}={buf[it*16+k]-t0}" for k in sorted(names)))

if int(os.environ.get("LDM_TC_DEBUG", "0")):
    import ctypes as C
    from ldm_b200 import _lib
    lib = _lib.load()
    buf = (C.c_ulonglong * 1024)()
    lib.ldm_debug_read_tc(buf, 1024)
    t0 = buf[0]
    names = {0: "mma:pre_tempty", 1: "mma:got_tempty", 2: "mma:got_full0", 3: "mma:committed", 4: "epi:pre_tfull", 5: "epi:got_tfull", 6: "epi:arrived"}
    for it in range(20, 28):
        print(f"tile {it}: " + "  ".join(f"{names[k]}={buf[it*16+k]-t0}" for k in sorted(names)))
