import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
import torch.nn.functional as F
from ldm_b200 import _lib, ops
lib = _lib.load()
dev = torch.device("cuda:0")
B, cin, cout, R, k = [int(v) for v in sys.argv[1:6]]
g = torch.Generator().manual_seed(0)
x = torch.randn(B, cin, R, R, generator=g).to(dev).bfloat16().float()
dy = torch.randn(B, cout, R, R, generator=g).to(dev).bfloat16().float()
w = torch.zeros(cout, cin, k, k, device=dev, requires_grad=True)
F.conv2d(x, w, None, padding=k // 2).backward(dy)
xh, dyh = ops.to_nhwc(x, "bf16"), ops.to_nhwc(dy, "bf16")
dw = torch.zeros(cout, cin, k, k, device=dev)
db = torch.zeros(cout, device=dev)
n = lib.ldm_conv2d_wgrad_scratch_bytes(cin, cout, B, R, R, k, 1)
print("scratch", n, flush=True)
scr = torch.empty(n, dtype=torch.uint8, device=dev)
rc = lib.ldm_conv2d_wgrad_tc(xh.data_ptr(), cin, cin, dyh.data_ptr(), cout, cout, dw.data_ptr(), db.data_ptr(), B, R, R, k, scr.data_ptr(), None)
print("rc", rc, lib.ldm_last_error(), flush=True)
torch.cuda.synchronize()
print("dw nan", int(torch.isnan(dw).sum()), "inf", int(torch.isinf(dw).sum()), "absmax", float(dw.nan_to_num().abs().max()), "ref absmax", float(w.grad.abs().max()))
print("dw[0,:4]", dw[0, :4].flatten()[:8].tolist(), "ref", w.grad[0, :4].flatten()[:8].tolist())
print("rel err dw", float((dw - w.grad).norm() / w.grad.norm()), "db", float((db - dy.sum((0, 2, 3))).norm() / dy.sum((0, 2, 3)).norm()))
