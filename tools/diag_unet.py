"""Print stage-by-stage relative errors of the native UNet against the oracle (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ldm_b200, oracle
from oracle import unet_oracle as U
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
def rel(a, b): a=a.double().cpu(); b=b.double().cpu(); return float((a-b).norm()/b.norm())
names = ["temb", "initial", "enc0.res", "enc0.attn", "enc1.res", "enc1.attn", "enc2.attn", "enc3.res", "enc3.attn", "bottleneck", "dec0", "dec1", "dec2", "dec3", "final.res"]
for tag, cin in (("cifar", 3), ("mnist", 1)):
    g = np.load(f"tests/golden/g1_unet_{tag}.npz")
    sd = oracle.init_state_dict(0, cin, cin, 64, (1, 2, 4, 8), True, 10)
    x, t, y = (torch.from_numpy(g[k]) for k in ("x", "t", "y"))
    taps = {}
    with torch.no_grad():
        U.unet_forward(sd, x, t, y, taps=taps)
        # the reference under bf16 autocast on this GPU, for calibration
        sdd = {k: v.to(dev) for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ac = U.unet_forward(sdd, x.to(dev), t.to(dev), y.to(dev))
        print(tag, "oracle-under-bf16-autocast vs golden:", rel(ac.float(), torch.from_numpy(g["eps_cond"])))
    for dtype, impl in (("fp32", 0), ("bf16", 0), ("bf16", 1)):
        m = ldm_b200.UNet(cin, cin, 64, (1, 2, 4, 8), True, 10, dtype=dtype, conv_impl=impl).to(dev)
        m.load_state_dict(sd)
        with torch.no_grad():
            out = m(x.to(dev), t.to(dev), y.to(dev))
            line = [f"{tag} {dtype} impl={impl} eps: {rel(out, torch.from_numpy(g['eps_cond'])):.3e} |"]
            for nme in names:
                want = taps.get(nme if nme != "final.res" else None)
                if want is None: continue
                buf = torch.empty(want.shape, dtype=torch.float32, device=dev)
                m.set_tap(nme, buf, 32); m(x.to(dev), t.to(dev), y.to(dev)); m.set_tap(None, None, 32)
                line.append(f"{nme}={rel(buf, want):.2e}")
        print(" ".join(line), flush=True)
