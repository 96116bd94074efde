#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (test infrastructure).

Run in the build container, where the reference is mounted read-only:

    python oracle/make_golden.py [--ref /root/reference] [--out tests/golden] [--skip-trajectory]

The reference has no tests or fixtures of its own, so these
files -- outputs of the reference's own ``src.UNet`` / ``src.DDPM`` modules on seeded
inputs -- are the parity pin for both the oracle and the CUDA path.  Weights are
never stored (81 MB): every case draws them with ``torch.manual_seed(seed)`` +
PyTorch default init, which ``oracle.init_state_dict`` reproduces bit for bit
(asserted below and in tests via a checksum).
"""
from __future__ import annotations

import argparse
import hashlib
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)


def sd_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "golden"))
    ap.add_argument("--skip-trajectory", action="store_true")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    sys.path.insert(0, args.ref)
    os.environ.setdefault("WANDB_MODE", "disabled")
    np.Inf = np.inf  # shim for src/EarlyStopping.py:31 under NumPy 2 (harness-side only)

    from src.UNet import UNet, Block, ResNetBlock, LinearAttention, Attention, TimeEmbedding  # noqa
    from src.DDPM import Diffusion
    from oracle import init_state_dict

    torch.set_num_threads(os.cpu_count())
    save = lambda name, **kw: np.savez_compressed(os.path.join(args.out, name), **{
        k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in kw.items()})

    # ---------------- G1: single UNet pass, CIFAR- and MNIST-shaped ----------------
    for tag, cin in (("cifar", 3), ("mnist", 1)):
        torch.manual_seed(0)
        model = UNet(cin, cin, 64, [1, 2, 4, 8], True, 10).eval()
        sd = model.state_dict()
        ours = init_state_dict(0, cin, cin, 64, (1, 2, 4, 8), True, 10)
        assert list(ours.keys()) == list(sd.keys())
        assert all(torch.equal(ours[k], sd[k]) for k in sd), "default-init replication broke"
        g = torch.Generator().manual_seed(1)
        B = 4
        x = torch.randn(B, cin, 32, 32, generator=g)
        t = torch.randint(0, 1000, (B,), generator=g)
        y = torch.randint(0, 10, (B,), generator=g)
        with torch.no_grad():
            eps_c = model(x, t, y)
            eps_u = model(x, t, None)
            eps_b = model(x, t, torch.tensor([3]))
        save(f"g1_unet_{tag}.npz", x=x, t=t, y=y, eps_cond=eps_c, eps_uncond=eps_u, eps_bcast3=eps_b,
             weight_seed=0, weight_sha256=np.frombuffer(bytes.fromhex(sd_checksum(sd)), dtype=np.uint8))
        print("G1", tag, float(eps_c.std()))

        if tag == "cifar":
            # ---------------- G5: training-step gradients (fwd + bwd, MSE) ----------------
            model.train()
            g = torch.Generator().manual_seed(5)
            x0 = torch.rand(B, cin, 32, 32, generator=g) * 2 - 1
            noise = torch.randn(B, cin, 32, 32, generator=g)
            tt = torch.randint(0, 1000, (B,), generator=g)
            yy = torch.randint(0, 10, (B,), generator=g)
            diff = Diffusion(1000, "cpu")
            xt = diff.q_sample(x0, tt, eps=noise).requires_grad_(True)
            eps = model(xt, tt, yy)
            loss = torch.nn.functional.mse_loss(noise, eps)
            loss.backward()
            names, gsum, gnorm, ghead, hasg = [], [], [], [], []
            for k, p in model.named_parameters():
                names.append(k)
                if p.grad is None:
                    hasg.append(0); gsum.append(0.0); gnorm.append(0.0); ghead.append(np.zeros(4, np.float32))
                else:
                    hasg.append(1)
                    gsum.append(float(p.grad.double().sum()))
                    gnorm.append(float(p.grad.double().norm()))
                    hd = np.zeros(4, np.float32); v = p.grad.reshape(-1)[:4].numpy(); hd[:v.size] = v; ghead.append(hd)
            save("g5_train_grads.npz", x0=x0, noise=noise, t=tt, y=yy, xt=xt, eps=eps, loss=loss,
                 dx=xt.grad, names=np.array(names), grad_sum=np.array(gsum), grad_norm=np.array(gnorm),
                 grad_head=np.stack(ghead), has_grad=np.array(hasg), weight_seed=0)
            print("G5 loss", float(loss), "params without grad:", [n for n, h in zip(names, hasg) if not h])

    # ---------------- G2: p_sample, with the reference's own RNG draw ----------------
    for T in (1000, 400):
        diff = Diffusion(T, "cpu")
        g = torch.Generator().manual_seed(2)
        xt = torch.randn(4, 3, 32, 32, generator=g) * 1.7
        eps = torch.randn(4, 3, 32, 32, generator=g)
        outs, noises, ts = [], [], []
        for step in (T - 1, T // 2, 1, 0):
            t = torch.full((4,), step, dtype=torch.long)
            torch.manual_seed(100 + step)
            out = diff.p_sample(xt, t, eps)
            torch.manual_seed(100 + step)
            noises.append(torch.randn(xt.shape))  # what the reference drew at :92 (unused when t==0)
            outs.append(out); ts.append(step)
        save(f"g2_p_sample_T{T}.npz", xt=xt, eps=eps, steps=np.array(ts), out=torch.stack(outs),
             noise=torch.stack(noises), beta=diff.beta, alpha=diff.alpha, alpha_bar=diff.alpha_bar)
    print("G2 done")

    # ---------------- G3: Diffusion.forward / q_sample ----------------
    diff = Diffusion(1000, "cpu")
    g = torch.Generator().manual_seed(3)
    x0 = torch.rand(8, 3, 32, 32, generator=g) * 2 - 1
    torch.manual_seed(33)
    noise, xt, t = diff(x0)
    save("g3_q_sample.npz", x0=x0, noise=noise, xt=xt, t=t, seed=33)
    print("G3 done")

    # ---------------- G4: sub-modules (weights small enough to store) ----------------
    def mod_case(name, mod, *inputs):
        mod.eval()
        with torch.no_grad():
            out = mod(*inputs)
        kw = {f"w::{k}": v for k, v in mod.state_dict().items()}
        kw.update({f"in{i}": v for i, v in enumerate(inputs) if v is not None})
        save(f"g4_{name}.npz", out=out, **kw)

    torch.manual_seed(40)
    mod_case("block_64_64_r8", Block(64, 64), torch.randn(2, 64, 8, 8))
    torch.manual_seed(41)
    mod_case("resblock_64_128_t_r8", ResNetBlock(64, 128, time_emb_dim=256), torch.randn(2, 64, 8, 8), torch.randn(2, 256))
    torch.manual_seed(42)
    mod_case("resblock_64_64_not_r16", ResNetBlock(64, 64), torch.randn(2, 64, 16, 16))
    torch.manual_seed(43)
    mod_case("linattn_64_n1024", LinearAttention(64), torch.randn(2, 64, 32, 32))
    torch.manual_seed(44)
    mod_case("linattn_128_n16", LinearAttention(128), torch.randn(3, 128, 4, 4))
    torch.manual_seed(45)
    mod_case("attn_512_n4", Attention(512), torch.randn(2, 512, 2, 2))
    torch.manual_seed(46)
    mod_case("attn_64_n64", Attention(64), torch.randn(2, 64, 8, 8))
    torch.manual_seed(47)
    mod_case("time_embedding_256", TimeEmbedding(256), torch.tensor([0, 1, 500, 999]))
    print("G4 done")

    # ---------------- G6: free-running trajectory through the reference's own loop ----------------
    if not args.skip_trajectory:
        for T, B in ((1000, 2),):
            torch.manual_seed(42)
            model = UNet(3, 3, 64, [1, 2, 4, 8], True, 10).eval()
            diff = Diffusion(T, "cpu")
            stats = np.zeros((T, 3), np.float64)   # per step: mean, std, absmax of x_t fed to the model
            keep = {}
            state = {"i": 0}

            def rec(x, t, y, _m=model):
                if y is not None:   # cond pass comes first each step
                    s = int(t[0])
                    stats[s] = (float(x.mean()), float(x.std()), float(x.abs().max()))
                    if s in (T - 1, T - 10, T - 100, T // 2, 0):
                        keep[s] = x.clone()
                return _m(x, t, y)
            rec.num_classes = 10
            torch.manual_seed(7)
            t0 = time.time()
            out = diff.sample(rec, torch.tensor([3]), shape=(B, 3, 32, 32), device="cpu", cfg_scale=3)
            print(f"G6 reference trajectory T={T} B={B}: {time.time()-t0:.1f}s, final std {float(out.std()):.3f}")
            save(f"g6_trajectory_T{T}.npz", out=out, stats=stats, weight_seed=42, noise_seed=7,
                 classes=np.array([3]), cfg_scale=3.0,
                 **{f"x_at_{s}": v for s, v in keep.items()})


if __name__ == "__main__":
    main()
