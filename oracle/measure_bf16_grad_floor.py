#!/usr/bin/env python
"""Measure the reference's own bf16-autocast gradient error (test infrastructure; build container only).

Runs the UNMODIFIED reference UNet on golden G5's inputs twice on the CPU -- fp32, and under
torch.autocast(bfloat16) as ``src/DiffusionModelTrainer.py:40`` does with AMP (the reference uses fp16 + GradScaler on
CUDA; bf16 is the same mechanism with the product path's operand type) -- and records, per parameter, the relative error of
the gradient norm and the element-wise relative L2 error.  The numbers calibrate the bf16 tolerances of
tests/test_train_gpu.py (about 3x the floor measured here); the output is committed as tests/golden/bf16_grad_floor.json.

    python oracle/measure_bf16_grad_floor.py [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "golden", "bf16_grad_floor.json"))
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    from src.UNet import UNet          # the unmodified reference
    g = np.load(os.path.join(REPO, "tests", "golden", "g5_train_grads.npz"))
    xt, noise = torch.from_numpy(g["xt"]), torch.from_numpy(g["noise"])
    t, y = torch.from_numpy(g["t"]), torch.from_numpy(g["y"])

    def grads(autocast: bool):
        torch.manual_seed(int(g["weight_seed"]))
        m = UNet(3, 3, 64, [1, 2, 4, 8], True, 10)
        m.train()
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            eps = m(xt, t, y)
            loss = torch.nn.functional.mse_loss(noise, eps.float())
        loss.backward()
        return float(loss), eps.detach().float(), {k: (p.grad.detach().double() if p.grad is not None else None) for k, p in m.named_parameters()}

    l32, e32, g32 = grads(False)
    l16, e16, g16 = grads(True)
    norm_err, l2_err = {}, {}
    for k, a in g32.items():
        if a is None:
            continue
        b = g16[k]
        norm_err[k] = float(abs(b.norm() - a.norm()) / a.norm())
        l2_err[k] = float((a - b).norm() / a.norm())
    out = {
        "what": "reference UNet, golden G5 inputs, CPU: bf16 autocast vs fp32",
        "loss_rel_err": abs(l16 - l32) / l32,
        "eps_rel_l2": float((e16 - e32).norm() / e32.norm()),
        "grad_norm_rel_err": {"max": max(norm_err.values()), "median": float(np.median(list(norm_err.values()))),
                              "argmax": max(norm_err, key=norm_err.get)},
        "grad_rel_l2": {"max": max(l2_err.values()), "median": float(np.median(list(l2_err.values()))),
                        "p90": float(np.percentile(list(l2_err.values()), 90)), "argmax": max(l2_err, key=l2_err.get)},
        "torch": torch.__version__,
    }
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
