#!/usr/bin/env python
"""Golden vectors for the steps either side of the hot path (test infrastructure; run in the build container):

    python oracle/make_golden_next.py [--ref /root/reference] [--out tests/golden]

g7_adam.npz      torch.optim.Adam(params, lr=5e-4) exactly as src/Trainer.py:68-71 constructs it, 6 steps, 3 tensors
g8_output.npz    the reference's reverse transform (src/transforms.py:22-35,58-66) and save_images (src/utils.py:121-130,
                 PNG bytes decoded again) on images that leave [-1, 1]
g9_val_loss.npz  one `_val_epoch` batch (src/DiffusionModelTrainer.py:94-107) with the reference UNet / Diffusion, cfg 3 and 0
g10_autoencoder_*.npz  src.Autoencoder.Autoencoder encode (mu, log_var, epsilon, sample), decode and forward, two configurations
g11_ldm_latent.npz     LatentDiffusionModel (src/LatentDiffusionModel.py): schedule, scaled encode, eps-prediction on the latent
                       shape, 3 reverse steps with the model's own schedule, decode (through self.autoencoder: the
                       reference's autoencoder_decode raises AttributeError on `first_stage_model`, :72)
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "golden"))
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    os.environ.setdefault("WANDB_MODE", "disabled")
    np.Inf = np.inf
    from PIL import Image
    from src.UNet import UNet
    from src.DDPM import Diffusion
    from src.transforms import get_reverse_image_transform, reverse_transform
    from src.utils import save_images

    save = lambda name, **kw: np.savez_compressed(os.path.join(args.out, name), **{
        k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in kw.items()})

    # ---------------- G7: Adam ----------------
    g = torch.Generator().manual_seed(11)
    shapes = [(64, 3, 3, 3), (257,), (5, 130)]
    params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    p0 = [p.detach().clone() for p in params]
    opt = torch.optim.Adam(params, lr=5e-4)                       # src/Trainer.py:69
    grads = [[torch.randn(s, generator=g) * (10.0 ** (i - 3)) for s in shapes] for i in range(6)]
    for gs in grads:
        opt.zero_grad(set_to_none=True)
        for p, gg in zip(params, gs):
            p.grad = gg.clone()
        opt.step()
    kw = {}
    for j in range(len(shapes)):
        kw[f"p0_{j}"] = p0[j]
        kw[f"p_final_{j}"] = params[j].detach()
        kw[f"exp_avg_{j}"] = opt.state[params[j]]["exp_avg"]
        kw[f"exp_avg_sq_{j}"] = opt.state[params[j]]["exp_avg_sq"]
        kw[f"grads_{j}"] = torch.stack([gs[j] for gs in grads])
    save("g7_adam.npz", lr=5e-4, steps=6, **kw)
    print("G7 ok")

    # ---------------- G8: output stage ----------------
    g = torch.Generator().manual_seed(12)
    x = torch.randn(5, 3, 32, 32, generator=g) * 0.9           # leaves [-1,1]: exercises the clamp and the wrap
    x[0, 0, 0, :8] = torch.tensor([-1.0, 1.0, 0.0, 1.004, -1.004, 3.2, -2.9, 0.99999])
    tr = get_reverse_image_transform()
    rev = np.stack([np.array(reverse_transform(img, transform=tr)) for img in x])     # PIL -> HWC uint8
    with tempfile.TemporaryDirectory() as d:
        cwd = os.getcwd()
        os.chdir(d)
        try:
            save_images(x, "sample")                                                   # ./sample_{i}.png
            sav = np.stack([np.array(Image.open(f"sample_{i}.png").convert("RGB")) for i in range(x.shape[0])])
        finally:
            os.chdir(cwd)
    x1 = torch.randn(3, 1, 32, 32, generator=g) * 0.9
    rev1 = np.stack([np.array(reverse_transform(img, transform=tr)) for img in x1])  # [1,H,W] -> HxWx1 -> mode L
    save("g8_output.npz", x=x, reverse_transform=rev, save_image=sav, x_gray=x1, reverse_transform_gray=rev1)
    print("G8 ok", rev.shape, sav.shape, rev1.shape)

    # ---------------- G9: validation loss ----------------
    torch.manual_seed(0)
    model = UNet(3, 3, 64, [1, 2, 4, 8], True, 10).eval()
    diffusion = Diffusion(1000, "cpu")
    g = torch.Generator().manual_seed(13)
    B = 4
    x0 = torch.rand(B, 3, 32, 32, generator=g) * 2 - 1
    noise = torch.randn(B, 3, 32, 32, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    y = torch.randint(0, 10, (B,), generator=g)
    with torch.inference_mode():
        xt = diffusion.q_sample(x0, t, noise)
        eps_c = model(xt, t, y)
        eps_u = model(xt, t, None)
        loss_cfg3 = torch.nn.functional.mse_loss(noise, torch.lerp(eps_u, eps_c, 3.0))   # :100-105
        loss_cfg0 = torch.nn.functional.mse_loss(noise, eps_c)
    save("g9_val_loss.npz", x0=x0, noise=noise, t=t, y=y, loss_cfg3=loss_cfg3, loss_cfg0=loss_cfg0, weight_seed=0)
    print("G9 ok", float(loss_cfg3), float(loss_cfg0))

    # ---------------- G10: autoencoder ----------------
    import hashlib
    from src.Autoencoder import Autoencoder
    from src.LatentDiffusionModel import LatentDiffusionModel

    def sha(sd):
        h = hashlib.sha256()
        for k in sd:
            h.update(k.encode())
            h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
        return np.frombuffer(h.digest(), dtype=np.uint8)

    for tag, cfgv, cin in (("ldm", (3, 4, 3, 64, [1, 2], 2), 3), ("deep", (1, 8, 1, 64, [1, 2, 4], 1), 1)):
        torch.manual_seed(21)
        ae = Autoencoder(*cfgv).eval()
        g = torch.Generator().manual_seed(22)
        img = torch.rand(2, cin, 32, 32, generator=g) * 2 - 1
        with torch.no_grad():
            torch.manual_seed(23)
            dist = ae.encode(img)
            zs = dist.sample()
            rec = ae.decode(zs)
            torch.manual_seed(23)
            fwd_img, fwd_mu, fwd_lv = ae(img)
        assert torch.equal(fwd_mu, dist.mu)
        save(f"g10_autoencoder_{tag}.npz", img=img, mu=dist.mu, log_var=dist.log_var, epsilon=dist.epsilon, z=zs, recon=rec,
             forward_img=fwd_img, weight_seed=21, weight_sha256=sha(ae.state_dict()), config=np.array(
                 [cfgv[0], cfgv[1], cfgv[2], cfgv[3], cfgv[5]] + list(cfgv[4])))
        print("G10", tag, tuple(zs.shape), float(rec.std()))

    # ---------------- G11: latent diffusion model ----------------
    torch.manual_seed(31)
    unet = UNet(4, 4, 64, [1, 2, 4, 8], True, 10).eval()
    torch.manual_seed(21)
    ae = Autoencoder(3, 4, 3, 64, [1, 2], 2).eval()
    ldm = LatentDiffusionModel(unet, ae, 0.18215, 1000, 0.00085, 0.012).eval()
    g = torch.Generator().manual_seed(32)
    img = torch.rand(2, 3, 32, 32, generator=g) * 2 - 1
    y = torch.tensor([3, 7])
    with torch.no_grad():
        torch.manual_seed(33)
        z0 = ldm.autoencoder_encode(img)
        t = torch.tensor([999, 500])
        eps_pred = ldm(z0, t, y)
        # three reverse steps with the LDM's own schedule (DDPM.p_sample arithmetic, src/DDPM.py:71-96) from z0 as x_T
        diff = Diffusion(1000, "cpu")
        diff.beta = ldm.beta.data.clone()
        diff.alpha = 1.0 - diff.beta
        diff.alpha_bar = ldm.alpha_bar.data.clone()
        diff.sigma2 = diff.beta
        x = z0.clone()
        zn = torch.randn(3, *z0.shape, generator=g)
        for i, step in enumerate((999, 998, 997)):
            tt = torch.full((2,), step, dtype=torch.long)
            e_c, e_u = ldm(x, tt, y), ldm(x, tt, None)
            eps = torch.lerp(e_u, e_c, 3.0)
            ab, al = diff.alpha_bar[step], diff.alpha[step]
            mean = (x - (1 - al) / (1 - ab) ** 0.5 * eps) / al ** 0.5
            x = mean + diff.sigma2[step] ** 0.5 * zn[i]
        dec = ldm.autoencoder.decode(x / ldm.latent_scaling_factor)
    torch.manual_seed(33)
    with torch.no_grad():
        eps_used = ae.encode(img).epsilon
    save("g11_ldm_latent.npz", img=img, y=y, t=t, z0=z0, encode_epsilon=eps_used, eps_pred=eps_pred, step_noise=zn, x_after3=x,
         decoded=dec, beta=ldm.beta.data, alpha_bar=ldm.alpha_bar.data, unet_seed=31, ae_seed=21)
    print("G11", float(eps_pred.std()), float(dec.std()))


if __name__ == "__main__":
    main()
