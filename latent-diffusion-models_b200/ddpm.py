"""Drop-in ``Diffusion`` process backed by libldm_b200.so.

Same surface as the reference's ``src/DDPM.py:22-149`` -- ``Diffusion(n_steps, device, n_samples)``,
attributes ``beta/alpha/alpha_bar/sigma2/n_steps/n_samples/device``, ``q_sample``, ``p_sample``,
``sample`` (returns a CPU tensor) and ``__call__(x0, noise) -> (noise, xt, t)`` -- but every tensor
operation is one fused kernel, the reverse loop is a replayed CUDA graph with a device-resident
timestep, and nothing synchronises with the host until the final ``.cpu()``.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib


def _unwrap_eps_model(m):
    """Find the native UNet behind the reference's pass-through wrappers (LatentDiffusionModel -> DiffusionWrapper)."""
    from .unet import UNet
    seen = 0
    while m is not None and seen < 4:
        if isinstance(m, UNet):
            return m
        m = getattr(m, "model", None) or getattr(m, "diffusion_model", None)
        seen += 1
    return None


class Diffusion(nn.Module):
    def __init__(self, n_steps: int, device, n_samples: int = 1):
        super().__init__()
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.n_samples = n_samples
        # linear beta schedule; plain tensor attributes on `device`, exactly as the reference (:31-43)
        self.beta = torch.linspace(0.0001, 0.02, n_steps).to(self.device)
        self.alpha = 1.0 - self.beta
        self.alpha_bar = torch.cumprod(self.alpha, dim=0)
        self.n_steps = n_steps
        self.sigma2 = self.beta
        self._coef: Optional[torch.Tensor] = None
        self._samplers = {}
        self.last_launches = 0

    # ------------------------------------------------------------------ helpers
    def set_schedule(self, beta: torch.Tensor, alpha_bar: Optional[torch.Tensor] = None) -> None:
        """Replace the schedule (e.g. LatentDiffusionModel's sqrt-linear betas, src/LatentDiffusionModel.py:41-55)."""
        self.beta = beta.detach().to(self.device, torch.float32).contiguous()
        self.alpha = 1.0 - self.beta
        self.alpha_bar = (alpha_bar.detach().to(self.device, torch.float32).contiguous()
                          if alpha_bar is not None else torch.cumprod(self.alpha, dim=0))
        self.sigma2 = self.beta
        self.n_steps = self.beta.numel()
        self._coef = None
        self._destroy_samplers()

    def _destroy_samplers(self, dead_only: bool = False) -> None:
        """Free native samplers (all, or those whose UNet object is gone)."""
        try:
            lib = _lib.load()
        except Exception:
            return
        for key in list(self._samplers):
            ent = self._samplers[key]
            if dead_only and ent["unet"]() is not None:
                continue
            _lib.destroy_native("ldm_sampler_destroy", ent["s"])    # parked while a CUDA-graph capture is in progress
            del self._samplers[key]

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_samplers"] = {}       # native handles belong to this object only
        state["_coef"] = None
        state.pop("_abar_dev", None)
        return state

    def _require_cuda(self, t: torch.Tensor) -> None:
        if not t.is_cuda:
            raise _lib.LdmError("ldm_b200.Diffusion runs on CUDA tensors only (no CPU fallback)")

    def _coef_table(self, dev: torch.device) -> torch.Tensor:
        if self._coef is None or self._coef.device != dev:
            sched = [v.to(dev, torch.float32).contiguous() for v in (self.beta, self.alpha, self.alpha_bar)]
            coef = torch.empty(self.n_steps, 4, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(_lib.load().ldm_build_coef_table(sched[0].data_ptr(), sched[1].data_ptr(),
                                                            sched[2].data_ptr(), self.n_steps, coef.data_ptr(),
                                                            _lib.stream_ptr()))
            self._coef = coef
            self._abar_dev = sched[2]
        return self._coef

    # ------------------------------------------------------------------ q(x_t | x_0)   src/DDPM.py:46-68
    def q_xt_x0(self, x0: torch.Tensor, t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        abar = self.alpha_bar.to(x0.device).gather(-1, t).reshape(-1, 1, 1, 1)
        return abar ** 0.5 * x0, 1 - abar

    def q_sample(self, x0: torch.Tensor, t: torch.Tensor, eps: Optional[torch.Tensor] = None,
                 return_eps: bool = False):
        self._require_cuda(x0)
        dev = x0.device
        x0c = x0.detach().to(torch.float32).contiguous()
        tc = t.detach().to(dev, torch.int64).contiguous()
        B = x0c.shape[0]
        n = x0c[0].numel()
        self._coef_table(dev)
        xt = torch.empty_like(x0c)
        eps_out = None
        if eps is None:
            eps_out = torch.empty_like(x0c)
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        else:
            eps = eps.detach().to(dev, torch.float32).contiguous()
            seed = 0
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ldm_q_sample(x0c.data_ptr(), tc.data_ptr(), self._abar_dev.data_ptr(), self.n_steps,
                                                _lib.ptr(eps), _lib.ptr(eps_out), xt.data_ptr(), B, n, seed, 0,
                                                _lib.stream_ptr()))
        if return_eps:
            return xt, (eps if eps is not None else eps_out)
        return xt

    # ------------------------------------------------------------------ p_theta(x_{t-1} | x_t)   src/DDPM.py:71-96
    def p_sample(self, xt: torch.Tensor, t: torch.Tensor, eps_theta: torch.Tensor,
                 noise: Optional[torch.Tensor] = None, eps_uncond: Optional[torch.Tensor] = None,
                 cfg_scale: float = 0.0, seed: Optional[int] = None):
        """One reverse step.  ``noise`` injects z (parity); otherwise z is drawn in-kernel.
        ``eps_uncond``/``cfg_scale`` fuse the classifier-free-guidance lerp (src/DDPM.py:124)."""
        self._require_cuda(xt)
        dev = xt.device
        x = xt.detach().to(torch.float32).contiguous()
        e = eps_theta.detach().to(dev, torch.float32).contiguous()
        u = eps_uncond.detach().to(dev, torch.float32).contiguous() if eps_uncond is not None else None
        z = noise.detach().to(dev, torch.float32).contiguous() if noise is not None else None
        tc = t.detach().to(dev, torch.int64).contiguous()
        B = x.shape[0]
        if tc.numel() not in (1, B):
            raise ValueError("t must have 1 or batch entries")
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if z is None else 0
        coef = self._coef_table(dev)
        out = torch.empty_like(x)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ldm_p_sample(x.data_ptr(), e.data_ptr(), _lib.ptr(u), float(cfg_scale), tc.data_ptr(),
                                                tc.numel(), coef.data_ptr(), self.n_steps, _lib.ptr(z), seed, 0,
                                                out.data_ptr(), B, x[0].numel(), _lib.stream_ptr()))
        return out

    # ------------------------------------------------------------------ reverse process   src/DDPM.py:98-130
    @torch.no_grad()
    def sample(self, eps_model, classes, shape, device, cfg_scale=3, *, x_T: Optional[torch.Tensor] = None,
               noise: Optional[torch.Tensor] = None, seed: Optional[int] = None, sample_offset: int = 0,
               use_graph: bool = True, return_device: bool = False, first_step: Optional[int] = None,
               num_steps: Optional[int] = None):
        """Same call as the reference; the extra keyword-only arguments are for parity and sharding:
        ``x_T`` fixes the initial noise, ``noise`` ([T,B,C,S,S], indexed by t) fixes every step's z,
        ``seed``/``sample_offset`` key the in-kernel Philox streams by global sample index;
        ``first_step``/``num_steps`` run only timesteps first_step, first_step-1, ... (chunked trajectories)."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.LdmError("ldm_b200.Diffusion.sample needs a CUDA device (no CPU fallback)")
        unet = _unwrap_eps_model(eps_model)
        B, Cc, S, S2 = shape
        if unet is None:
            return self._sample_generic(eps_model, classes, shape, dev, cfg_scale, x_T, noise, seed, return_device)
        lib = _lib.load()
        with torch.cuda.device(dev):
            y = None
            y_len = 0
            if classes is not None:
                y = classes.detach().to(dev, torch.int64).contiguous()   # H2D when the caller passes host labels
                y_len = y.numel()
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            if x_T is not None:
                x = x_T.detach().to(dev, torch.float32).contiguous().clone()   # H2D of x_T when it lives on the host
                x_init = 1
            else:
                x = torch.empty(shape, dtype=torch.float32, device=dev)
                x_init = 0
            z = noise.detach().to(dev, torch.float32).contiguous() if noise is not None else None
            h = unet.native(S, dev)
            # keyed by the UNet OBJECT's identity (handle addresses and id() are reused after garbage collection)
            key = (unet._uid, h, B, float(cfg_scale), y_len, bool(use_graph))
            ent = self._samplers.get(key)
            if ent is None:
                self._destroy_samplers(dead_only=True)
                d = _lib.SamplerDesc(batch=B, n_steps=self.n_steps, cfg_scale=float(cfg_scale), y_len=y_len,
                                     use_graph=int(use_graph))
                sp = C.c_void_p()
                _lib.check(lib.ldm_sampler_create(h, C.byref(d), C.byref(sp)))
                nbytes = lib.ldm_sampler_workspace_bytes(sp.value)
                ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
                # persistent buffers so that the captured graph can be replayed across calls
                xbuf = torch.empty(shape, dtype=torch.float32, device=dev)
                ybuf = torch.zeros(max(y_len, 1), dtype=torch.int64, device=dev)
                ent = {"s": sp.value, "ws": ws, "nbytes": nbytes, "x": xbuf, "y": ybuf, "unet": weakref.ref(unet)}
                self._samplers[key] = ent
            ws = ent["ws"]
            off = (-ws.data_ptr()) % 1024
            xbuf, ybuf = ent["x"], ent["y"]
            if x_init:
                xbuf.copy_(x)
            if y is not None:
                ybuf.copy_(y)
            coef = self._coef_table(dev)
            before = _lib.launch_count()
            _lib.check(lib.ldm_sampler_run(ent["s"], xbuf.data_ptr(), x_init, ybuf.data_ptr() if y is not None else None,
                                           coef.data_ptr(), _lib.ptr(z), seed, sample_offset,
                                           self.n_steps - 1 if first_step is None else int(first_step),
                                           (self.n_steps if first_step is None else int(first_step) + 1)
                                           if num_steps is None else int(num_steps),
                                           ws.data_ptr() + off, ent["nbytes"], _lib.stream_ptr()))
            self.last_launches = _lib.launch_count() - before
            if return_device:
                return xbuf.clone()
            return xbuf.cpu()   # the reference's xt.detach().cpu() (:128): the only host sync of the call

    def _sample_generic(self, eps_model, classes, shape, dev, cfg_scale, x_T, noise, seed, return_device):
        """Duck-typed eps_model (not our UNet): Python loop, fused CFG + p_sample kernel per step."""
        xt = x_T.to(dev, torch.float32).clone() if x_T is not None else torch.randn(shape, device=dev)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        tt = torch.empty(shape[0], dtype=torch.int64, device=dev)
        for step in reversed(range(self.n_steps)):
            tt.fill_(step)
            eps = eps_model(xt, tt, classes)
            eps_u = eps_model(xt, tt, None) if cfg_scale > 0 else None
            z = noise[step] if noise is not None else None
            xt = self.p_sample(xt, tt[:1], eps, noise=z, eps_uncond=eps_u, cfg_scale=cfg_scale, seed=seed)
        return xt if return_device else xt.detach().cpu()

    # ------------------------------------------------------------------ training-time noising   src/DDPM.py:133-149
    def forward(self, x0: torch.Tensor, noise: Optional[torch.Tensor] = None):
        batch_size = x0.shape[0]
        t = torch.randint(0, self.n_steps, (batch_size,), device=x0.device, dtype=torch.long)
        xt, noise = self.q_sample(x0, t, eps=noise, return_eps=True)
        return noise, xt, t

    def __del__(self):
        try:
            self._destroy_samplers()
        except Exception:
            pass
