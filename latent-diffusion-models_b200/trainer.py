"""What the reference does either side of the denoising hot path in its trainer:
the optimizer step, the validation (CFG) loss and the image output stage, on the C ABI.

* ``FlatAdam``        -- ``torch.optim.Adam(model.parameters(), lr)`` as ``Trainer._get_optimizer`` builds it
                         (src/Trainer.py:68-71): parameters, gradients and both moments live in four flat fp32 buffers;
                         one all-reduce over the gradient buffer (data parallel) and ONE kernel launch update the model.
* ``train_step``      -- the body of ``DiffusionModelTrainer._train_epoch`` (src/DiffusionModelTrainer.py:36-67).
* ``val_step``        -- the body of ``_val_epoch`` (:79-118): cond + uncond in one 2B-row UNet pass, lerp, MSE.
* ``DiffusionModelTrainer`` -- the reference class's compute methods (forward / sample / _train_epoch / _val_epoch);
                         logging, early stopping and checkpoint files stay with the caller (out of scope).

CUDA only: there is no CPU path.
"""
from __future__ import annotations

import os
import sys
from typing import Callable, Iterable, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .ops import images_to_uint8  # noqa: F401  (re-exported: the output stage belongs to this layer)


class FlatAdam:
    """Adam with torch's defaults (betas 0.9/0.999, eps 1e-8, no weight decay, no amsgrad) over flat buffers.

    On construction every parameter's storage is moved into one flat fp32 buffer (``p.data`` becomes a view into it, so
    the module, its ``state_dict()`` and the native weight re-pack keep working); gradients and both moments have flat
    buffers of the same layout (every parameter starts on a 256-byte boundary; the padding stays zero for ever).

    The flat gradient buffer IS the gradient bucket: each parameter carries its slice as ``p._ldm_grad_slot`` and the
    backward kernels of ``ldm_b200.train`` accumulate dW / db / dgamma / dbeta straight into it (the training forward zeroes
    the bucket once, ``begin_step``), so after ``loss.backward()`` ``p.grad`` is that slice and ``step()`` has nothing to
    pack.  Gradients that arrive some other way (a torch-autograd model, the concatenated time-projection weights) are
    copied in; parameters without one (the reference's four dead bottleneck ``mlp_t`` tensors) contribute zeros, so every
    rank reduces the same bytes.  ``step()`` all-reduces the bucket in ``n_buckets`` contiguous pieces, LAST piece first
    (parameters() order is forward order, so the tail -- decoder, final conv -- is what backward finishes first), each
    piece's Adam launch (``ldm_adam_step``) overlapping the next piece's NCCL all-reduce.  The mean's 1/world_size is folded
    into the kernel's ``grad_scale``."""

    ALIGN = 64   # elements: 256 bytes (the weight-gradient kernel reduces with 16-byte vector atomics)

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 n_buckets: int = 3):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        if not self.params:
            raise ValueError("FlatAdam: no parameters")
        dev = self.params[0].device
        for p in self.params:
            if p.device != dev or not p.is_cuda or p.dtype != torch.float32:
                raise _lib.LdmError("FlatAdam: parameters must be fp32 CUDA tensors on one device (no CPU path)")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.device = dev
        sizes = [p.numel() for p in self.params]
        padded = [(k + self.ALIGN - 1) // self.ALIGN * self.ALIGN for k in sizes]
        self.offsets = np.concatenate([[0], np.cumsum(padded)]).astype(np.int64)
        n = int(self.offsets[-1])
        self.flat_param = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad_views = []
        with torch.no_grad():
            for p, o, k in zip(self.params, self.offsets[:-1], sizes):
                view = self.flat_param[int(o):int(o) + k].view_as(p)
                view.copy_(p.data)
                p.data = view
                slot = self.flat_grad[int(o):int(o) + k].view_as(p)
                self.grad_views.append(slot)
                p._ldm_grad_slot = slot          # read by ldm_b200.train's backward functions
                p._ldm_grad_bucket = self
        # contiguous pieces of about equal size, cut at parameter boundaries
        n_buckets = max(1, min(int(n_buckets), len(self.params)))
        cuts = [0]
        for b in range(1, n_buckets):
            target = n * b // n_buckets
            cuts.append(int(self.offsets[int(np.searchsorted(self.offsets, target))]))
        cuts.append(n)
        self.bucket_bounds = sorted(set(cuts))
        # data parallel: every replica starts from rank 0's parameters (the reference has one process; nothing else would make
        # the replicas agree unless every rank happened to seed torch identically)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(self.flat_param, src=0)
        self.step_count = 0
        self.generation = 0                      # bumped by begin_step: a slot is handed out once per generation
        self._slot_gen: dict = {}
        self.overlap = True                      # data parallel: start the tail's all-reduce from the backward pass
        self._pending = None                     # (offset, [(lo, hi, work)]) of the early reduction of this step

    # ---- the bucket protocol used by ldm_b200.train
    def begin_step(self) -> None:
        """Zero the gradient bucket (one memset; CUDA-graph capturable) and open a new generation of slot hand-outs."""
        self.flat_grad.zero_()
        self.generation += 1

    def take_slot(self, p: torch.Tensor) -> Optional[torch.Tensor]:
        """The zero-filled slice a backward kernel may accumulate ``p``'s gradient into, or None when in-place accumulation
        would be wrong: the bucket was not zeroed this step, the slot was already handed out (two backward passes through
        one forward), or ``p.grad`` already holds something (autograd would then ADD the returned tensor to it)."""
        if self.generation == 0 or p.grad is not None or self._slot_gen.get(id(p)) == self.generation:
            return None
        self._slot_gen[id(p)] = self.generation
        return p._ldm_grad_slot.detach()     # a fresh alias: autograd adopts a gradient only if nobody else holds the object

    # ---- data parallel: the tail of the bucket is reduced while the head is still being computed
    def wants_overlap(self) -> bool:
        return self.overlap and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def reduce_tail_async(self, first_tail_param: torch.Tensor) -> None:
        """All-reduce flat_grad[offset(first_tail_param):] now (async; NCCL's stream first waits for everything queued on the
        current stream so far).  Called by ldm_b200.train from a gradient hook on the encoder/decoder boundary: by then the
        backward kernels of every parameter from ``first_tail_param`` on have been queued.  ``step()`` reduces the rest."""
        if self._pending is not None:
            return                               # one early reduction per step
        idx = next(i for i, p in enumerate(self.params) if p is first_tail_param)
        lo = int(self.offsets[idx])
        pieces = [b for b in self.bucket_bounds if b > lo]
        works = []
        cuts = [lo] + pieces                     # (lo, b0], (b0, b1], ... up to the end of the bucket; last piece first
        for a, b in reversed(list(zip(cuts[:-1], cuts[1:]))):
            works.append((a, b, dist.all_reduce(self.flat_grad[a:b], op=dist.ReduceOp.SUM, async_op=True)))
        self._pending = (lo, works)

    # torch.optim.Optimizer surface used by the reference (src/DiffusionModelTrainer.py:55-63)
    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if set_to_none or p.grad is None:
                p.grad = None
            else:
                p.grad.zero_()

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0) -> None:
        pending, self._pending = self._pending, None
        reduced_from = pending[0] if pending is not None else int(self.offsets[-1])
        src, dst, dead = [], [], []
        for p, v, o in zip(self.params, self.grad_views, self.offsets[:-1]):
            if p.grad is None:
                if int(o) < reduced_from:
                    dead.append(v)               # (in the early-reduced tail it is the bucket's zero fill, summed over ranks)
            elif p.grad.data_ptr() != v.data_ptr():
                if int(o) >= reduced_from:
                    raise _lib.LdmError("FlatAdam: a gradient of the early-reduced bucket tail was not accumulated in place; "
                                        "set optimizer.overlap = False for this model")
                src.append(p.grad)
                dst.append(v)
        if dead:
            torch._foreach_zero_(dead)
        if dst:
            torch._foreach_copy_(dst, src)
        scale = float(grad_scale)
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.step_count += 1
        lib = _lib.load()
        with torch.cuda.device(self.device):
            def adam(lo: int, hi: int) -> None:
                _lib.check(lib.ldm_adam_step(self.flat_param.data_ptr() + 4 * lo, self.flat_grad.data_ptr() + 4 * lo,
                                             self.exp_avg.data_ptr() + 4 * lo, self.exp_avg_sq.data_ptr() + 4 * lo, hi - lo,
                                             self.lr, self.betas[0], self.betas[1], self.eps, self.step_count,
                                             scale / world, _lib.stream_ptr()))
            if world > 1:
                # NCCL over NVLink.  The head of the bucket (what the early reduction did not cover; everything without one)
                # goes now, last piece first; the Adam launch of a piece runs under the all-reduce of the next one.
                cuts = [0] + [b for b in self.bucket_bounds if 0 < b < reduced_from] + [reduced_from]
                works = [(lo, hi, dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
                         for lo, hi in reversed(list(zip(cuts[:-1], cuts[1:]))) if hi > lo]
                for lo, hi, w in (pending[1] if pending is not None else []) + works:
                    w.wait()                 # stream-level wait on CUDA: the host does not block
                    adam(lo, hi)
            else:
                adam(0, int(self.offsets[-1]))
        torch.autograd.graph.increment_version(self.params)   # the kernel wrote through raw pointers: tell torch (and the
        #                                                       UNet's weight re-pack, which keys on ._version)

    def state_dict(self) -> dict:
        return {"step": self.step_count, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "lr": self.lr, "betas": self.betas, "eps": self.eps}

    def load_state_dict(self, sd: dict) -> None:
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.lr, self.betas, self.eps = float(sd["lr"]), tuple(sd["betas"]), float(sd["eps"])


def mse_loss(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """F.mse_loss(a, b) for fp32 CUDA tensors without gradients -> 0-dim device tensor (ldm_mse)."""
    a = a.detach().to(torch.float32).contiguous()
    b = b.detach().to(torch.float32).contiguous()
    if not a.is_cuda or a.shape != b.shape:
        raise _lib.LdmError("mse_loss: two CUDA tensors of one shape expected (no CPU path)")
    out = torch.empty((), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().ldm_mse(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _lib.stream_ptr()))
    return out


class _MSELoss(torch.autograd.Function):
    """``F.mse_loss(input, target)`` (mean) with both directions on the C ABI: ``ldm_mse`` forward, ``ldm_mse_backward``
    for whichever side needs a gradient -- the training loss of src/DiffusionModelTrainer.py:48 / src/Trainer.py:60."""

    @staticmethod
    def forward(ctx, a, b):
        a32, b32 = a.detach().to(torch.float32).contiguous(), b.detach().to(torch.float32).contiguous()
        ctx.save_for_backward(a32, b32)
        ctx.dtypes = (a.dtype, b.dtype)
        out = torch.empty((), dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device):
            _lib.check(_lib.load().ldm_mse(a32.data_ptr(), b32.data_ptr(), out.data_ptr(), a32.numel(), _lib.stream_ptr()))
        return out

    @staticmethod
    def backward(ctx, gout):
        a, b = ctx.saved_tensors
        g = gout.detach().to(torch.float32).contiguous()
        lib = _lib.load()
        grads = [None, None]
        with torch.cuda.device(a.device):
            for i, (pred, target) in enumerate(((a, b), (b, a))):
                if ctx.needs_input_grad[i]:
                    d = torch.empty_like(pred)
                    _lib.check(lib.ldm_mse_backward(pred.data_ptr(), target.data_ptr(), g.data_ptr(), d.data_ptr(), pred.numel(),
                                                    _lib.stream_ptr()))
                    grads[i] = d.to(ctx.dtypes[i])
        return tuple(grads)


def mse_loss_autograd(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Differentiable ``F.mse_loss(a, b)`` for CUDA tensors of one shape (no CPU path): the default training loss."""
    if not a.is_cuda or a.shape != b.shape:
        raise _lib.LdmError("mse_loss_autograd: two CUDA tensors of one shape expected (no CPU path)")
    return _MSELoss.apply(a, b)


def train_step(model, diffusion, optimizer, data: torch.Tensor, targets: Optional[torch.Tensor], *, drop_labels: bool = False,
               forward: Optional[Callable] = None, loss_fn: Callable = mse_loss_autograd) -> torch.Tensor:
    """One iteration of ``_train_epoch`` (src/DiffusionModelTrainer.py:36-67): noise, forward, MSE, backward, optimizer.
    ``drop_labels`` is the reference's 10 % coin (:44), drawn by the caller so that all ranks agree.  ``forward`` may be a
    CUDA-graphed callable from ``ldm_b200.train.make_graphed``.  Returns the (device) loss; the caller decides when to sync."""
    noise, xt, t = diffusion(data)
    y = None if drop_labels else targets
    fwd = forward if forward is not None else model
    eps_theta = fwd(xt, t, y) if y is not None else fwd(xt, t)
    loss = loss_fn(noise, eps_theta)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    optimizer.step()
    return loss.detach()


@torch.no_grad()
def val_step(model, diffusion, data: torch.Tensor, targets: Optional[torch.Tensor], cfg_scale: float) -> torch.Tensor:
    """One iteration of ``_val_epoch`` (:94-107): eps(x_t, t, y), and when cfg_scale > 0 also eps(x_t, t, None) and
    lerp(eps_u, eps_c, cfg_scale); the two predictions come from ONE UNet pass over 2B rows (first B conditional)."""
    noise, xt, t = diffusion(data)
    B = xt.shape[0]
    if cfg_scale > 0 and targets is not None:
        both = model._forward_nograd(torch.cat((xt, xt)), torch.cat((t, t)), targets, y_rows=B)
        eps = torch.lerp(both[B:], both[:B], float(cfg_scale))
    else:   # cfg_scale == 0, or no labels (then the reference's two passes coincide and its lerp is the identity)
        eps = model._forward_nograd(xt, t, targets)
    return mse_loss(noise, eps)


class NoOpGradScaler:
    """What ``torch.cuda.amp.GradScaler`` is to the reference's fp16 autocast (src/Trainer.py:43,
    src/DiffusionModelTrainer.py:57-60) when the reduced-precision kernels are bf16: bf16 has fp32's exponent range, so
    there is nothing to scale -- the object keeps the call protocol (scale / step / update) and changes no number."""

    def scale(self, loss):
        return loss

    def unscale_(self, optimizer) -> None:
        return None

    def step(self, optimizer, *args, **kwargs):
        return optimizer.step(*args, **kwargs)

    def update(self, new_scale=None) -> None:
        return None

    def get_scale(self) -> float:
        return 1.0

    def state_dict(self) -> dict:
        return {}

    def load_state_dict(self, sd) -> None:
        return None


class DiffusionModelTrainer:
    """Compute methods of ``src.DiffusionModelTrainer.DiffusionModelTrainer`` on the native path, with the reference's
    constructor order ``(config, model, train_loader, val_loader, classes, diffusion, cfg_scale)`` (:14-17).

    ``config`` needs ``lr``, ``epochs`` and ``data.image_channels`` / ``data.image_size`` (mapping or attribute access).
    ``config["use_amp"]`` is honoured as the reference means it (src/Trainer.py:43, :40): True -> reduced-precision kernels
    (bf16 here, with a no-op scaler object in ``self.scaler``), False -> the fp32 kernels and ``self.scaler is None``; when the
    key is absent the model keeps the dtype it was built with.
    The reference's wandb logging, early stopping and checkpoint writing are the caller's business."""

    def __init__(self, config, model, train_loader=None, val_loader=None, classes=None, diffusion=None, cfg_scale: float = 0.0,
                 *, device=None, rng: Optional[np.random.Generator] = None, use_cuda_graphs: bool = True):
        if diffusion is None:
            raise TypeError("DiffusionModelTrainer: `diffusion` is required (reference order: config, model, train_loader, "
                            "val_loader, classes, diffusion, cfg_scale)")
        get = (lambda k: config[k]) if hasattr(config, "__getitem__") else (lambda k: getattr(config, k))
        self.config, self.model, self.diffusion = config, model, diffusion
        self.train_loader, self.val_loader = train_loader, val_loader
        self.classes, self.cfg_scale = classes, cfg_scale
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.epochs = int(get("epochs")) if self._has(config, "epochs") else 1
        self.scaler = None
        if self._has(config, "use_amp"):
            self.use_amp = bool(get("use_amp"))
            want = "bf16" if self.use_amp else "fp32"
            if hasattr(model, "set_compute_dtype"):
                model.set_compute_dtype(want)       # native handles are rebuilt lazily in the new precision
            self.scaler = NoOpGradScaler() if self.use_amp else None
        else:
            self.use_amp = getattr(model, "compute_dtype", "bf16") == "bf16"
        self.optimizer = FlatAdam(model.parameters(), lr=float(get("lr")))
        self.loss_fn = mse_loss_autograd          # F.mse_loss (src/Trainer.py:60) as ldm_mse / ldm_mse_backward
        # the label-drop coin (:44) must fall the same way on every rank of a data-parallel job (a rank that drops its labels
        # has no label_emb gradient): one seed for all ranks unless the caller passes a generator
        self._rng = rng if rng is not None else np.random.default_rng(0x5EED)
        self._get = get
        # forward + backward of the UNet captured as CUDA graphs, one pair per (batch shape, labels given?): a step is
        # ~560 kernel launches, and at the reference's batch 64 issuing them from Python takes twice as long as running them
        self.use_cuda_graphs = use_cuda_graphs
        self._graphed: dict = {}
        self.graph_capture_error: Optional[str] = None

    @staticmethod
    def _has(config, key) -> bool:
        try:
            return key in config
        except TypeError:
            return hasattr(config, key)

    def forward(self, x, t, targets=None):                      # src/DiffusionModelTrainer.py:151-160
        return self.model(x, t) if targets is None else self.model(x, t, targets)

    def _graphed_forward(self, data: torch.Tensor, targets: Optional[torch.Tensor]):
        """The CUDA-graphed UNet callable for this batch shape (captured on first use), or None to run eagerly."""
        if not self.use_cuda_graphs:
            return None
        key = (tuple(data.shape), targets is not None)
        if key not in self._graphed:
            from .train import make_graphed
            try:
                noise, xt, t = self.diffusion(data)
                self._graphed[key] = make_graphed(self.model, xt, t, targets)
            except Exception as e:   # noqa: BLE001 -- capture is an optimisation; the eager autograd path is always there
                if os.environ.get("LDM_STRICT_GRAPHS"):
                    raise
                self.graph_capture_error = f"{type(e).__name__}: {e}"
                sys.stderr.write(f"ldm_b200: CUDA-graph capture of the training step failed ({self.graph_capture_error}); "
                                 "running eagerly (about 2x slower at batch 64). LDM_STRICT_GRAPHS=1 re-raises.\n")
                self._graphed[key] = None
        return self._graphed[key]

    def _train_epoch(self, epoch: int) -> float:                 # :28-77
        self.model.train()
        total = torch.zeros((), dtype=torch.float64, device=self.device)
        for data, targets in self.train_loader:
            data, targets = data.to(self.device, non_blocking=True), targets.to(self.device, non_blocking=True)
            drop = bool(self._rng.random() < 0.1)
            loss = train_step(self.model, self.diffusion, self.optimizer, data, targets, drop_labels=drop,
                              forward=self._graphed_forward(data, None if drop else targets), loss_fn=self.loss_fn)
            total += loss.double() * data.size(0)                # accumulated on the device: one sync per epoch, not per step
        return float(total.item()) / len(self.train_loader)

    def _val_epoch(self, epoch: int) -> float:                   # :79-118
        self.model.eval()
        total = torch.zeros((), dtype=torch.float64, device=self.device)
        for data, targets in self.val_loader:
            data, targets = data.to(self.device, non_blocking=True), targets.to(self.device, non_blocking=True)
            total += val_step(self.model, self.diffusion, data, targets, self.cfg_scale).double() * data.size(0)
        return float(total.item()) / len(self.val_loader)

    def sample(self, classes, cfg_scale=0, as_uint8: bool = False):   # :162-180
        data = self._get("data")
        dget = (lambda k: data[k]) if hasattr(data, "__getitem__") else (lambda k: getattr(data, k))
        shape = (len(classes), int(dget("image_channels")), int(dget("image_size")), int(dget("image_size")))
        x = self.diffusion.sample(self.model, classes, shape=shape, device=self.device, cfg_scale=cfg_scale,
                                  return_device=as_uint8)
        if as_uint8:   # the reverse transform (src/transforms.py:22-35) on the device; HWC uint8 arrays, ready for PIL
            return list(images_to_uint8(x, "reverse_transform").cpu().numpy())
        return x

    def to(self, device):                                        # :182-185
        self.device = torch.device(device)
        self.model.to(self.device)
        self.diffusion.to(self.device)
