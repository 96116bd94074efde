"""Multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous only.

Sampling shards the batch (no data-path collective): rank r owns global samples
[r*B, (r+1)*B) and keys its Philox streams by that global index, so the union of the shards is
bit-identical to a single-GPU run of the whole batch.  Training is data parallel with one gradient
all-reduce per step (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process when absent)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_offset(batch_per_rank: int, rank: int) -> int:
    """Global index of this rank's first sample (weak scaling: every rank generates batch_per_rank images)."""
    return batch_per_rank * rank


def split_batch(total: int, rank: int, world: int) -> Tuple[int, int]:
    """(offset, count) of rank's slice of `total` samples (strong scaling; remainder to the low ranks)."""
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def allreduce_mean_(flat_grad: torch.Tensor) -> torch.Tensor:
    """Data-parallel gradient exchange: one all-reduce(sum) over the flat gradient bucket, then 1/world
    (20,350,915 grads; parameters without a gradient are zero-filled by the caller)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        flat_grad.mul_(1.0 / dist.get_world_size())
    return flat_grad


def flatten_grads(params: Iterable[torch.nn.Parameter], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Pack .grad of every parameter (zeros where None, e.g. the reference's dead bottleneck mlp_t, App. D-3)."""
    params = list(params)
    n = sum(p.numel() for p in params)
    if out is None:
        out = torch.zeros(n, dtype=torch.float32, device=params[0].device)
    off = 0
    for p in params:
        k = p.numel()
        if p.grad is None:
            out[off:off + k].zero_()
        else:
            out[off:off + k].copy_(p.grad.reshape(-1))
        off += k
    return out


def unflatten_grads(params: Iterable[torch.nn.Parameter], flat: torch.Tensor, skip_none: bool = True) -> None:
    off = 0
    for p in params:
        k = p.numel()
        if p.grad is not None:
            p.grad.copy_(flat[off:off + k].view_as(p.grad))
        elif not skip_none:
            p.grad = flat[off:off + k].view_as(p).clone()
        off += k


def sync_gradients(params: Iterable[torch.nn.Parameter], bucket: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Data-parallel step between ``loss.backward()`` and ``optimizer.step()``: pack every gradient into one flat fp32
    bucket (zeros for parameters without a gradient -- the reference's four dead bottleneck mlp_t tensors -- so that all
    ranks reduce the same length), one all-reduce(sum) over NCCL / NVLink, scale by 1/world, unpack.  Returns the bucket so
    the caller can reuse it next step.  No-op on a single process."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return bucket
    params = list(params)
    bucket = flatten_grads(params, bucket)
    allreduce_mean_(bucket)
    # every parameter receives the reduced gradient, also those without a local one (a rank whose label-drop coin fell
    # differently has no label_emb gradient; Adam leaves all-zero gradients' parameters untouched, as the reference does)
    unflatten_grads(params, bucket, skip_none=False)
    return bucket
