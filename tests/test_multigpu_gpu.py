"""GPU, two ranks over NCCL (run with `gpurun --gpus 2`; skipped on a one-GPU box): the two ways the path is partitioned
across GPUs give the single-GPU result.

* training is data parallel: gradients of the two half batches, all-reduced through FlatAdam's flat bucket, equal the
  single-GPU gradients of the concatenated batch (src/DiffusionModelTrainer.py:36-63 on one device);
* sampling is batch-sharded without any collective: the union of the two ranks' images equals the single-GPU batch BIT
  FOR BIT (Philox streams keyed by global sample index, statistics summed in a fixed order).
"""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT, rel_l2

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import ldm_b200
    import oracle
    from ldm_b200 import dist as D, trainer
    D.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    sd = oracle.init_state_dict(0, 3, 3, 64, (1, 2, 4, 8), True, 10)

    def model():
        m = ldm_b200.UNet(3, 3, 64, (1, 2, 4, 8), True, 10, dtype="bf16").to(dev)
        m.load_state_dict(sd)
        return m

    gen = torch.Generator().manual_seed(3)
    B = 8
    xt = torch.randn(B, 3, 32, 32, generator=gen)
    noise = torch.randn(B, 3, 32, 32, generator=gen)
    t = torch.randint(0, 1000, (B,), generator=gen)
    y = torch.randint(0, 10, (B,), generator=gen)
    # ---- data-parallel gradients through FlatAdam's bucket (lr = 0: the step only exchanges and scales)
    m = model()
    opt = trainer.FlatAdam(m.parameters(), lr=0.0)
    off, cnt = D.split_batch(B, rank, world)
    sl = slice(off, off + cnt)
    loss = torch.nn.functional.mse_loss(noise[sl].to(dev), m(xt[sl].to(dev), t[sl].to(dev), y[sl].to(dev)))
    loss.backward()
    opt.step()                                   # the bucket's tail was all-reduced from inside backward, the head goes now
    dp = torch.cat([v.reshape(-1) for v in opt.grad_views]) / world     # (the bucket itself is padded per parameter)
    out = {"rank": rank, "early_reduce_used": True}
    # ... and the same through the CUDA-graphed halves (what bench.py's training leg and the trainer run)
    from ldm_b200.train import make_graphed
    mg = model()
    optg = trainer.FlatAdam(mg.parameters(), lr=0.0)
    fwd = make_graphed(mg, torch.randn_like(xt[sl]).to(dev), t[sl].to(dev), y[sl].to(dev))
    for _ in range(2):
        lossg = trainer.mse_loss_autograd(noise[sl].to(dev), fwd(xt[sl].to(dev), t[sl].to(dev), y[sl].to(dev)))
        optg.zero_grad(set_to_none=True)
        lossg.backward()
        had_pending = optg._pending is not None
        optg.step()
    dpg = torch.cat([v.reshape(-1) for v in optg.grad_views]) / world
    out["early_reduce_used"] = had_pending
    out["graphed_vs_eager_dp"] = rel_l2(dpg, dp)
    if rank == 0:
        m1 = model()
        torch.nn.functional.mse_loss(noise.to(dev), m1(xt.to(dev), t.to(dev), y.to(dev))).backward()
        single = D.flatten_grads(list(m1.parameters()))
        out["grad_rel_l2"] = rel_l2(dp, single)
        worst = 0.0
        o = 0
        for p in m1.parameters():
            k = p.numel()
            if p.grad is not None:
                worst = max(worst, rel_l2(dp[o:o + k], single[o:o + k]))
            o += k
        out["grad_worst_param"] = worst
    # ---- batch-sharded sampling, no collective on the data path
    d = ldm_b200.Diffusion(1000, dev)
    ms = model()
    part = d.sample(ms, torch.tensor([3]), (cnt, 3, 32, 32), dev, cfg_scale=3, seed=77, sample_offset=off, first_step=999,
                    num_steps=4, return_device=True)
    parts = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(parts, part)                 # test plumbing only: bring the shards together for the comparison
    if rank == 0:
        whole = d.sample(ms, torch.tensor([3]), (B, 3, 32, 32), dev, cfg_scale=3, seed=77, first_step=999, num_steps=4,
                         return_device=True)
        out["sampling_bitwise"] = bool(torch.equal(torch.cat(parts), whole))
    dist.barrier()
    q.put(out)
    dist.destroy_process_group()


def test_two_gpu_dp_gradients_and_sharded_sampling():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + os.getpid() % 100
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    r0 = [r for r in res if r["rank"] == 0][0]
    print("two-GPU check:", r0)
    assert r0["sampling_bitwise"], "2-shard sampling must equal the 1-GPU batch bit for bit"
    # the half batches take different split-K partitions in the weight-gradient kernels: fp32 summation order only
    assert r0["grad_rel_l2"] < 2e-3 and r0["grad_worst_param"] < 2e-2, r0
    assert r0["early_reduce_used"], "the bucket tail must have been all-reduced from the backward hook"
    assert r0["graphed_vs_eager_dp"] < 2e-3, r0
