"""Batch-sharded data parallelism for the train step: bucketed gradient all-reduce overlapped with the backward.

The reference is single-device; the semantics here are DistributedDataParallel's: every rank
runs forward/backward on its own images (per-rank BatchNorm statistics, per-rank batch-global Dice), gradients are
AVERAGED over ranks, every rank applies the same optimizer update.  Parameters and buffers are broadcast from rank 0
when data parallelism is enabled.

The backward of the network is cut into 4 stages whose parameter gradients are final when the stage ends
(`unetb200_grad_bucket_range`): stage 0 = decoder + head, 1 = encoder.layer4 (54 % of all parameters, finished early
because it runs at 16x16), 2 = layer3, 3 = layer2 + layer1 + stem.  After stage k the bucket k all-reduce is launched
on NCCL's stream (it waits for the compute stream's stage-k work only) and runs over NVLink while stage k+1 computes.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib


class GradBucketReducer:
    """All-reduce (average) of the flat gradient array, one contiguous bucket per backward stage."""

    def __init__(self, flat_grads: torch.Tensor, ranges=None, group=None):
        self.flat = flat_grads
        self.ranges = list(ranges) if ranges is not None else _lib.grad_bucket_ranges()
        self.group = group
        self.world = dist.get_world_size(group)
        backend = dist.get_backend(group)
        self._avg = backend == "nccl"  # gloo has no ReduceOp.AVG: sum, then scale
        self._pending = []
        # True: the backward leaves the all-reduces in flight and the optimizer waits bucket by bucket
        # (FusedAdamW(overlap_allreduce=True)): the update of buckets 0-2 (95 % of the parameters) then runs while the
        # last, smallest bucket — which can only start when the backward ends — is still being reduced.
        self.defer_finish = False
        covered = sorted(self.ranges)
        assert covered[0][0] == 0 and covered[-1][1] == flat_grads.numel() and all(
            a[1] == b[0] for a, b in zip(covered, covered[1:])), "buckets must tile the flat gradient array"

    def bucket(self, stage: int) -> torch.Tensor:
        b, e = self.ranges[stage]
        return self.flat[b:e]

    def reduce(self, stage: int):
        t = self.bucket(stage)
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        work = dist.all_reduce(t, op=op, group=self.group, async_op=True)
        self._pending.append((work, t, stage))

    def wait(self, stage: int):
        """Wait for the all-reduce of one bucket (no-op if it is not in flight)."""
        keep = []
        for work, t, st in self._pending:
            if st != stage:
                keep.append((work, t, st))
                continue
            work.wait()  # CUDA: makes the current stream wait for the collective; CPU (gloo): blocks
            if not self._avg:
                t.mul_(1.0 / self.world)
        self._pending = keep

    def finish(self):
        for stage in sorted({st for _, _, st in self._pending}):
            self.wait(stage)


def _high_priority_nccl_group():
    """A process group whose NCCL kernels run on a HIGH-PRIORITY stream: the backward's persistent one-CTA-per-SM compute
    grids keep every SM busy, so an all-reduce launched at normal priority mostly waits for SM slots and runs when the
    backward drains (round 1: the N = 8 step was 0.25 ms = one un-overlapped 98 MB all-reduce longer than N = 1).  At high
    priority its few CTAs take the first SMs any compute CTA frees.  UNETB200_DP_NORMAL_PRIORITY=1 switches it off."""
    import os
    if os.environ.get("UNETB200_DP_NORMAL_PRIORITY") or dist.get_backend() != "nccl":
        return None
    try:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        return dist.new_group(backend="nccl", pg_options=opts)
    except Exception:  # pragma: no cover  (older torch without the option)
        return None


def enable_data_parallel(model, group=None, broadcast: bool = True):
    """Turn `model` (unet_b200.Unet on this rank's GPU) into a data-parallel replica.  Returns the model.
    group=None: a dedicated NCCL group with a high-priority stream is created (collective call on every rank)."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    if group is None:
        group = _high_priority_nccl_group()
    if broadcast:
        dist.broadcast(model.flat_params, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        dist.broadcast(model.flat_buffers, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        model._params_epoch += 1
        model._buffers_epoch += 1
    model._dp = GradBucketReducer(model._grad_buffer(), group=group)
    return model


def shard_batch(n_total: int, rank: int, world: int):
    """Contiguous image range [begin, end) of rank `rank` for a global batch of n_total images (inference sharding)."""
    per, rem = divmod(n_total, world)
    b = rank * per + min(rank, rem)
    return b, b + per + (1 if rank < rem else 0)


def dp_gradient_parity(dev, rank: int, world: int, batch: int = 2, size: int = 128):
    """Multi-GPU correctness self-check (needs an initialised NCCL group, one process per GPU): the all-reduced gradients
    of ONE data-parallel step must equal the single-process emulation "run every rank's batch slice separately on one
    GPU, average the flat gradients" (DDP semantics) and must be bit-identical on every rank.
    Returns (rel_l2, max_abs, identical_across_ranks) on every rank.  Used by scripts/dp_check.py, tests/test_gpu_dp.py
    and bench.py (`train.dp_parity_rel_l2`), so that every multi-GPU bench line carries its own correctness figure."""
    from . import losses
    from .unet import Unet

    torch.manual_seed(42 + rank)  # deliberately different init per rank: enable_data_parallel must broadcast rank 0's
    model = Unet("resnet34").to(dev).train()
    enable_data_parallel(model)
    crit = losses.BCEDiceLoss()
    g = torch.Generator().manual_seed(99)
    X = torch.randn(world * batch, 3, size, size, generator=g)
    Y = (torch.rand(world * batch, 1, size, size, generator=g) < 0.1).float()
    b, e = shard_batch(world * batch, rank, world)
    p0 = model.flat_params.clone()
    buf0 = model.flat_buffers.clone()
    crit(model(X[b:e].to(dev)), Y[b:e].to(dev)).backward()
    torch.cuda.synchronize(dev)
    got = model.flat_grads.clone()
    ref0 = got.clone()
    dist.broadcast(ref0, 0)
    same = torch.tensor([int(torch.equal(ref0, got))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out = torch.zeros(2, device=dev, dtype=torch.float64)
    if rank == 0:
        single = Unet("resnet34").to(dev).train()
        acc = torch.zeros_like(got)
        for r in range(world):
            with torch.no_grad():
                single.flat_params.copy_(p0)
                single.flat_buffers.copy_(buf0)
            single._params_epoch += 1
            bb, ee = shard_batch(world * batch, r, world)
            single.zero_grad(set_to_none=True)
            crit(single(X[bb:ee].to(dev)), Y[bb:ee].to(dev)).backward()
            acc += single.flat_grads
        acc /= world
        out[0] = float((got - acc).norm() / acc.norm())
        out[1] = float((got - acc).abs().max())
    dist.broadcast(out, 0)
    return float(out[0]), float(out[1]), bool(int(same.item()))


def dp_overlap_update_parity(dev, rank: int, world: int, steps: int = 3):
    """Second multi-GPU self-check: with FusedAdamW(overlap_allreduce=True) the optimizer updates bucket by bucket while
    later all-reduces are still in flight; the parameters must be BITWISE what the joined path gives.  The gradients are
    synthetic (seeded per rank and step, written into the flat gradient array; a real backward sums weight gradients with
    fp32 atomics and AdamW turns last-bit noise of near-zero gradients into full-size steps, which would drown the
    comparison), the buckets are reduced in backward-completion order exactly as `_UnetTrainFn.backward` does.
    Returns the number of parameters that differ (max over ranks): must be 0."""
    from .optim import FusedAdamW
    from .unet import Unet

    finals = []
    for overlap in (False, True):
        torch.manual_seed(11)
        model = Unet("resnet34").to(dev).eval()
        with torch.no_grad():
            model(torch.zeros(1, 3, 64, 64, device=dev))     # creates the native context the optimizer call needs
        enable_data_parallel(model)
        dp = model._dp
        opt = FusedAdamW(model, lr=1e-3, weight_decay=1e-4, overlap_allreduce=overlap)
        dp.defer_finish = overlap
        g = model.flat_grads
        for s in range(steps):
            gen = torch.Generator(device=dev).manual_seed(1000 * s + rank)
            g.copy_(torch.randn(g.shape, device=dev, generator=gen))
            for stage in range(len(dp.ranges)):
                dp.reduce(stage)
            if not dp.defer_finish:
                dp.finish()
            opt.step()
        torch.cuda.synchronize(dev)
        finals.append(model.flat_params.clone())
    t = torch.tensor([float((finals[0] != finals[1]).sum())], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t)
