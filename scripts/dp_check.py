"""2+ rank data-parallel parity (torchrun, one process per GPU): the all-reduced gradients of a DP step must equal the
single-process emulation "run every rank's batch slice separately, average the flat gradients" (SURVEY.md section 8e)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import vickers_hardness_unet_b200 as vb


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, S = 2, 128
    torch.manual_seed(42 + rank)  # deliberately different init per rank: enable_data_parallel must broadcast rank 0's
    model = vb.Unet("resnet34").to(dev).train()
    vb.distributed.enable_data_parallel(model)
    crit = vb.losses.BCEDiceLoss()
    g = torch.Generator().manual_seed(99)
    X = torch.randn(world * B, 3, S, S, generator=g)
    Y = (torch.rand(world * B, 1, S, S, generator=g) < 0.1).float()
    b, e = vb.distributed.shard_batch(world * B, rank, world)
    p0 = model.flat_params.clone()
    buf0 = model.flat_buffers.clone()
    loss = crit(model(X[b:e].to(dev)), Y[b:e].to(dev))
    loss.backward()
    torch.cuda.synchronize()
    got = model.flat_grads.clone()
    # every rank must hold identical reduced gradients
    ref0 = got.clone()
    dist.broadcast(ref0, 0)
    same = bool(torch.equal(ref0, got))
    ok = True
    if rank == 0:
        single = vb.Unet("resnet34").to(dev).train()
        acc = torch.zeros_like(got)
        for r in range(world):
            with torch.no_grad():
                single.flat_params.copy_(p0)
                single.flat_buffers.copy_(buf0)
            single._params_epoch += 1
            bb, ee = vb.distributed.shard_batch(world * B, r, world)
            single.zero_grad(set_to_none=True)
            crit(single(X[bb:ee].to(dev)), Y[bb:ee].to(dev)).backward()
            acc += single.flat_grads
        acc /= world
        rel = float((got - acc).norm() / acc.norm())
        mx = float((got - acc).abs().max())
        print(f"dp_check world={world}: reduced-vs-emulated gradient rel-L2 {rel:.3e} max-abs {mx:.3e}; "
              f"identical across ranks: {same}")
        ok = rel < 1e-3
    flags = torch.tensor([int(ok and same)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flags.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("DP_CHECK_OK")


if __name__ == "__main__":
    main()
