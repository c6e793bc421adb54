"""BASELINE.json configs[3] (1024x1024, batch 8/GPU: inference + train step) and configs[4] (512x512 inference batch sweep
1..512 per GPU): throughput per configuration (CUDA events, inputs resident in HBM), one JSON line each.
On one GPU: `python scripts/configs_sweep.py`; across N GPUs, batch-sharded (one process per GPU, no collective for
inference, bucketed NCCL gradient all-reduce for the train step): `torchrun --nproc-per-node N scripts/configs_sweep.py`
— every rank runs the per-GPU batch, the time is the MAX over ranks, images/s the aggregate of all ranks.
These are the parity-test shapes of the other configs run at their full sizes; bench.py keeps configs[1]/[2]."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vickers_hardness_unet_b200 as vb


RANK = int(os.environ.get("RANK", "0"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))


def timed(fn, steps, warmup=3):
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    if WORLD > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if WORLD > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms


def emit(rec):
    rec["n_gpus"] = WORLD
    if RANK == 0:
        print(json.dumps(rec), flush=True)


def main():
    torch.cuda.set_device(LOCAL)
    dev = torch.device("cuda", LOCAL)
    if WORLD > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(42)
    model = vb.Unet("resnet34").to(dev)
    out = []
    # ---- configs[4]: batch sweep at 512^2 (inference)
    model.eval()
    for B in [1, 2, 4, 8, 16, 32, 64, 128, 256, 512]:
        x = torch.randn(B, 3, 512, 512, device=dev)
        with torch.no_grad():
            ms = timed(lambda: model.predict_mask(x, 0.5), steps=max(3, min(20, 256 // B)))
        out.append({"config": "infer 512x512", "batch_per_gpu": B, "ms": ms, "images_per_s": WORLD * B / ms * 1e3})
        emit(out[-1])
        del x
        torch.cuda.empty_cache()
    # ---- configs[3]: 1024^2, batch 8
    x = torch.randn(8, 3, 1024, 1024, device=dev)
    with torch.no_grad():
        ms = timed(lambda: model.predict_mask(x, 0.5), steps=10)
    out.append({"config": "infer 1024x1024", "batch_per_gpu": 8, "ms": ms, "images_per_s": WORLD * 8 / ms * 1e3})
    emit(out[-1])
    model.train()
    if WORLD > 1:
        vb.distributed.enable_data_parallel(model)
    y = (torch.rand(8, 1, 1024, 1024, device=dev) < 0.1).float()
    crit = vb.losses.BCEDiceLoss()
    opt = vb.optim.FusedAdamW(model, lr=1e-4, weight_decay=1e-4)
    losses = []

    def step():
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), y)
        loss.backward()
        opt.step()
        losses.append(loss)

    ms = timed(step, steps=8)
    out.append({"config": "train 1024x1024", "batch_per_gpu": 8, "ms": ms, "images_per_s": WORLD * 8 / ms * 1e3,
                "loss_first": float(losses[0]), "loss_last": float(losses[-1])})
    emit(out[-1])
    assert all(torch.isfinite(l) for l in losses) and float(losses[-1]) < float(losses[0])
    assert model._ctx.device_error_flag() == 0
    if WORLD > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if RANK == 0:
        print("CONFIGS_SWEEP_OK")


if __name__ == "__main__":
    main()
