"""BASELINE.json configs[3] (1024x1024, batch 8/GPU: inference + train step) and configs[4] (512x512 inference batch sweep
1..512) on one GPU: throughput per configuration (CUDA events, inputs resident in HBM), one JSON line each.
These are the parity-test shapes of the other configs run at their full sizes; bench.py keeps configs[1]/[2]."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vickers_hardness_unet_b200 as vb


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    model = vb.Unet("resnet34").to(dev)
    out = []
    # ---- configs[4]: batch sweep at 512^2 (inference)
    model.eval()
    for B in [1, 2, 4, 8, 16, 32, 64, 128, 256, 512]:
        x = torch.randn(B, 3, 512, 512, device=dev)
        with torch.no_grad():
            ms = timed(lambda: model.predict_mask(x, 0.5), steps=max(3, min(20, 256 // B)))
        out.append({"config": "infer 512x512", "batch": B, "ms": ms, "images_per_s": B / ms * 1e3})
        print(json.dumps(out[-1]), flush=True)
        del x
        torch.cuda.empty_cache()
    # ---- configs[3]: 1024^2, batch 8
    x = torch.randn(8, 3, 1024, 1024, device=dev)
    with torch.no_grad():
        ms = timed(lambda: model.predict_mask(x, 0.5), steps=10)
    out.append({"config": "infer 1024x1024", "batch": 8, "ms": ms, "images_per_s": 8 / ms * 1e3})
    print(json.dumps(out[-1]), flush=True)
    model.train()
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
    out.append({"config": "train 1024x1024", "batch": 8, "ms": ms, "images_per_s": 8 / ms * 1e3,
                "loss_first": float(losses[0]), "loss_last": float(losses[-1])})
    print(json.dumps(out[-1]), flush=True)
    assert all(torch.isfinite(l) for l in losses) and float(losses[-1]) < float(losses[0])
    assert model._ctx.device_error_flag() == 0
    print("CONFIGS_SWEEP_OK")


if __name__ == "__main__":
    main()
