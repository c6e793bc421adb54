"""Exploration (GPU): which training recipe on the Vickers fixture gives a stable, well-trained network in few steps.
Runs on the CUDA path (8.5 ms / step); prints validation trajectories."""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import vickers_hardness_unet_b200 as vb
import vickers_data as vd
from oracle import build_oracle

data = vd.load_vickers()
import sys
BATCH = int(os.environ.get("EXPLORE_BATCH", "16"))
for lr, steps, warm in ((3e-4, 1500, 30), (3e-4, 2500, 30), (2e-4, 2500, 30), (1.5e-4, 1500, 30)):
    m = vb.Unet("resnet34")
    m.load_state_dict(build_oracle(42).state_dict(), strict=True)
    m = m.cuda()
    opt = vb.FusedAdamW(m, lr=lr, weight_decay=1e-4)
    crit = vb.losses.BCEDiceLoss()
    k = [0]

    def step(x, y):
        k[0] += 1
        s = k[0]
        f = min(1.0, s / warm) * 0.5 * (1 + math.cos(math.pi * min(1.0, s / steps)))
        opt.param_groups[0]["lr"] = lr * f
        opt.zero_grad(set_to_none=True)
        loss = crit(m(x), y)
        loss.backward()
        opt.step()
        return loss.detach()
    t0 = time.time()
    print(f"== lr {lr} steps {steps} cosine, warm-up {warm}")
    h = vd.train(m, step, data, "cuda", steps, BATCH, 1234, 250, log=print)
    print(f"   {time.time() - t0:.1f} s")
