"""Test infrastructure: the reference's own micrographs (tests/golden/vickers_512.npz, built by
tests/golden/make_vickers_fixture.py from /root/reference/data) and a small training harness that follows
/root/reference/train.py — used to put the parity tests in the regime BASELINE.json's tolerance was written for: a
TRAINED network on real indentation images.

Pre-processing = the deterministic part of VickersDataset (train.py:70-75 / 116-126): LongestMaxSize(512) is already in
the fixture; here PadIfNeeded(512, 512, constant 0, centred), BGR->RGB, Normalize(ImageNet mean / std)
(train.py:108-112), masks > 0 -> 1.0 (train.py:166-173).  The random augmentation is reduced to the dihedral group
(flips / 90-degree rotations: train.py:82-86), drawn from a seeded generator so that two training runs see the same
batches.  Split: train.py:560-565 (seed 42, first 10 % = validation), stored in the fixture.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn.functional as F

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)
SIZE = 512


def load_vickers(size: int = SIZE):
    """-> dict(train_u8 [Nt,S,S,3] uint8 BGR, train_y [Nt,1,S,S] float, val_u8, val_y, names)."""
    import cv2

    z = np.load(os.path.join(GOLD, "vickers_512.npz"))
    imgs, masks = [], []
    for i in range(len(z["names"])):
        im = cv2.imdecode(z["jpeg"][z["jpeg_off"][i]:z["jpeg_off"][i + 1]], cv2.IMREAD_COLOR)
        mk = cv2.imdecode(z["mask_png"][z["mask_off"][i]:z["mask_off"][i + 1]], cv2.IMREAD_UNCHANGED)
        h, w = im.shape[:2]
        top, left = (size - h) // 2, (size - w) // 2          # A.PadIfNeeded default position: centre
        canvas = np.zeros((size, size, 3), np.uint8)
        canvas[top:top + h, left:left + w] = im
        mc = np.zeros((size, size), np.uint8)
        mc[top:top + h, left:left + w] = mk
        imgs.append(canvas)
        masks.append((mc > 0).astype(np.float32))
    u8 = torch.from_numpy(np.stack(imgs))
    y = torch.from_numpy(np.stack(masks)).unsqueeze(1)
    val = torch.from_numpy(z["is_val"])
    return {"train_u8": u8[~val], "train_y": y[~val], "val_u8": u8[val], "val_y": y[val],
            "names": [str(n) for n in z["names"]]}


def normalise(u8_bgr: torch.Tensor) -> torch.Tensor:
    """uint8 BGR HWC [N,S,S,3] -> fp32 RGB NCHW, (x/255 - mean) / std  (train.py:108-112; infer_pth_gui.py:46-48)."""
    rgb = u8_bgr.flip(-1).float() / 255.0
    mean = torch.tensor(MEAN, device=u8_bgr.device)
    std = torch.tensor(STD, device=u8_bgr.device)
    return ((rgb - mean) / std).permute(0, 3, 1, 2).contiguous()


def dihedral(x: torch.Tensor, k: int) -> torch.Tensor:
    """k in 0..7: rotation by 90 degrees * (k & 3), then a horizontal flip if k & 4 (NCHW, square images)."""
    x = torch.rot90(x, k & 3, (2, 3))
    return x.flip(3) if k & 4 else x


def batches(n_train: int, batch: int, steps: int, seed: int):
    """Deterministic (indices, dihedral ids) per step: epochs of a seeded permutation, drop_last."""
    g = torch.Generator().manual_seed(seed)
    out, perm, pos = [], torch.randperm(n_train, generator=g), 0
    for _ in range(steps):
        if pos + batch > n_train:
            perm, pos = torch.randperm(n_train, generator=g), 0
        idx = perm[pos:pos + batch]
        pos += batch
        out.append((idx, torch.randint(0, 8, (batch,), generator=g)))
    return out


def make_batch(data, idx, ks, device):
    x = normalise(data["train_u8"][idx].to(device))
    y = data["train_y"][idx].to(device)
    x = torch.stack([dihedral(x[i:i + 1], int(k))[0] for i, k in enumerate(ks)])
    y = torch.stack([dihedral(y[i:i + 1], int(k))[0] for i, k in enumerate(ks)])
    return x.contiguous(), y.contiguous()


@torch.no_grad()
def val_metrics(model, data, device, batch: int = 18):
    """dice_coef / iou_coef of train.py:230-281 over the validation images (sigmoid(logits) > 0.5 <=> logits > 0)."""
    model.eval()
    dices, ious, loss = [], [], []
    xs = normalise(data["val_u8"].to(device))
    ys = data["val_y"].to(device)
    for i in range(0, xs.shape[0], batch):
        lg = model(xs[i:i + batch]).float()
        y = ys[i:i + batch]
        pred = (lg > 0).float()
        inter = (pred * y).flatten(1).sum(1)
        sp, st = pred.flatten(1).sum(1), y.flatten(1).sum(1)
        dices.append((2 * inter + 1e-7) / (sp + st + 1e-7))
        ious.append((inter + 1e-7) / (sp + st - inter + 1e-7))
        loss.append(F.binary_cross_entropy_with_logits(lg, y, reduction="none").flatten(1).mean(1))
    return float(torch.cat(dices).mean()), float(torch.cat(ious).mean()), float(torch.cat(loss).mean())


def lr_at(step: int, steps: int, peak: float, warmup: int = 30) -> float:
    """Learning rate of step `step` (1-based): linear warm-up, then cosine annealing to 0 at `steps` (the reference
    anneals per epoch, train.py:607 CosineAnnealingLR(T_max=epochs); here per step because the run is a few epochs)."""
    import math
    return peak * min(1.0, step / warmup) * 0.5 * (1.0 + math.cos(math.pi * min(1.0, step / steps)))


def train(model, step_fn, data, device, steps: int, batch: int, seed: int, eval_every: int, log=None, opt=None,
          peak_lr: float = None):
    """Runs `steps` optimisation steps (step_fn(x, y) -> loss tensor/float does zero_grad/forward/loss/backward/step,
    train.py:428-449) and evaluates the validation split every `eval_every` steps.  -> history list of dicts.
    opt + peak_lr: set param_groups[*]["lr"] = lr_at(step) before every step."""
    hist = []
    sched = batches(data["train_u8"].shape[0], batch, steps, seed)
    for s, (idx, ks) in enumerate(sched, 1):
        model.train()
        if opt is not None and peak_lr is not None:
            for gp in opt.param_groups:
                gp["lr"] = lr_at(s, steps, peak_lr)
        x, y = make_batch(data, idx, ks, device)
        loss = step_fn(x, y)
        if s % eval_every == 0 or s == steps:
            d, i, vl = val_metrics(model, data, device)
            hist.append({"step": s, "train_loss": float(loss), "val_dice": d, "val_iou": i, "val_bce": vl})
            if log:
                log(f"    step {s:4d} train_loss {float(loss):.4f} val_dice {d:.4f} val_iou {i:.4f} val_bce {vl:.4f}")
    return hist
