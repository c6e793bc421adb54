"""Loss entry points mirroring `smp.losses` (only DiceLoss is used: /root/reference/train.py:601)."""
from __future__ import annotations

import torch.nn as nn


class DiceLoss(nn.Module):
    """smp.losses.DiceLoss(mode="binary") — batch-global soft Dice on logits (smooth=0, eps=1e-7)."""

    def __init__(self, mode: str = "binary", classes=None, log_loss: bool = False, from_logits: bool = True,
                 smooth: float = 0.0, ignore_index=None, eps: float = 1e-7):
        super().__init__()
        if mode != "binary" or classes is not None or log_loss or not from_logits or smooth != 0.0 \
                or ignore_index is not None:
            raise ValueError("unet_b200 implements DiceLoss(mode='binary') with smp defaults only")
        self.eps = float(eps)

    def forward(self, y_pred, y_true):
        from .train import dice_loss
        return dice_loss(y_pred, y_true, self.eps)


class BCEDiceLoss(nn.Module):
    """`nn.BCEWithLogitsLoss()(x, y) + DiceLoss("binary")(x, y)` (/root/reference/train.py:438) as ONE fused pass."""

    def __init__(self, eps: float = 1e-7):
        super().__init__()
        self.eps = float(eps)

    def forward(self, y_pred, y_true):
        from .train import bce_dice
        bce, dice = bce_dice(y_pred, y_true, self.eps)
        return bce + dice
