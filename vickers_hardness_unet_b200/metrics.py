"""Validation metrics on the device: drop-ins for `dice_coef` / `iou_coef` of /root/reference/train.py:230-281.

The reference thresholds the probability map at 0.5, computes per-image Dice / IoU and averages over the batch on the
host after a `.item()` per metric per batch (train.py:518-522).  Here one pair of kernels produces both numbers from the
logits (or probabilities) without extra passes; the host reads two floats.
"""
from __future__ import annotations

import torch

from . import _lib


def dice_iou(pred: torch.Tensor, target: torch.Tensor, eps: float = 1e-7, from_logits: bool = False) -> torch.Tensor:
    """Returns a CUDA tensor [2] = (mean Dice, mean IoU) over the batch.  pred: [N,1,H,W] probabilities, or logits with
    from_logits=True (sigmoid(x) > 0.5 <=> x > 0); target: [N,1,H,W] {0,1}."""
    if not pred.is_cuda or not target.is_cuda:
        raise _lib.UnetB200Error("metrics run on CUDA tensors only (no CPU fallback)")
    if pred.shape != target.shape or pred.dim() != 4:
        raise ValueError(f"expected equal [N,1,H,W] shapes, got {tuple(pred.shape)} and {tuple(target.shape)}")
    lib = _lib.load()
    p = pred.detach().to(torch.float32).contiguous()
    t = target.detach().to(torch.float32).contiguous()
    N = p.shape[0]
    hw = p.numel() // N
    scratch = torch.empty(lib.unetb200_seg_metrics_scratch_floats(N), dtype=torch.float32, device=p.device)
    out = torch.empty(2, dtype=torch.float32, device=p.device)
    rc = lib.unetb200_seg_metrics(p.data_ptr(), t.data_ptr(), N, hw, 0.0 if from_logits else 0.5, float(eps),
                                  scratch.data_ptr(), out.data_ptr(), torch.cuda.current_stream(p.device).cuda_stream)
    _lib.check_global(rc, "seg_metrics")
    return out


def dice_coef(prob: torch.Tensor, target: torch.Tensor, eps: float = 1e-7) -> float:
    """Same signature and value as /root/reference/train.py:230 (one device sync)."""
    return float(dice_iou(prob, target, eps)[0])


def iou_coef(prob: torch.Tensor, target: torch.Tensor, eps: float = 1e-7) -> float:
    """Same signature and value as /root/reference/train.py:262."""
    return float(dice_iou(prob, target, eps)[1])
