"""Training-step host code: autograd bridge to libunetb200's train_forward / train_backward, the fused BCE+Dice loss.

Mirrors /root/reference/train.py:428-449 (`zero_grad -> logits = model(x) -> bce + dice -> backward -> step`).
The autograd graph of the network is written out by hand inside the library; PyTorch only sees ONE node
(`_UnetTrainFn`) whose backward fills the model's flat fp32 gradient array and attaches per-parameter views of it
as `.grad` (so torch.optim.AdamW, GradScaler.unscale_ and clip_grad_norm_ keep working unchanged).
"""
from __future__ import annotations

import torch

from . import _lib


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


class _UnetTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(fctx, x, anchor, model, frames=None):  # noqa: D401  (anchor: any parameter, makes the output require grad)
        ctx = model._ctx
        flat = model._flat
        g = model._grad_buffer()
        if frames is None:
            N, _, H, W = x.shape
            logits = torch.empty((N, 1, H, W), dtype=torch.float32, device=x.device)
            ctx.check(ctx.lib.unetb200_train_forward(ctx.handle, x.data_ptr(), logits.data_ptr(), flat["p"].data_ptr(),
                                                     flat["b"].data_ptr(), flat["c"].data_ptr(), g.data_ptr(), N,
                                                     _stream(x)), "train_forward")
        else:   # uint8 HWC frames: the reference's host pre-processing (train.py:108-112) runs in the input pack
            import ctypes as C
            bgr, mean, std = frames
            N, H, W, _ = x.shape
            logits = torch.empty((N, 1, H, W), dtype=torch.float32, device=x.device)
            ctx.check(ctx.lib.unetb200_train_forward_u8(ctx.handle, x.data_ptr(), int(bool(bgr)), (C.c_float * 3)(*mean),
                                                        (C.c_float * 3)(*std), logits.data_ptr(), flat["p"].data_ptr(),
                                                        flat["b"].data_ptr(), flat["c"].data_ptr(), g.data_ptr(), N,
                                                        _stream(x)), "train_forward_u8")
        model._buffers_epoch += 1  # running statistics changed behind PyTorch's back: eval must re-fold BatchNorm
        model._fwd_seq += 1        # the arena now holds THIS forward's activations
        fctx.model, fctx.N, fctx.seq = model, N, model._fwd_seq
        return logits

    @staticmethod
    def backward(fctx, dlogits):
        model, N = fctx.model, fctx.N
        ctx = model._ctx
        if fctx.seq != model._fwd_seq:
            raise _lib.UnetB200Error(
                "backward of a train-mode forward that is not the latest one: the library keeps the activations of ONE "
                "forward (one arena per model); call loss.backward() before the next model(x) in train() mode, or run "
                "the extra forward under model.eval() / on a copy of the model")
        dl = dlogits.detach().to(torch.float32).contiguous()
        g = model._grad_buffer()
        params = list(model.parameters())
        # accumulation semantics (no zero_grad between two backwards): the library overwrites, so keep the old sum
        attached = [p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, model._grad_views)]
        prev = g.clone() if any(attached) else None
        stream = _stream(dl)
        dp = getattr(model, "_dp", None)
        if dp is None:
            ctx.check(ctx.lib.unetb200_train_backward(ctx.handle, dl.data_ptr(), N, 0, 3, stream), "train_backward")
        else:
            for stage in range(4):
                ctx.check(ctx.lib.unetb200_train_backward(ctx.handle, dl.data_ptr(), N, stage, stage, stream),
                          "train_backward")
                dp.reduce(stage)  # async all-reduce of this stage's bucket, overlapped with the next stage
            if not dp.defer_finish:
                dp.finish()
        if prev is not None:
            if dp is not None:
                dp.finish()   # accumulation reads / writes the buckets: the all-reduces must have landed
            g.add_(prev)
        for p, v, a in zip(params, model._grad_views, attached):
            if p.grad is None:
                p.grad = v
            elif not a:
                p.grad.add_(v)
        return None, None, None, None


def unet_train_forward(model, ctx, x, stream, frames=None):
    anchor = next(model.parameters())
    return _UnetTrainFn.apply(x, anchor, model, frames)


# ------------------------------------------------------------------------------------------------ loss
_scratch = {}


def _loss_scratch(device):
    t = _scratch.get(device)
    if t is None:
        t = torch.empty(_lib.load().unetb200_loss_scratch_floats(), dtype=torch.float32, device=device)
        _scratch[device] = t
    return t


class _BceDiceFn(torch.autograd.Function):
    """(bce, dice) = (nn.BCEWithLogitsLoss()(x, y), smp DiceLoss("binary")(x, y)) in one pass over the logits."""

    @staticmethod
    def forward(fctx, logits, target, eps):
        if not logits.is_cuda:
            raise _lib.UnetB200Error("unet_b200 losses run on CUDA sm_100a only — there is no CPU fallback")
        x = logits.detach().to(torch.float32).contiguous()
        y = target.detach().to(torch.float32).contiguous()
        if x.numel() != y.numel():
            raise ValueError(f"logits {tuple(logits.shape)} and target {tuple(target.shape)} differ in size")
        lib = _lib.load()
        result = torch.empty(8, dtype=torch.float32, device=x.device)
        _lib.check_global(lib.unetb200_loss_bce_dice_forward(x.data_ptr(), y.data_ptr(), x.numel(), float(eps),
                                                             _loss_scratch(x.device).data_ptr(), result.data_ptr(),
                                                             _stream(x)), "loss_forward")
        fctx.save_for_backward(x, y, result)
        fctx.eps = float(eps)
        fctx.in_dtype = logits.dtype
        return result[0].clone(), result[1].clone()

    @staticmethod
    def backward(fctx, g_bce, g_dice):
        x, y, result = fctx.saved_tensors
        lib = _lib.load()
        dx = torch.empty_like(x)
        gb = g_bce.detach().to(torch.float32).contiguous() if g_bce is not None else None
        gd = g_dice.detach().to(torch.float32).contiguous() if g_dice is not None else None
        _lib.check_global(lib.unetb200_loss_bce_dice_backward(
            x.data_ptr(), y.data_ptr(), result.data_ptr(), gb.data_ptr() if gb is not None else None,
            gd.data_ptr() if gd is not None else None, 1.0, fctx.eps, dx.data_ptr(), x.numel(), _stream(x)),
            "loss_backward")
        return dx.to(fctx.in_dtype), None, None


def bce_dice(logits, target, eps: float = 1e-7):
    return _BceDiceFn.apply(logits, target, eps)


def dice_loss(logits, target, eps: float = 1e-7):
    return _BceDiceFn.apply(logits, target, eps)[1]
