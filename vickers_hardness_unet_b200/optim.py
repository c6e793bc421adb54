"""Fused AdamW over the model's flat parameter array (one kernel instead of torch's multi-tensor loop).

Same update rule and defaults as the optimizer the reference builds at /root/reference/train.py:606
(`torch.optim.AdamW(model.parameters(), lr=..., weight_decay=1e-4)`): decoupled decay on every tensor (BatchNorm
gamma/beta and the head bias included), betas (0.9, 0.999), eps 1e-8, bias-corrected moments.
`param_groups[0]["lr"]` is read every step, so torch LR schedulers (CosineAnnealingLR, train.py:607) work unchanged.
"""
from __future__ import annotations

import torch

from . import _lib


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 grad_scale: float = 1.0, zero_grad: bool = False, overlap_allreduce: bool = False):
        if not hasattr(model, "flat_params"):
            raise TypeError("FusedAdamW takes the unet_b200.Unet module itself (it updates its flat parameter array)")
        super().__init__(list(model.parameters()), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._model = model
        self._step = 0
        self._m = self._v = None
        self.grad_scale = float(grad_scale)
        self._zero = bool(zero_grad)
        # data parallel only: leave the bucketed gradient all-reduces in flight at the end of backward and update bucket
        # by bucket as they land (nothing may read `.grad` between loss.backward() and step(): no clipping / unscale_)
        self._overlap = bool(overlap_allreduce)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        model = self._model
        p = model.flat_params
        if not p.is_cuda:
            raise _lib.UnetB200Error("FusedAdamW runs on CUDA sm_100a only — there is no CPU fallback")
        g = model._grad_buffer()
        if self._m is None or self._m.device != p.device:
            self._m, self._v = torch.zeros_like(p), torch.zeros_like(p)
        grp = self.param_groups[0]
        self._step += 1
        ctx = model._ctx
        if ctx is None:
            raise _lib.UnetB200Error("FusedAdamW.step() before any forward/backward")
        stream = torch.cuda.current_stream(p.device).cuda_stream
        b1, b2 = grp["betas"]
        dp = getattr(model, "_dp", None)
        if dp is not None and self._overlap:
            dp.defer_finish = True              # from the next backward on
            ranges = [(stage, b, e) for stage, (b, e) in enumerate(dp.ranges)]
        else:
            if dp is not None:
                dp.finish()                     # a previous deferred backward may have left buckets in flight
            ranges = [(None, 0, p.numel())]
        for stage, b, e in ranges:              # bucket order = backward completion order = all-reduce launch order
            if stage is not None:
                dp.wait(stage)
            ctx.check(ctx.lib.unetb200_adamw_step(ctx.handle, p.data_ptr() + 4 * b, g.data_ptr() + 4 * b,
                                                  self._m.data_ptr() + 4 * b, self._v.data_ptr() + 4 * b, e - b,
                                                  float(grp["lr"]), float(b1), float(b2), float(grp["eps"]),
                                                  float(grp["weight_decay"]), self._step, self.grad_scale,
                                                  int(self._zero), stream), "adamw_step")
        model._params_epoch += 1  # the library wrote the parameters: the bf16 operand caches are stale
        return loss
