"""`Unet`: drop-in for `segmentation_models_pytorch.Unet("resnet34", in_channels=3, classes=1, activation=None)`.

Same constructor call, same `nn.Module` surface and the same 278-entry state_dict key layout as the model the
reference builds at /root/reference/train.py:372-378, infer_pth_gui.py:31-33, ui_infer_rectangle.py:496-499 and
ui_infer_quadrilateral.py:638-641 — but `forward` runs the hand-written sm_100a kernels of libunetb200.so.
Parameters are fp32 master copies living in ONE flat tensor (views per state_dict entry), which is what the fused
optimizer and the bucketed gradient all-reduce operate on.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib


class _Node(nn.Module):
    """Anonymous container so that parameters get smp's dotted key names."""


def _init_tensor(name: str, t: torch.Tensor):
    """Default init of smp / torchvision; only matters without a checkpoint."""
    if name.endswith("running_var"):
        t.fill_(1.0)
    elif name.endswith("running_mean"):
        t.zero_()
    elif t.dim() == 4:
        if name.startswith("encoder."):  # torchvision/models/resnet.py:208-210
            nn.init.kaiming_normal_(t, mode="fan_out", nonlinearity="relu")
        elif name.startswith("decoder."):
            nn.init.kaiming_uniform_(t, mode="fan_in", nonlinearity="relu")
        else:
            nn.init.xavier_uniform_(t)
    elif name.endswith(".weight"):
        t.fill_(1.0)  # BatchNorm gamma
    else:
        t.zero_()  # BatchNorm beta / head bias


def _find_resnet34_checkpoint() -> str:
    """`encoder_weights="imagenet"` (/root/reference/train.py:753,595) without a download: a torchvision-format
    resnet34 state_dict on local disk — $UNETB200_RESNET34_WEIGHTS or torch hub's checkpoint cache (where smp /
    torchvision would have put `resnet34-*.pth`)."""
    import glob
    import os

    cands = []
    env = os.environ.get("UNETB200_RESNET34_WEIGHTS")
    if env:
        cands.append(env)
    try:
        hub = torch.hub.get_dir()
    except Exception:  # pragma: no cover
        hub = os.path.expanduser("~/.cache/torch/hub")
    cands += sorted(glob.glob(os.path.join(hub, "checkpoints", "resnet34-*.pth")))
    for c in cands:
        if os.path.isfile(c):
            return c
    raise FileNotFoundError(
        "encoder_weights='imagenet' needs a local torchvision-format resnet34 checkpoint (there is no network "
        "download here): set UNETB200_RESNET34_WEIGHTS=/path/to/resnet34-b627a593.pth or place the file in "
        f"{os.path.join(hub, 'checkpoints')}; or pass encoder_weights=None and load a state_dict")


class _ShapeProxy:
    """What `_context` needs to know about an input: device and the NCHW shape it stands for."""

    def __init__(self, t, N, H, W):
        self.is_cuda, self.device, self.shape = t.is_cuda, t.device, (N, 3, H, W)


class Unet(nn.Module):
    def __init__(self, encoder_name: str = "resnet34", encoder_depth: int = 5, encoder_weights=None,
                 decoder_use_batchnorm: bool = True, decoder_channels=(256, 128, 64, 32, 16),
                 decoder_attention_type=None, in_channels: int = 3, classes: int = 1, activation=None,
                 aux_params=None, precision: str = "bf16", **kwargs):
        super().__init__()
        if precision not in _lib.LIB_PATHS:
            raise ValueError(f"precision must be one of {sorted(_lib.LIB_PATHS)} (got {precision!r})")
        # 16-bit storage format of activations / conv operands: "bf16" (default, training + inference) or "fp16"
        # (inference only: same tensor-core rate, 10 instead of 7 mantissa bits).  May be changed before a forward.
        self.precision = precision
        if encoder_name != "resnet34":
            raise ValueError(f"unet_b200 implements encoder_name='resnet34' only (got {encoder_name!r})")
        if encoder_weights not in (None, "imagenet"):
            raise ValueError(f"encoder_weights must be None or 'imagenet' (got {encoder_weights!r})")
        pretrained = _find_resnet34_checkpoint() if encoder_weights == "imagenet" else None
        if (encoder_depth != 5 or tuple(decoder_channels) != (256, 128, 64, 32, 16) or not decoder_use_batchnorm
                or decoder_attention_type is not None or in_channels != 3 or classes != 1
                or activation is not None or aux_params is not None):
            raise ValueError("unet_b200 implements exactly smp.Unet('resnet34', in_channels=3, classes=1, "
                             "activation=None) with default decoder settings")
        table = _lib.tensor_table()
        lib = _lib.load()
        self._table = table
        n_p, n_b, n_c = lib.unetb200_num_params(), lib.unetb200_num_buffers(), lib.unetb200_num_counters()
        flat_p = torch.zeros(n_p, dtype=torch.float32)
        flat_b = torch.zeros(n_b, dtype=torch.float32)
        flat_c = torch.zeros(n_c, dtype=torch.int64)
        self._flat = {"p": flat_p, "b": flat_b, "c": flat_c}
        self._entries = []  # (owner module, attr, kind, offset, shape)
        for name, shape, off, kind in table:
            parts = name.split(".")
            mod = self
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, _Node())
                mod = mod._modules[p]
            numel = int(math.prod(shape)) if shape else 1
            if kind == 0:
                view = flat_p[off:off + numel].view(shape)
                with torch.no_grad():
                    _init_tensor(name, view)
                mod.register_parameter(parts[-1], nn.Parameter(view))
            elif kind == 1:
                view = flat_b[off:off + numel].view(shape)
                _init_tensor(name, view)
                mod.register_buffer(parts[-1], view)
            else:
                mod.register_buffer(parts[-1], flat_c[off:off + 1].view(()))
            self._entries.append((mod, parts[-1], kind, off, shape, numel))
        self._ctx = None
        self._packed_version = None
        self._params_epoch = 0   # bumped when the library itself writes parameters (FusedAdamW)
        self._buffers_epoch = 0  # bumped when the library updates the BatchNorm running statistics (train forward)
        self._grad_views = None
        self._dp = None          # distributed.GradBucketReducer when data parallelism is enabled
        self._fwd_seq = 0        # id of the last train-mode forward (its activations are the ones in the arena)
        self._param_list = None
        self.name = "u-resnet34"
        if pretrained is not None:
            self._load_torchvision_encoder(pretrained)

    def _load_torchvision_encoder(self, path: str):
        """Copy a torchvision resnet34 state_dict (keys `conv1.weight`, `layer1.0.bn1.running_mean`, ..., `fc.*`) into
        `encoder.*`, as smp's ResNetEncoder.load_state_dict does (it drops `fc.weight` / `fc.bias`)."""
        sd = torch.load(path, map_location="cpu", weights_only=True)
        sd = {k: v for k, v in sd.items() if not k.startswith("fc.")}
        own = {k[len("encoder."):]: v for k, v in self.state_dict().items() if k.startswith("encoder.")}
        missing = sorted(set(own) - set(sd))
        unexpected = sorted(set(sd) - set(own))
        if missing or unexpected:
            raise RuntimeError(f"{path} is not a torchvision resnet34 state_dict: missing {missing[:4]}, "
                               f"unexpected {unexpected[:4]}")
        with torch.no_grad():
            for k, v in sd.items():
                own[k].copy_(v)

    # ------------------------------------------------------------------ copying / pickling (EMA, torch.save(model))
    def __getstate__(self):
        """The native context (a ctypes handle) and everything derived from it stay behind: a copy re-creates its own
        context at its first forward.  Makes copy.deepcopy(model) and torch.save(model) work after a forward."""
        st = self.__dict__.copy()
        st["_ctx"] = None
        st["_dp"] = None
        st["_packed_version"] = None
        st["_grad_views"] = None
        st["_param_list"] = None
        st["_flat"] = {k: v for k, v in self._flat.items() if k != "g"}
        return st

    def __setstate__(self, state):
        super().__setstate__(state)
        self._reflatten()   # parameters / buffers back to views of ONE flat array each

    # ------------------------------------------------------------------ flat storage upkeep
    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self._reflatten()
        return self

    def _reflatten(self):
        """nn.Module._apply re-creates every tensor separately; gather them back into the flat arrays."""
        first = next(p for p in self.parameters())
        dev = first.device
        flat = {"p": torch.empty(self._flat["p"].numel(), dtype=torch.float32, device=dev),
                "b": torch.empty(self._flat["b"].numel(), dtype=torch.float32, device=dev),
                "c": torch.empty(self._flat["c"].numel(), dtype=torch.int64, device=dev)}
        with torch.no_grad():
            for mod, attr, kind, off, shape, numel in self._entries:
                if kind == 0:
                    p = mod._parameters[attr]
                    view = flat["p"][off:off + numel].view(shape)
                    view.copy_(p.data.to(torch.float32))
                    p.data = view
                elif kind == 1:
                    view = flat["b"][off:off + numel].view(shape)
                    view.copy_(mod._buffers[attr].to(torch.float32))
                    mod._buffers[attr] = view
                else:
                    view = flat["c"][off:off + 1].view(())
                    view.copy_(mod._buffers[attr])
                    mod._buffers[attr] = view
        self._flat = flat
        self._packed_version = None
        self._grad_views = None
        self._param_list = None
        for p in self.parameters():
            p.grad = None
        if self._ctx is not None and self._ctx.device != (dev.index if dev.type == "cuda" else -1):
            self._ctx = None

    @property
    def flat_params(self) -> torch.Tensor:
        return self._flat["p"]

    @property
    def flat_buffers(self) -> torch.Tensor:
        return self._flat["b"]

    @property
    def flat_grads(self) -> torch.Tensor:
        return self._grad_buffer()

    def _grad_buffer(self) -> torch.Tensor:
        """Flat fp32 gradient array (same layout as flat_params); `.grad` of every parameter is a view of it."""
        g = self._flat.get("g")
        p = self._flat["p"]
        if g is None or g.device != p.device:
            g = torch.zeros_like(p)
            self._flat["g"] = g
            self._grad_views = None
        if self._grad_views is None:
            self._grad_views = [g[off:off + numel].view(shape) for (_, _, kind, off, shape, numel) in self._entries
                                if kind == 0]
        return g

    # ------------------------------------------------------------------ native context
    def _context(self, x: torch.Tensor) -> "_lib.Context":
        if not x.is_cuda:
            raise _lib.UnetB200Error(
                "unet_b200.Unet runs on CUDA sm_100a only — there is no CPU fallback. "
                "Move the model and the input to a B200 (`.to('cuda')`).")
        if self._flat["p"].device != x.device:
            raise _lib.UnetB200Error(f"model is on {self._flat['p'].device}, input on {x.device}")
        N, Cin, H, W = x.shape
        if Cin != 3:
            raise ValueError(f"expected [N,3,H,W] input, got {tuple(x.shape)}")
        if H % 32 or W % 32:
            raise ValueError(f"H and W must be divisible by 32 (got {H}x{W})")  # same rule as smp>=0.5
        c = self._ctx
        if self.training and self.precision != "bf16":
            raise _lib.UnetB200Error("precision='fp16' is an inference mode: training needs bfloat16's range "
                                     "(set model.precision = 'bf16' or call model.eval())")
        if (c is None or c.H != H or c.W != W or c.max_batch < N or c.device != x.device.index
                or c.precision != self.precision):
            if c is not None:
                torch.cuda.synchronize(x.device)
                c.close()
            mb = N if c is None or c.H != H or c.W != W else max(N, c.max_batch)
            self._ctx = c = _lib.Context(x.device.index, mb, H, W, self.precision)
            self._packed_version = None
        return c

    def _sync_weights(self, ctx, stream, fold_bn=None):
        """Refresh the library's bf16 operand caches when the master tensors changed.  fold_bn: also fold eval-mode
        BatchNorm (default: only in eval mode — a training step normalises with batch statistics)."""
        fold_bn = (not self.training) if fold_bn is None else fold_bn
        # `p.data = view` (_reflatten) leaves every Parameter with its OWN version counter, so an in-place update through
        # the parameter (torch.optim.AdamW, p.add_()) does not bump the flat tensor's version: sum the per-tensor ones
        if self._param_list is None:
            self._param_list = list(self.parameters()) + [b for b in self.buffers() if b.dtype == torch.float32]
        ver = (self._flat["p"]._version, self._flat["b"]._version, sum(t._version for t in self._param_list), id(ctx),
               self._params_epoch, self._buffers_epoch if fold_bn else -1)
        if ver != self._packed_version:
            ctx.check(ctx.lib.unetb200_load_weights_ex(ctx.handle, self._flat["p"].data_ptr(),
                                                       self._flat["b"].data_ptr(), int(not fold_bn), stream),
                      "load_weights")
            self._packed_version = ver

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dtype == torch.uint8:
            return self.forward_frames(x)
        ctx = self._context(x)
        x = x.detach().to(torch.float32).contiguous()
        stream = torch.cuda.current_stream(x.device).cuda_stream
        self._sync_weights(ctx, stream)
        if self.training:
            # batch-statistics BatchNorm + running-stat update, activations kept for loss.backward()
            from .train import unet_train_forward
            return unet_train_forward(self, ctx, x, stream)
        N, _, H, W = x.shape
        logits = torch.empty((N, 1, H, W), dtype=torch.float32, device=x.device)
        ctx.check(ctx.lib.unetb200_forward_infer(ctx.handle, x.data_ptr(), logits.data_ptr(), None, None, 0.5, N,
                                                 stream), "forward_infer")
        return logits

    def forward_frames(self, frames: torch.Tensor, bgr: bool = True, mean=None, std=None) -> torch.Tensor:
        """`model(x)` on uint8 HWC frames [N,H,W,3] already on the GPU (what cv2.imread + the letterbox produce): the
        reference's host pre-processing — BGR->RGB, /255, (x - mean) / std (train.py:108-112, infer_pth_gui.py:46-48) —
        runs inside the input-pack kernel, in train and eval mode alike.  `model(frames_u8)` dispatches here with the
        ImageNet constants."""
        import ctypes as C
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
            raise ValueError(f"expected uint8 [N,H,W,3] frames, got {frames.dtype} {tuple(frames.shape)}")
        mean = self.IMAGENET_MEAN if mean is None else mean
        std = self.IMAGENET_STD if std is None else std
        N, H, W, _ = frames.shape
        ctx = self._context(_ShapeProxy(frames, N, H, W))
        frames = frames.detach().contiguous()
        stream = torch.cuda.current_stream(frames.device).cuda_stream
        self._sync_weights(ctx, stream)
        if self.training:
            from .train import unet_train_forward
            return unet_train_forward(self, ctx, frames, stream, frames=(bgr, mean, std))
        logits = torch.empty((N, 1, H, W), dtype=torch.float32, device=frames.device)
        ctx.check(ctx.lib.unetb200_forward_infer_u8(ctx.handle, frames.data_ptr(), int(bool(bgr)), (C.c_float * 3)(*mean),
                                                    (C.c_float * 3)(*std), logits.data_ptr(), None, None, 0.5, N, stream),
                  "forward_infer_u8")
        return logits

    @torch.no_grad()
    def predict_mask(self, x: torch.Tensor, threshold: float = 0.5, return_prob: bool = False):
        """Fused `sigmoid(model(x)) >= threshold` (infer_pth_gui.py:50-52): uint8 {0,255} mask [N,1,H,W]."""
        ctx = self._context(x)
        x = x.detach().to(torch.float32).contiguous()
        stream = torch.cuda.current_stream(x.device).cuda_stream
        self._sync_weights(ctx, stream, fold_bn=True)
        N, _, H, W = x.shape
        mask = torch.empty((N, 1, H, W), dtype=torch.uint8, device=x.device)
        prob = torch.empty((N, 1, H, W), dtype=torch.float32, device=x.device) if return_prob else None
        ctx.check(ctx.lib.unetb200_forward_infer(ctx.handle, x.data_ptr(), None,
                                                 prob.data_ptr() if return_prob else None, mask.data_ptr(),
                                                 float(threshold), N, stream), "forward_infer")
        return (mask, prob) if return_prob else mask

    # ------------------------------------------------------------------ host-buffer inference (numpy / pinned tensors)
    IMAGENET_MEAN = (0.485, 0.456, 0.406)   # /root/reference/infer_pth_gui.py:15-16, ui_infer_rectangle.py:41-42
    IMAGENET_STD = (0.229, 0.224, 0.225)

    def _host_context(self, N: int, H: int, W: int) -> "_lib.Context":
        dev = self._flat["p"].device
        if dev.type != "cuda":
            raise _lib.UnetB200Error("the model must live on a CUDA (sm_100a) device; there is no CPU fallback")
        if H % 32 or W % 32:
            raise ValueError(f"H and W must be divisible by 32 (got {H}x{W})")
        c = self._ctx
        if c is None or c.H != H or c.W != W or c.max_batch < N or c.device != dev.index or c.precision != self.precision:
            if c is not None:
                torch.cuda.synchronize(dev)
                c.close()
            self._ctx = c = _lib.Context(dev.index, N, H, W, self.precision)
            self._packed_version = None
        self._sync_weights(c, torch.cuda.current_stream(dev).cuda_stream, fold_bn=True)
        return c

    @torch.no_grad()
    def submit_host(self, slot: int, x: torch.Tensor, mask_out: torch.Tensor = None, prob_out: torch.Tensor = None,
                    logits_out: torch.Tensor = None, threshold: float = 0.5, bgr: bool = True,
                    mean=IMAGENET_MEAN, std=IMAGENET_STD):
        """Enqueue one inference request from HOST memory (ideally pinned) and return immediately; `wait_host(slot)`
        blocks until the requested outputs are in the given host tensors.  x: fp32 [N,3,H,W] (already normalised, as
        infer_pth_gui.py:46-49 builds it) or uint8 [N,H,W,3] camera frames, in which case BGR->RGB, /255 and
        (x-mean)/std run on the device.  Two slots let request k+1's upload overlap request k's compute."""
        import ctypes as C
        if x.is_cuda or not x.is_contiguous():
            raise ValueError("submit_host takes a contiguous CPU tensor")
        u8 = x.dtype == torch.uint8
        if u8:
            N, H, W, Cc = x.shape
        else:
            if x.dtype != torch.float32:
                raise ValueError("submit_host takes float32 NCHW or uint8 NHWC input")
            N, Cc, H, W = x.shape
        if Cc != 3:
            raise ValueError(f"expected 3 channels, got {tuple(x.shape)}")
        outs = []
        for t, dt in ((logits_out, torch.float32), (prob_out, torch.float32), (mask_out, torch.uint8)):
            if t is not None and (t.is_cuda or t.dtype != dt or t.numel() != N * H * W or not t.is_contiguous()):
                raise ValueError("output tensors must be contiguous CPU tensors of N*H*W elements (float32 / uint8)")
            outs.append(t.data_ptr() if t is not None else None)
        ctx = self._host_context(N, H, W)
        if u8:
            m3 = (C.c_float * 3)(*mean)
            s3 = (C.c_float * 3)(*std)
            rc = ctx.lib.unetb200_infer_host_u8_submit(ctx.handle, slot, x.data_ptr(), int(bool(bgr)), m3, s3, outs[0],
                                                       outs[1], outs[2], float(threshold), N)
        else:
            rc = ctx.lib.unetb200_infer_host_submit(ctx.handle, slot, x.data_ptr(), outs[0], outs[1], outs[2],
                                                    float(threshold), N)
        ctx.check(rc, "infer_host_submit")

    def wait_host(self, slot: int):
        ctx = self._ctx
        if ctx is None:
            raise _lib.UnetB200Error("wait_host: nothing was submitted")
        ctx.check(ctx.lib.unetb200_infer_host_wait(ctx.handle, slot), "infer_host_wait")

    def predict_mask_host(self, x: torch.Tensor, threshold: float = 0.5, **kw) -> torch.Tensor:
        """Blocking host-to-host `sigmoid(model(x)) >= threshold`: uint8 {0,255} CPU tensor [N,1,H,W]."""
        N = x.shape[0]
        H, W = (x.shape[1], x.shape[2]) if x.dtype == torch.uint8 else (x.shape[2], x.shape[3])
        mask = torch.empty((N, 1, H, W), dtype=torch.uint8)
        self.submit_host(0, x, mask_out=mask, threshold=threshold, **kw)
        self.wait_host(0)
        return mask
