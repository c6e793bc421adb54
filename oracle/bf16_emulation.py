"""fp32 emulation of the bf16 storage points of the CUDA path (TEST INFRASTRUCTURE, see oracle/__init__.py).

The product keeps activations and conv weights in bf16 (fp32 accumulate).  Measured on this model (random init,
calibrated BatchNorm statistics, noise input) rounding ONLY the conv weights and the input image to bf16 already moves
the logits by ~0.03-0.09 mean-abs / 0.2-0.7 max-abs against the fp32 oracle (|logit| mean ~0.8, max ~5-7), i.e. the
BASELINE north_star figure "2e-2 max-abs / 1e-3 mean-abs" is below what bf16 operands can deliver on this network
(fp16 operands: ~5e-3..2e-2 mean, 3e-2..0.15 max).  The parity tests therefore check the CUDA path against THIS
emulation (same rounding points, fp32 math in between) and report the distance to the fp32 oracle beside it.

Rounding points emulated (eval mode):
  input image -> bf16; every conv weight except the seg head -> bf16; the decoder conv1 weights that act on the
  up-sampled channels are first summed per output parity (3x3 -> 2x2 on the low-res tensor) and then rounded, exactly
  as the packed operand of the CUDA kernel; every unit output (after BN(+residual)+ReLU, and after the downsample BN)
  -> bf16.  BatchNorm is applied in fp32 on the fp32 accumulator, the seg head runs in fp32 on bf16 inputs.
"""
from __future__ import annotations

import copy

import torch
import torch.nn.functional as F


class _RoundSTE(torch.autograd.Function):
    """bf16 rounding with a straight-through (identity) gradient: the backward of the emulated network is then the
    exact gradient of the rounded-forward computation, which is what the CUDA backward approximates."""

    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


def _r(t: torch.Tensor) -> torch.Tensor:
    return _RoundSTE.apply(t)


def round_mantissa(t: torch.Tensor, bits: int) -> torch.Tensor:
    """fp32 -> `bits` explicit mantissa bits, round-to-nearest-even, fp32 exponent range (bits = 7 is bf16, 10 is
    fp16 / tf32 precision, 15 is a bf16 hi + lo pair, 23 is the identity).  Used by the precision sweep that answers
    "how many mantissa bits does the north_star logit tolerance need" (tests/test_gpu_trained.py)."""
    if bits >= 23:
        return t
    sh = 23 - bits
    i = t.contiguous().view(torch.int32)
    i = ((i + ((1 << (sh - 1)) - 1) + ((i >> sh) & 1)) >> sh) << sh
    return i.view(torch.float32)


def _dec_conv1_parity(x_low, skip, w, cup, round_up_part=False, _r=_r):
    """cat(nearest2x(x_low), skip) (*) w, computed as the CUDA path does: 4 output parities, summed 2x2 weights.
    round_up_part: the contribution of the up-sampled channels is stored in bf16 before the skip part is added (decoder
    block 3: two tconv launches, csrc/unet.cuh tc == 3)."""
    N, _, Hl, Wl = x_low.shape
    cout = w.shape[0]
    w_up, w_sk = w[:, :cup], w[:, cup:]
    rsets = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}  # parity -> row sets of a=0, a=1
    full = F.conv2d(skip, _r(w_sk), padding=1) if skip is not None else None
    rows = []
    for ph in (0, 1):
        cols = []
        for pw in (0, 1):
            weff = torch.stack([torch.stack([sum(w_up[:, :, r, s] for r in rsets[ph][a] for s in rsets[pw][b])
                                             for b in (0, 1)], -1) for a in (0, 1)], -2)  # [cout, cup, 2(a), 2(b)]
            weff = _r(weff)
            # low-res taps at offsets (a-1+ph, b-1+pw): pad so that a 2x2 VALID conv lines up
            xp = F.pad(x_low, (1 - pw, pw, 1 - ph, ph))
            acc = F.conv2d(xp, weff)
            if round_up_part:
                acc = _r(acc)
            if full is not None:
                acc = acc + full[:, :, ph::2, pw::2]
            cols.append(acc)
        rows.append(torch.stack(cols, -1).reshape(N, cout, Hl, 2 * Wl))       # interleave the two column parities
    return torch.stack(rows, -2).reshape(N, cout, 2 * Hl, 2 * Wl)              # interleave the two row parities


# every decoder conv1 uses the parity-folded 2x2 weights on the low-res tensor for its up-sampled channels; block 3
# (and, in inference only, blocks 0-2: csrc/unet.cuh tc == 4) stores that part in bf16 and adds it to the skip-channel
# conv through the residual path (two launches); block 4 has no skip
SPLIT_DECODER_BLOCKS = (3,)
SPLIT_DECODER_BLOCKS_EVAL = (0, 1, 2, 3)


def emulated_forward(o, x, train: bool, rnd=None, taps=None):
    """Forward of oracle `o` with the CUDA path's bf16 rounding points.

    rnd : the rounding applied at those points (default: bf16 with a straight-through gradient; `lambda t: t` gives the
          plain fp32 oracle forward, `lambda t: round_mantissa(t, k)` a k-bit-mantissa storage format).
    taps: optional dict that receives the named activations the CUDA path materialises ("<conv weight key>/out",
          "encoder.maxpool/out"), for the per-layer error-growth report.

    eval : BatchNorm folded into the conv epilogue: a = bf16(relu(acc*scale + shift (+ identity)))
    train: raw conv output stored first, z = bf16(acc); batch statistics are taken from the ROUNDED z; then
           a = bf16(relu(bn(z) (+ identity)))  -- running statistics of `o` are updated like nn.BatchNorm2d.
    """
    def bn(t, m):
        if train:
            if m.num_batches_tracked is not None:
                m.num_batches_tracked += 1
            return F.batch_norm(_r(t), m.running_mean, m.running_var, m.weight, m.bias, True, m.momentum, m.eps)
        return F.batch_norm(t, m.running_mean, m.running_var, m.weight, m.bias, False, 0.0, m.eps)

    _r = rnd if rnd is not None else globals()["_r"]

    def tap(name, t):
        if taps is not None:
            taps[name] = t.detach()
        return t

    e = o.encoder
    x = _r(x)
    f1 = tap("encoder.conv1.weight/out", _r(F.relu(bn(F.conv2d(x, _r(e.conv1.weight), None, 2, 3), e.bn1))))
    t = tap("encoder.maxpool/out", F.max_pool2d(f1, 3, 2, 1))
    feats = [f1]
    for li, layer in enumerate((e.layer1, e.layer2, e.layer3, e.layer4)):
        for bi, blk in enumerate(layer):
            pre = f"encoder.layer{li + 1}.{bi}"
            idn = t
            u = tap(pre + ".conv1.weight/out", _r(F.relu(bn(F.conv2d(t, _r(blk.conv1.weight), None, blk.stride, 1), blk.bn1))))
            if blk.downsample is not None:
                idn = tap(pre + ".downsample.0.weight/out",
                          _r(bn(F.conv2d(t, _r(blk.downsample[0].weight), None, blk.stride, 0), blk.downsample[1])))
            t = tap(pre + ".conv2.weight/out", _r(F.relu(bn(F.conv2d(u, _r(blk.conv2.weight), None, 1, 1), blk.bn2) + idn)))
        feats.append(t)
    skips = [feats[3], feats[2], feats[1], feats[0], None]
    cups = [512, 256, 128, 64, 32]
    for i, blk in enumerate(o.decoder.blocks):
        z = _dec_conv1_parity(t, skips[i], blk.conv1[0].weight, cups[i],
                              round_up_part=i in (SPLIT_DECODER_BLOCKS if train else SPLIT_DECODER_BLOCKS_EVAL), _r=_r)
        u = tap(f"decoder.blocks.{i}.conv1.0.weight/out", _r(F.relu(bn(z, blk.conv1[1]))))
        t = tap(f"decoder.blocks.{i}.conv2.0.weight/out",
                _r(F.relu(bn(F.conv2d(u, _r(blk.conv2[0].weight), None, 1, 1), blk.conv2[1]))))
    head = o.segmentation_head[0]
    return F.conv2d(t, head.weight, head.bias, 1, 1)


class Bf16EmulatedUnet(torch.nn.Module):
    """Eval-mode forward of the oracle with the CUDA path's bf16 rounding points."""

    def __init__(self, oracle):
        super().__init__()
        self.o = copy.deepcopy(oracle).eval()

    @torch.no_grad()
    def forward(self, x):
        return emulated_forward(self.o, x, False)


class Bf16EmulatedTrainUnet(torch.nn.Module):
    """Train-mode forward (batch statistics) with the CUDA path's rounding points and straight-through gradients.
    Owns a deep copy of the oracle: `.o.named_parameters()` receive the gradients of the rounded-forward network."""

    def __init__(self, oracle):
        super().__init__()
        self.o = copy.deepcopy(oracle).train()

    def forward(self, x):
        return emulated_forward(self.o, x, True)
