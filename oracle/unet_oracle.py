"""fp32 PyTorch restatement of ``segmentation_models_pytorch.Unet("resnet34")`` + ``DiceLoss("binary")``.

TEST INFRASTRUCTURE — not part of the product path (see oracle/__init__.py).

PARITY PIN STATUS: **partially pinned**.
  * ``segmentation_models_pytorch`` (PyPI, version un-pinned by the reference: it ships no requirements file)
    is neither vendored in /root/reference nor installed here, and the reference has no tests, golden vectors
    or checkpoints (``.MISSING_LARGE_BLOBS``), so the decoder / head / DiceLoss restatement is **parity unpinned**
    against smp itself: it follows smp's published source from memory
    (decoders/unet/decoder.py, base/modules.py::Conv2dReLU, base/heads.py::SegmentationHead,
    base/initialization.py, losses/dice.py, losses/_functional.py::soft_dice_score).
  * The encoder IS pinned: it is torchvision's own ``ResNet(BasicBlock,[3,4,6,3])``
    (torchvision/models/resnet.py:166-276, installed 0.26.0), and tests/test_oracle.py checks the feature
    pyramid against ``torchvision.models.resnet34`` run stage by stage.
  * Structural anchors from the reference call sites: ``smp.Unet(encoder_name="resnet34", encoder_weights=None,
    in_channels=3, classes=1, activation=None)`` (/root/reference/train.py:372-378, infer_pth_gui.py:31-33);
    24,436,369 parameters / 278 state-dict entries.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.models.resnet import BasicBlock, ResNet


class _Encoder(ResNet):
    """smp.encoders.resnet.ResNetEncoder for resnet34: torchvision ResNet minus fc/avgpool, returns 6 features.

    Op order follows torchvision/models/resnet.py:266-276 (conv1, bn1, relu, maxpool, layer1..4).
    """

    def __init__(self):
        super().__init__(BasicBlock, [3, 4, 6, 3])
        del self.fc
        del self.avgpool

    def forward(self, x):  # type: ignore[override]
        feats = [x]
        x = self.relu(self.bn1(self.conv1(x)))
        feats.append(x)
        x = self.layer1(self.maxpool(x))
        feats.append(x)
        x = self.layer2(x)
        feats.append(x)
        x = self.layer3(x)
        feats.append(x)
        x = self.layer4(x)
        feats.append(x)
        return feats


class _Conv2dReLU(nn.Sequential):
    """smp.base.modules.Conv2dReLU with use_batchnorm=True: conv(bias=False) -> BN -> ReLU."""

    def __init__(self, cin, cout):
        super().__init__(
            nn.Conv2d(cin, cout, 3, padding=1, bias=False),
            nn.BatchNorm2d(cout),
            nn.ReLU(inplace=True),
        )


class _DecoderBlock(nn.Module):
    """smp UnetDecoder DecoderBlock (attention_type=None): nearest 2x -> cat([x, skip]) -> conv1 -> conv2."""

    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = _Conv2dReLU(cin + cskip, cout)
        self.conv2 = _Conv2dReLU(cout, cout)

    def forward(self, x, skip=None):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)
        return self.conv2(self.conv1(x))


class _Decoder(nn.Module):
    """smp UnetDecoder(encoder_channels=(3,64,64,128,256,512), decoder_channels=(256,128,64,32,16))."""

    def __init__(self):
        super().__init__()
        cin = [512, 256, 128, 64, 32]
        cskip = [256, 128, 64, 64, 0]
        cout = [256, 128, 64, 32, 16]
        self.blocks = nn.ModuleList(_DecoderBlock(a, b, c) for a, b, c in zip(cin, cskip, cout))

    def forward(self, feats):
        feats = feats[1:][::-1]  # drop the input-resolution feature, deepest first
        x, skips = feats[0], feats[1:]
        for i, blk in enumerate(self.blocks):
            x = blk(x, skips[i] if i < len(skips) else None)
        return x


class OracleUnet(nn.Module):
    """Same module tree / state_dict keys as smp.Unet("resnet34", in_channels=3, classes=1, activation=None)."""

    def __init__(self):
        super().__init__()
        self.encoder = _Encoder()
        self.decoder = _Decoder()
        self.segmentation_head = nn.Sequential(nn.Conv2d(16, 1, 3, padding=1))
        self._init_smp()

    def _init_smp(self):
        # smp.base.initialization.initialize_decoder / initialize_head
        for m in self.decoder.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        for m in self.segmentation_head.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        return self.segmentation_head(self.decoder(self.encoder(x)))


class OracleDiceLoss(nn.Module):
    """smp.losses.DiceLoss(mode="binary") defaults: from_logits=True, smooth=0, eps=1e-7, log_loss=False.

    One score for the whole batch (dims=(0,2)), zeroed when the batch has no positive pixel
    (call site: /root/reference/train.py:601).
    """

    def __init__(self, eps: float = 1e-7):
        super().__init__()
        self.eps = eps

    def forward(self, logits, target):
        bs = target.size(0)
        p = F.logsigmoid(logits).exp().view(bs, 1, -1)
        t = target.view(bs, 1, -1).type_as(p)
        inter = torch.sum(p * t, dim=(0, 2))
        card = torch.sum(p + t, dim=(0, 2))
        score = (2.0 * inter) / card.clamp_min(self.eps)
        loss = (1.0 - score) * (t.sum((0, 2)) > 0).to(p.dtype)
        return loss.mean()


def build_oracle(seed: int = 42) -> OracleUnet:
    """Random-init oracle, seeded like /root/reference/train.py:765 (`seed=42`)."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = OracleUnet()
    torch.random.set_rng_state(g)
    return m


def oracle_train_step(model, opt, x, y, dice=None):
    """One step of /root/reference/train.py:428-449 on CPU (autocast disabled there): returns loss value."""
    dice = dice or OracleDiceLoss()
    opt.zero_grad(set_to_none=True)
    logits = model(x)
    loss = F.binary_cross_entropy_with_logits(logits, y) + dice(logits, y)
    loss.backward()
    opt.step()
    return float(loss.detach())
