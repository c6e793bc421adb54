"""The reference's OWN loop functions, unmodified, on the CUDA path (B200).

`/root/reference/train.py` does not exist on the GPU box, and reference sources are never committed: `build()` of
__graft_entry__.py stages an unmodified copy as `baseline/_ref/train.py` (git-ignored, travels with the snapshot — the
place the bench contract reserves for the unmodified reference).  This test imports THAT file with only its data
augmentation dependency (`albumentations`, not installed) stubbed, lets its `import segmentation_models_pytorch as smp`
resolve to the shim, and runs `build_model`, `train_one_epoch` (autocast fp16 + GradScaler + stock AdamW,
train.py:381-459) and `validate` (train.py:461-529) on batches of the reference's micrographs — once on the CUDA path
and once on the fp32 oracle, which must agree.  Skipped when the staged copy is absent.
"""
import importlib.util
import os
import sys
import types

import pytest
import torch

import vickers_hardness_unet_b200 as vb
from oracle import OracleDiceLoss, build_oracle

import vickers_data as vd

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TRAIN = os.path.join(ROOT, "baseline", "_ref", "train.py")


def _import_reference_train():
    if not os.path.exists(REF_TRAIN):
        pytest.skip("baseline/_ref/train.py not staged (run __graft_entry__.build() where /root/reference exists)")
    stub = types.ModuleType("albumentations")
    for n in ("Compose", "LongestMaxSize", "PadIfNeeded", "OneOf", "HorizontalFlip", "VerticalFlip", "RandomRotate90",
              "Rotate", "RandomBrightnessContrast", "CLAHE", "GaussianBlur", "GaussNoise", "Normalize"):
        setattr(stub, n, lambda *a, **k: None)
    sub = types.ModuleType("albumentations.pytorch")
    sub.ToTensorV2 = lambda *a, **k: None
    stub.pytorch = sub
    saved = {k: sys.modules.get(k) for k in ("albumentations", "albumentations.pytorch")}
    sys.modules["albumentations"], sys.modules["albumentations.pytorch"] = stub, sub
    threads = torch.get_num_threads()
    try:
        spec = importlib.util.spec_from_file_location("reference_train", REF_TRAIN)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)       # runs torch.set_num_threads(4) (train.py:19)
    finally:
        torch.set_num_threads(threads)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def test_reference_train_one_epoch_and_validate_run_unmodified():
    ref = _import_reference_train()
    import segmentation_models_pytorch as smp
    assert ref.smp is smp and smp.Unet is vb.Unet
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    data = vd.load_vickers()
    sched = vd.batches(data["train_u8"].shape[0], 8, 2, 5)
    train_loader = []
    for idx, ks in sched:
        x, y = vd.make_batch(data, idx, ks, "cpu")
        train_loader.append((x, y, [f"img{i}" for i in idx.tolist()]))       # (x, y, names) as VickersDataset yields
    val_loader = [(vd.normalise(data["val_u8"][i:i + 6]), data["val_y"][i:i + 6], ["v"] * 6) for i in (0, 6)]

    ref.set_seed(42)
    model = ref.build_model("resnet34", None).to("cuda")                     # train.py:357-378,595
    assert isinstance(model, vb.Unet)
    oracle = build_oracle(42)
    model.load_state_dict(oracle.state_dict(), strict=True)                  # identical init for the comparison
    oracle = oracle.cuda()
    from torch.amp import GradScaler
    out = {}
    for tag, mdl, dice in (("cuda", model, smp.losses.DiceLoss(mode="binary")), ("oracle", oracle, OracleDiceLoss())):
        bce = torch.nn.BCEWithLogitsLoss()                                   # train.py:600-601
        opt = torch.optim.AdamW(mdl.parameters(), lr=5e-5, weight_decay=1e-4)  # train.py:606, RECOMMENDED_CFG lr
        scaler = GradScaler("cuda", enabled=True)                            # train.py:610-611
        tl = ref.train_one_epoch(mdl, train_loader, opt, bce, dice, "cuda", scaler)
        vl, vdice, viou = ref.validate(mdl, val_loader, bce, dice, "cuda", out_vis_dir=None)
        out[tag] = (tl, vl, vdice, viou)
        print(f"\n[reference loop on {tag}] train_loss {tl:.4f} val_loss {vl:.4f} val_dice {vdice:.4f} val_iou {viou:.4f}")
    for a, b in zip(out["cuda"], out["oracle"]):
        assert a == a and abs(a - b) <= 0.03 * max(abs(b), 0.05), out      # the oracle arm ran under fp16 autocast
    assert model._ctx.device_error_flag() == 0
