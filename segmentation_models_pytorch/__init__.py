"""Import shim: `import segmentation_models_pytorch as smp` in the reference scripts resolves to the B200-native
implementation when the repo root is put on PYTHONPATH *on purpose* (it shadows a real smp install, so do not leave the
repo root on a default import path).  Only the two symbols the reference uses exist: `smp.Unet` and
`smp.losses.DiceLoss` (/root/reference/train.py:24,372,601; infer_pth_gui.py:6,32; ui_infer_rectangle.py:496;
ui_infer_quadrilateral.py:638).  `encoder_weights="imagenet"` (train.py:753) works when a torchvision-format resnet34
checkpoint is on local disk ($UNETB200_RESNET34_WEIGHTS or torch hub's cache) and raises FileNotFoundError with that
hint otherwise; the loop functions `train_one_epoch` / `validate` of train.py run unmodified on this module
(tests/test_gpu_reference_loop.py).
"""
from vickers_hardness_unet_b200 import Unet, losses  # noqa: F401

__version__ = "0.0.0+unet_b200"
