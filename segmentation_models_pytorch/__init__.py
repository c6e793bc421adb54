"""Import shim: lets the reference scripts run unmodified (`import segmentation_models_pytorch as smp`) on the
B200-native implementation.  Only the two symbols the reference uses exist: `smp.Unet` and `smp.losses.DiceLoss`
(/root/reference/train.py:24,372,601; infer_pth_gui.py:6,32; ui_infer_rectangle.py:496; ui_infer_quadrilateral.py:638).
"""
from vickers_hardness_unet_b200 import Unet, losses  # noqa: F401

__version__ = "0.0.0+unet_b200"
