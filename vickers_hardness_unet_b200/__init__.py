"""B200-native U-Net(ResNet-34) hot path of ZooMEISTER/vickers-hardness-Unet.

Public surface mirrors the two smp symbols the reference uses (`Unet`, `losses.DiceLoss`), see SURVEY.md section 8b.
"""
from ._lib import UnetB200Error, LIB_PATH  # noqa: F401
from .unet import Unet  # noqa: F401
from . import losses  # noqa: F401
from . import distributed  # noqa: F401
from . import metrics  # noqa: F401
from . import data  # noqa: F401
from .optim import FusedAdamW  # noqa: F401

__all__ = ["Unet", "losses", "distributed", "metrics", "data", "FusedAdamW", "UnetB200Error"]
