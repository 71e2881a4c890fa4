"""B200-native forward/backward engine for the dual-head U-Net of
rg2/DeepFluoroLabeling-IPCAI2020 (drop-in for train_test_code/unet.py).

    import importlib
    dfl = importlib.import_module("deepfluorolabeling-ipcai2020_b200")
    net = dfl.UNet(n_classes=7, depth=6, wf=5, batch_norm=True, padding=True,
                   max_pool=False, num_lands=14).cuda()
"""
from .unet import UNet, UNetConvBlock, UNetUpBlock  # noqa: F401
from .util import center_crop  # noqa: F401
from . import parallel  # noqa: F401
from .losses import (DiceLoss2D, DiceAndHeatMapLoss2D, ncc_2d,  # noqa: F401
                     FusedDiceLoss2D, FusedDiceAndHeatMapLoss2D)
from .graphs import GraphedStep, GraphedForward  # noqa: F401
from . import prepost  # noqa: F401
from .build import build as build_library  # noqa: F401

__all__ = ["UNet", "UNetConvBlock", "UNetUpBlock", "center_crop", "parallel", "build_library",
           "DiceLoss2D", "DiceAndHeatMapLoss2D", "ncc_2d", "FusedDiceLoss2D", "FusedDiceAndHeatMapLoss2D", "GraphedStep", "GraphedForward", "prepost"]
