"""Drop-in for /root/reference/models/det_base_sunrgbd.py (SUN-RGBD, 5 scales).

Differences from the KITTI file (det_base_sunrgbd.py:115-128,178-200,278-279): five PointNet
scales with K=(128,128,256,256,256), block1 width 64, an extra block5 + 8x transposed conv, and
heads on 1024 channels.  Everything else is shared with ``det_base``.
"""
from __future__ import annotations

import os
import sys

# loadable as a top-level module through the reference's import_from_file (utils/utils.py:12-25)
_PKG_PARENT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_PARENT not in sys.path:
    sys.path.insert(0, _PKG_PARENT)

from frustum_convnet_b200 import det_base as _kitti  # noqa: E402
from frustum_convnet_b200.config import ARCH_SUNRGBD  # noqa: E402
from frustum_convnet_b200.det_base import PointNetModule, QueryDepthPoint, _block1d, _init_kaiming, _upblock1d  # noqa: E402

__all__ = ["QueryDepthPoint", "PointNetModule", "PointNetFeat", "ConvFeatNet", "PointNetDet"]


class PointNetFeat(_kitti.PointNetFeat):
    ARCH = ARCH_SUNRGBD

    def _engine_spec(self):
        return self.ARCH, self.num_vec, "SUNRGBD", self.dists, 12, "feat_net."


class ConvFeatNet(_kitti._EngineOwner):
    ARCH = ARCH_SUNRGBD

    def __init__(self, i_c=128, num_vec=10):
        super().__init__()
        self.num_vec = num_vec
        self.block1_conv1 = _block1d(i_c + num_vec, 64, 3, 1, 1)
        self.block2_conv1 = _block1d(64, 128, 3, 2, 1)
        self.block2_conv2 = _block1d(128, 128, 3, 1, 1)
        self.block2_merge = _block1d(128 + 128 + num_vec, 128, 1, 1)
        self.block3_conv1 = _block1d(128, 256, 3, 2, 1)
        self.block3_conv2 = _block1d(256, 256, 3, 1, 1)
        self.block3_merge = _block1d(256 + 256 + num_vec, 256, 1, 1)
        self.block4_conv1 = _block1d(256, 512, 3, 2, 1)
        self.block4_conv2 = _block1d(512, 512, 3, 1, 1)
        self.block4_merge = _block1d(512 + 512 + num_vec, 512, 1, 1)
        self.block5_conv1 = _block1d(512, 512, 3, 2, 1)
        self.block5_conv2 = _block1d(512, 512, 3, 1, 1)
        self.block5_merge = _block1d(512 + 512 + num_vec, 512, 1, 1)
        self.block5_deconv = _upblock1d(512, 256, 8, 8)
        self.block4_deconv = _upblock1d(512, 256, 4, 4)
        self.block3_deconv = _upblock1d(256, 256, 2, 2)
        self.block2_deconv = _upblock1d(128, 256, 1, 1)
        _init_kaiming(self)

    def _engine_spec(self):
        return self.ARCH, self.num_vec, "SUNRGBD", (0.0,) * self.ARCH.num_scales, 12, "conv_net."

    forward = _kitti.ConvFeatNet.forward


class PointNetDet(_kitti.PointNetDet):
    ARCH = ARCH_SUNRGBD
    FEAT_CLS = PointNetFeat
    FCN_CLS = ConvFeatNet
