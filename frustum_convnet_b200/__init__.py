"""frustum_convnet_b200 — Blackwell-native (sm_100a) per-frustum hot path of F-ConvNet.

Only the path named in BASELINE.json is built here: sliding-frustum grouping
(``QueryDepthPoint``), the grouped shared-MLP + max-over-K extractor
(``PointNetModule`` / ``PointNetFeat``), the 1-D FCN (``ConvFeatNet``) and the
detection heads + decode (``PointNetDet``).  All device work goes through the
C-ABI library ``libfrustum_b200.so`` (``include/frustum_b200.h``); there is no
CPU fallback — importing the compute modules without the built library raises.
"""
__version__ = "0.1.0"
