"""ORACLE (test infrastructure) — import the UNMODIFIED reference model in the
authoring container so that golden fixtures come from the reference's own code.

/root/reference exists only in the authoring container (never on the GPU box), so
this module is used exclusively by oracle/make_golden.py.  Three shims are needed
because two compiled extensions cannot be built here and PyYAML >= 6 changed
``yaml.load``:
  * ``ops.query_depth_point.query_depth_point`` -> a CPU ``QueryDepthPoint`` backed by
    the C restatement (the shipped one asserts ``.is_cuda`` and imports a missing .so,
    query_depth_point.py:6,23-24);
  * ``ops.pybind11.box_ops_cc.rbbox_iou_3d_pair`` -> zeros (needs Boost; metrics only,
    det_base.py:495);
  * ``yaml.load(s)`` -> ``yaml.load(s, Loader=yaml.FullLoader)`` (configs/config.py:228).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import yaml

from . import qdp

REF = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF, "models"))


def _install_shims():
    if "ops.query_depth_point.query_depth_point" in sys.modules:
        return

    class QueryDepthPoint(torch.nn.Module):
        def __init__(self, dis_z, nsample):
            super().__init__()
            self.dis_z, self.nsample = dis_z, nsample

        def forward(self, xyz1, xyz2):
            idx, cnt = qdp.qdp_c(xyz1.detach().numpy(), xyz2.detach().numpy(), self.dis_z,
                                 self.nsample, transposed_call=True)
            return torch.from_numpy(idx), torch.from_numpy(cnt)

    for name in ("ops", "ops.query_depth_point", "ops.pybind11"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    m = types.ModuleType("ops.query_depth_point.query_depth_point")
    m.QueryDepthPoint = QueryDepthPoint
    sys.modules[m.__name__] = m
    m = types.ModuleType("ops.pybind11.box_ops_cc")
    m.rbbox_iou_3d_pair = lambda a, b: np.zeros((a.shape[0], 2), dtype=np.float32)
    sys.modules[m.__name__] = m
    _orig = yaml.load
    yaml.load = lambda s, Loader=None: _orig(s, Loader=Loader or yaml.FullLoader)
    if REF not in sys.path:
        sys.path.insert(0, REF)


def load_reference_model(yaml_name: str, num_vec: int):
    """Returns (model, cfg) of the reference for ``cfgs/<yaml_name>`` (fresh cfg each call)."""
    _install_shims()
    for k in [k for k in sys.modules if k.startswith(("configs", "models", "datasets"))]:
        del sys.modules[k]
    from configs.config import cfg, merge_cfg_from_file  # type: ignore
    merge_cfg_from_file(os.path.join(REF, "cfgs", yaml_name))
    import importlib
    mod = importlib.import_module(cfg.MODEL.FILE[:-3].replace("/", "."))
    model = mod.PointNetDet(3, num_vec=num_vec, num_classes=cfg.MODEL.NUM_CLASSES)
    return model, cfg
