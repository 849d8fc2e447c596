"""Drop-in boundary on the CPU: the model files load exactly the way the
reference loads them, the fast configuration is the default, weight-pack invalidation reaches nested modules."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "frustum_convnet_b200")

# the body of import_from_file, /root/reference/utils/utils.py:12-25 (minus the copy into cfg.OUTPUT_DIR)
LOADER = r'''
import importlib, os, sys
def import_from_file(def_file):
    folder = os.path.dirname(def_file)
    file = os.path.basename(def_file)
    path = os.path.abspath(folder)
    sys.path.append(path)
    model_file = importlib.import_module(file[:-3])
    sys.path.remove(path)
    return model_file
assert not any(p.rstrip("/") == %(root)r for p in sys.path), "the repo root must not already be importable"
mod = import_from_file(%(file)r)
assert mod.__name__ == %(name)r and mod.__package__ in (None, ""), (mod.__name__, mod.__package__)
# outside the reference tree the cfg mirror is merged from our yaml (inside it: the reference's own cfg object)
from frustum_convnet_b200 import config
config.merge_cfg_from_file(%(yaml)r)
m = mod.PointNetDet(3, num_vec=%(nv)d)            # train_net_det.py:301-304
n = sum(p.numel() for p in m.parameters())
assert n == %(nparam)d, n
for cls in ("PointNetDet", "PointNetFeat", "PointNetModule", "ConvFeatNet", "QueryDepthPoint"):
    assert hasattr(mod, cls), cls
assert m.resolved_precision() == 1 and m.use_cuda_graph is True     # the benchmarked configuration is the default
print("OK", n)
'''


@pytest.mark.parametrize("fname,name,nv,nparam,yaml", [
    ("det_base.py", "det_base", 3, 3316777, "det_sample.yaml"),
    ("det_base_sunrgbd.py", "det_base_sunrgbd", 10, 6667589, "det_sample_sunrgbd.yaml"),
])
def test_model_file_loads_through_the_reference_loader(tmp_path, fname, name, nv, nparam, yaml):
    body = LOADER % dict(root=ROOT, file=os.path.join(PKG, fname), name=name, nv=nv, nparam=nparam,
                         yaml=os.path.join(ROOT, "cfgs", yaml))
    env = {k: v for k, v in os.environ.items() if k not in ("PYTHONPATH", "FCN_PRECISION", "FCN_CUDA_GRAPH")}
    r = subprocess.run([sys.executable, "-c", body], cwd=str(tmp_path), env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def _kitti_model():
    from frustum_convnet_b200 import config, synth
    from frustum_convnet_b200.det_base import PointNetDet
    cfg, w = config.load_workload("car")
    return PointNetDet(3, num_vec=3), w


def _clean(m):
    for sub in m.modules():
        if hasattr(sub, "_engine_dirty"):
            sub._engine_dirty = False


def test_pack_invalidation_reaches_nested_modules():
    m, w = _kitti_model()
    owners = [s for s in m.modules() if hasattr(s, "_engine_dirty")]
    assert len(owners) == 1 + 1 + 4 + 1                 # det, feat_net, 4 PointNetModules, conv_net

    class Wrapper(torch.nn.Module):                      # e.g. DataParallel-style ".module" checkpoints
        def __init__(self, inner):
            super().__init__()
            self.module = inner

    _clean(m)
    Wrapper(m).load_state_dict({"module." + k: v for k, v in m.state_dict().items()})
    assert all(o._engine_dirty for o in owners)
    _clean(m)
    m.conv_net.load_state_dict(m.conv_net.state_dict())
    assert m.conv_net._engine_dirty and not m.feat_net._engine_dirty
    _clean(m)
    m.train()
    assert all(o._engine_dirty for o in owners)
    _clean(m)
    m.float()
    assert all(o._engine_dirty for o in owners)
    # in-place updates bump Tensor._version, which the per-call scan compares
    v0 = m._param_version()
    with torch.no_grad():
        next(m.parameters()).add_(1.0)
    assert m._param_version() != v0


def test_defaults_follow_environment_and_replicas_raise(monkeypatch):
    from frustum_convnet_b200 import det_base
    monkeypatch.setenv("FCN_PRECISION", "0")
    monkeypatch.setenv("FCN_CUDA_GRAPH", "0")
    m, _ = _kitti_model()
    assert m.resolved_precision() == 0 and m.use_cuda_graph is False
    m.precision = 1
    assert m.resolved_precision() == 1
    m.eval()
    with pytest.raises(RuntimeError, match="CUDA only"):
        m.engine()                                       # CPU module: no fallback
    m._is_replica = True                                 # what torch.nn.parallel.replicate sets on replicas
    with pytest.raises(RuntimeError, match="DataParallel"):
        m.engine()
