"""ORACLE (test infrastructure) — generate tests/golden/*.npz FROM THE REFERENCE ITSELF.

Run in the authoring container only (needs /root/reference):
    python -m oracle.make_golden
For each case the UNMODIFIED reference ``PointNetDet`` (models/det_base.py or
models/det_base_sunrgbd.py, imported through oracle/ref_import.py) is loaded with the
seeded state dict of ``frustum_convnet_b200.synth.make_state_dict``, put in eval mode and
run on the seeded synthetic frustums of ``synth.make_frustums``; forward hooks capture
the intermediate tensors.  Inputs are NOT stored (they are regenerated from the seed; a
float64 checksum guards against generator drift); outputs are.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from frustum_convnet_b200 import config, synth  # noqa: E402
from oracle import ref_import  # noqa: E402

# name: (workload, B, data seed, weight seed, max_depth override, N override)
CASES = {
    "car_full_b1": ("car", 1, 101, 7, None, None),
    "car_small_b3": ("car", 3, 102, 8, 17.5, None),
    "car_oddn_b2": ("car", 2, 103, 8, 17.5, 777),
    "people_small_b2": ("people", 2, 104, 9, 7.0, None),
    "sunrgbd_full_b2": ("sunrgbd", 2, 105, 10, None, None),
    "refine_car_b4": ("refine_car", 4, 106, 11, None, None),
}


def checksum(d):
    return float(sum(np.asarray(v, dtype=np.float64).sum() for v in d.values()))


def run_case(name):
    workload, B, dseed, wseed, md, N = CASES[name]
    w = config.WORKLOADS[workload]
    arch = w["arch"]
    model, cfg = ref_import.load_reference_model(w["yaml"], w["num_vec"])
    dataset = cfg.DATA.DATASET_NAME
    sd = synth.make_state_dict(arch, w["num_vec"], dataset, seed=wseed)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    model.eval()
    data = synth.make_frustums(workload, B, seed=dseed, max_depth=md, N=N)
    cap = {}
    hooks = []
    S = arch.num_scales
    for i in range(S):
        q = getattr(model.feat_net, "pointnet%d" % (i + 1)).query_depth_point
        hooks.append(q.register_forward_hook(
            lambda m, inp, out, i=i: cap.__setitem__("group%d" % (i + 1), out)))
    hooks.append(model.feat_net.register_forward_hook(lambda m, i, o: cap.__setitem__("feats", o)))
    hooks.append(model.conv_net.register_forward_hook(lambda m, i, o: cap.__setitem__("x", o)))
    hooks.append(model.cls_out.register_forward_hook(lambda m, i, o: cap.__setitem__("cls", o)))
    hooks.append(model.reg_out.register_forward_hook(lambda m, i, o: cap.__setitem__("reg", o)))
    with torch.no_grad():
        out = model({k: torch.from_numpy(v) for k, v in data.items()})
    for h in hooks:
        h.remove()
    rec = {"input_checksum": np.array(checksum(data)),
           "meta": np.array([B, dseed, wseed, -1 if md is None else md, -1 if N is None else N],
                            dtype=np.float64)}
    for i in range(S):
        idx, cnt = cap["group%d" % (i + 1)]
        assert idx.max() < 32768
        rec["idx%d" % (i + 1)] = idx.numpy().astype(np.int16)
        rec["cnt%d" % (i + 1)] = cnt.numpy().astype(np.int16)
        rec["feat%d" % (i + 1)] = cap["feats"][i].numpy()
    rec["x"] = cap["x"].numpy()
    rec["cls"] = cap["cls"].numpy()   # (B,2,T2)
    rec["reg"] = cap["reg"].numpy()   # (B,out,T2)
    for j, o in enumerate(out):
        rec["out%d" % j] = o.numpy()
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **rec)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def run_train_case(name="refine_car_train_b4"):
    """Training branch (models/det_base.py:414-525): losses/metrics + a few gradients of the UNMODIFIED
    reference in train mode (batch-statistics BN) on seeded inputs with labels."""
    workload, B, dseed, wseed = "refine_car", 4, 206, 11
    w = config.WORKLOADS[workload]
    model, cfg = ref_import.load_reference_model(w["yaml"], w["num_vec"])
    sd = synth.make_state_dict(w["arch"], w["num_vec"], cfg.DATA.DATASET_NAME, seed=wseed)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    model.train()
    data = synth.make_frustums(workload, B, seed=dseed, with_labels=True)
    losses, metrics = model({k: torch.from_numpy(v) for k, v in data.items()})
    losses["total_loss"].backward()
    rec = {"input_checksum": np.array(checksum(data))}
    for k, v in losses.items():
        rec["loss_" + k] = v.detach().numpy()
    for k in ("cls_acc", "head_acc", "size_acc"):
        rec["metric_" + k] = metrics[k].detach().numpy()
    for pn in ("cls_out.weight", "reg_out.bias", "feat_net.pointnet1.conv1.0.weight",
               "conv_net.block4_merge.1.weight", "feat_net.pointnet4.conv3.1.bias"):
        rec["grad_" + pn] = dict(model.named_parameters())[pn].grad.numpy()
    rec["bn_running_mean"] = model.feat_net.pointnet1.conv1[1].running_mean.numpy()
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **rec)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    assert ref_import.reference_available(), "needs /root/reference (authoring container only)"
    for n in (sys.argv[1:] or CASES):
        if n != "train":
            run_case(n)
    if not sys.argv[1:] or "train" in sys.argv[1:]:
        run_train_case()
