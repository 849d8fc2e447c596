"""GPU: training branch (refine_car.yaml).  Grouping runs on libfrustum_b200, the
differentiable arithmetic is PyTorch autograd on the drop-in modules' own parameters; losses,
accuracies, gradients and the BN running-stat update must match the reference's train-mode forward
(fixture produced by the unmodified reference on CPU, oracle/make_golden.py::run_train_case)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


def test_train_forward_backward_matches_reference_golden():
    from frustum_convnet_b200 import config, synth
    g = dict(np.load(os.path.join(GOLDEN_DIR, "refine_car_train_b4.npz")))
    cfg, w = config.load_workload("refine_car")
    from frustum_convnet_b200.det_base import PointNetDet
    sd = synth.make_state_dict(w["arch"], 3, "KITTI", seed=11)
    m = PointNetDet(3, num_vec=3)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    m = m.cuda().train()
    m.train_kernels = False                          # this test pins the PyTorch-autograd composition of the branch
    data = synth.make_frustums("refine_car", 4, seed=206, with_labels=True)
    chk = float(sum(np.asarray(v, dtype=np.float64).sum() for v in data.values()))
    assert abs(chk - float(g["input_checksum"])) < 1e-6 * max(1.0, abs(chk))
    torch.backends.cudnn.allow_tf32 = False          # fp32 parity for the autograd path
    torch.backends.cuda.matmul.allow_tf32 = False
    losses, metrics = m({k: torch.from_numpy(v).cuda() for k, v in data.items()})
    losses["total_loss"].backward()
    for k, v in losses.items():
        ref = float(g["loss_" + k])
        assert abs(float(v) - ref) <= 2e-4 * max(1.0, abs(ref)), (k, float(v), ref)
    for k in ("cls_acc", "head_acc", "size_acc"):
        assert abs(float(metrics[k]) - float(g["metric_" + k])) < 1e-6
    params = dict(m.named_parameters())
    for key in [k for k in g if k.startswith("grad_")]:
        gr = params[key[5:]].grad.cpu().numpy()
        ref = g[key]
        assert np.abs(gr - ref).max() <= 2e-3 * max(1e-3, np.abs(ref).max()), key
    rm = m.feat_net.pointnet1.conv1[1].running_mean.cpu().numpy()
    assert np.abs(rm - g["bn_running_mean"]).max() < 1e-5
    # and the same module switches back to the kernel path in eval mode
    m.eval()
    out = m({k: torch.from_numpy(v).cuda() for k, v in data.items() if not k.startswith(("cls_", "box3d", "size_c"))})
    assert len(out) == 6 and all(torch.isfinite(o).all() for o in out)


def test_reference_style_training_loop_on_the_kernels():
    """The reference's loop (train_net_det.py:121-128) unchanged: model(data) -> loss.backward() -> optimizer.step(),
    with the drop-in module in its DEFAULT train configuration (hand-written kernels behind an autograd.Function);
    gradients must equal the reference fixture, a torch optimizer must be able to step."""
    from frustum_convnet_b200 import config, synth
    g = dict(np.load(os.path.join(GOLDEN_DIR, "refine_car_train_b4.npz")))
    cfg, w = config.load_workload("refine_car")
    from frustum_convnet_b200.det_base import PointNetDet
    sd = synth.make_state_dict(w["arch"], 3, "KITTI", seed=11)
    m = PointNetDet(3, num_vec=3)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    m = m.cuda().train()
    assert m.train_kernels
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-4)
    data = {k: torch.from_numpy(v).cuda() for k, v in synth.make_frustums("refine_car", 4, seed=206, with_labels=True).items()}
    opt.zero_grad()
    losses, metrics = m(data)
    losses["total_loss"].backward()
    for k, v in losses.items():
        ref = float(g["loss_" + k])
        assert abs(float(v.detach()) - ref) <= 2e-4 * max(1.0, abs(ref)), (k, float(v.detach()), ref)
    params = dict(m.named_parameters())
    for key in [k for k in g if k.startswith("grad_")]:
        gr = params[key[5:]].grad.cpu().numpy()
        assert np.abs(gr - g[key]).max() <= 2e-3 * max(1e-3, np.abs(g[key]).max()), key
    before = params["conv_net.block2_merge.0.weight"].detach().clone()
    opt.step()
    assert float((params["conv_net.block2_merge.0.weight"] - before).abs().max()) > 0
    losses2, _ = m(data)                               # second step: engine reuse, updated weights
    assert torch.isfinite(losses2["total_loss"])
