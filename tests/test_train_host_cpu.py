"""Host logic of the training path on the CPU (no CUDA): the static-shape loss formulation used for CUDA-graph capture
equals the reference-shaped one (values AND gradients w.r.t. the head logits), and the flat parameter bucket keeps
nn.Module / optimizer semantics."""
import numpy as np
import pytest
import torch

from frustum_convnet_b200 import config, synth


def _model():
    from frustum_convnet_b200.det_base import PointNetDet
    cfg, w = config.load_workload("refine_car")
    sd = synth.make_state_dict(w["arch"], 3, "KITTI", seed=11)
    m = PointNetDet(3, num_vec=3)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return m, cfg


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_masked_losses_equal_reference_shaped_losses(seed):
    """losses_masked (foreground WEIGHTS, static shapes) == losses_from_logits (nonzero() row selection, the
    reference's formulation, det_base.py:414-476): every loss term, the accuracies and d(total)/d(logits)."""
    from frustum_convnet_b200 import train_path as tp
    m, cfg = _model()
    m.gpu_iou_metrics = False                       # the IoU metric is a CUDA kernel; everything else is torch
    B = 5
    data = {k: torch.from_numpy(v) for k, v in synth.make_frustums("refine_car", B, seed=40 + seed, with_labels=True).items()}
    T2 = data["center_ref2"].shape[2]
    g = torch.Generator().manual_seed(seed)
    cls = torch.randn(B * T2, 2, generator=g).requires_grad_(True)
    reg = (torch.randn(B * T2, 39, generator=g) * 0.5).requires_grad_(True)
    la, ma = tp.losses_from_logits(m, cls, reg, data["center_ref2"], data)
    la["total_loss"].backward()
    ga = (cls.grad.clone(), reg.grad.clone())
    cls.grad = reg.grad = None
    ref2 = data["center_ref2"].permute(0, 2, 1).reshape(-1, 3)
    mean_size = torch.from_numpy(np.asarray(m.mean_size_array)).float()
    bidx = torch.arange(B * T2) // T2
    lb, mb = tp.losses_masked(m, cls, reg, ref2, data["cls_label"].reshape(-1), data["box3d_center"],
                              data["box3d_heading"], data["box3d_size"], data["size_class"], mean_size, bidx, None)
    lb["total_loss"].backward()
    for k in la:
        assert abs(float(la[k]) - float(lb[k])) <= 2e-6 * max(1.0, abs(float(la[k]))), (k, float(la[k]), float(lb[k]))
    for k in ("cls_acc", "head_acc", "size_acc"):
        assert abs(float(ma[k]) - float(mb[k])) < 1e-6, k
    for a, b in zip(ga, (cls.grad, reg.grad)):
        assert float((a - b).abs().max()) <= 1e-6 * max(1.0, float(a.abs().max()))


def test_flat_parameter_bucket_semantics():
    from frustum_convnet_b200.train_engine import FlatParams
    m, _ = _model()
    before = {k: v.detach().clone() for k, v in m.state_dict().items()}
    FlatParams.__init__.__globals__["torch"]          # (import check)
    m_cuda_like = m                                   # FlatParams asserts CUDA: emulate with a CPU variant
    ps = list(m_cuda_like.parameters())

    class CpuFlat(FlatParams):
        def __init__(self, module):
            ps_ = [p for p in module.parameters()]
            n = sum(p.numel() for p in ps_)
            self.param = torch.empty(n)
            self.grad = torch.zeros(n)
            self.offsets, off = {}, 0
            for p in ps_:
                k = p.numel()
                self.param[off: off + k].copy_(p.data.reshape(-1))
                p.data = self.param[off: off + k].view(p.shape)
                self.offsets[id(p)] = off
                off += k
            self.numel = n

    flat = CpuFlat(m)
    assert flat.numel == 3316777 == sum(p.numel() for p in ps)
    for k, v in m.state_dict().items():               # values unchanged, names unchanged
        assert torch.equal(v, before[k]), k
    # parameters are views of the bucket: an in-place update of the bucket is what the module sees
    first = next(m.parameters())
    flat.param[:first.numel()].add_(1.0)
    assert torch.equal(first.detach().reshape(-1), before[next(iter(before))].reshape(-1) + 1.0)
    # gradient views tile the bucket without gaps, in parameter order (feat_net first, conv_net + heads last)
    off = 0
    for p in m.parameters():
        assert flat.offsets[id(p)] == off
        assert flat.grad_view(p).shape == p.shape
        off += p.numel()
    split = flat.offsets[id(next(m.conv_net.parameters()))]
    assert 0 < split < flat.numel and split == sum(p.numel() for p in m.feat_net.parameters())
    # a stock optimizer built on the SAME Parameter objects keeps working after the re-pointing
    flat.expose_grads(m)
    flat.grad.fill_(0.5)
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    w0 = flat.param.clone()
    opt.step()
    assert torch.allclose(flat.param, w0 - 0.05)
