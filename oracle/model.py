"""ORACLE (test infrastructure) — fp32 CPU restatement of the reference forward.

Functional (state-dict driven) restatement, with torch CPU fp32 ops, of
/root/reference/models/det_base.py:
  PointNetModule.forward  :62-103   -> ``pointnet_module``
  PointNetFeat.forward    :126-159  -> ``pointnet_feat``
  ConvFeatNet.forward     :196-224  -> ``conv_feat_net``
  PointNetDet.forward     :334-412  -> ``pointnet_det_eval`` (eval branch)
and of the 5-scale variant /root/reference/models/det_base_sunrgbd.py:115-252.
Layer factories follow models/common.py:38-63 (conv(bias=False) -> BN(eps=1e-5) -> ReLU).
Box decode follows models/box_transform.py:5-12,28-41.

The grouping op is ``oracle.qdp.qdp_c``.  The conv/BN arithmetic itself lives in
PyTorch (third-party, not under /root/reference); parity for it is pinned by the
golden fixtures in tests/golden/, which were produced by importing and running
the reference's own modules (oracle/make_golden.py) — this file must reproduce
those fixtures (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import qdp

EPS = 1e-5


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))


def to_torch_state(sd):
    return {k: _t(v) for k, v in sd.items()}


def _bn(x, sd, p, training=False):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], training, 0.1, EPS)


def conv2d_bn_relu(x, sd, p):
    return F.relu(_bn(F.conv2d(x, sd[p + ".0.weight"]), sd, p + ".1"))


def conv1d_bn_relu(x, sd, p, stride=1, pad=0):
    return F.relu(_bn(F.conv1d(x, sd[p + ".0.weight"], None, stride, pad), sd, p + ".1"))


def deconv1d_bn_relu(x, sd, p, stride):
    return F.relu(_bn(F.conv_transpose1d(x, sd[p + ".0.weight"], None, stride, 0), sd, p + ".1"))


def query_depth_point(pc, new_pc, dist, nsample):
    idx, cnt = qdp.qdp_c(pc.numpy(), new_pc.numpy(), float(dist), int(nsample))
    return torch.from_numpy(idx), torch.from_numpy(cnt)


def pointnet_module(pc, new_pc, sd, prefix, dist, nsample):
    """det_base.py:62-103 (use_xyz=True, no extra features). Returns ((B,C3,T,K), idx, cnt)."""
    B = pc.shape[0]
    T = new_pc.shape[2]
    idx, cnt = query_depth_point(pc, new_pc, dist, nsample)
    g = torch.gather(pc, 2, idx.view(B, 1, T * nsample).expand(-1, 3, -1)).view(B, 3, T, nsample)
    g = g - new_pc.unsqueeze(3)
    x = conv2d_bn_relu(g, sd, prefix + ".conv1")
    x = conv2d_bn_relu(x, sd, prefix + ".conv2")
    x = conv2d_bn_relu(x, sd, prefix + ".conv3")
    valid = (cnt > 0).view(B, 1, -1, 1)
    return x * valid.float(), idx, cnt


def pointnet_feat(pc, centers, one_hot, sd, dists, nsamples, prefix="feat_net", keep_groups=False):
    """det_base.py:126-159. Returns list of (B, C3+V, T_i) and the per-scale (idx,cnt)."""
    feats, groups = [], []
    for i, (c, d, k) in enumerate(zip(centers, dists, nsamples)):
        x, idx, cnt = pointnet_module(pc, c, sd, "%s.pointnet%d" % (prefix, i + 1), d, k)
        f = x.max(-1)[0]
        if one_hot is not None:
            f = torch.cat([f, one_hot.unsqueeze(-1).expand(-1, -1, f.shape[-1])], 1)
        feats.append(f)
        groups.append((idx, cnt))
    return feats, groups


def conv_feat_net(feats, sd, prefix="conv_net"):
    """det_base.py:196-224 / det_base_sunrgbd.py:213-252 for len(feats) in (4, 5)."""
    S = len(feats)
    x = conv1d_bn_relu(feats[0], sd, prefix + ".block1_conv1", 1, 1)
    branches = []
    for i in range(2, S + 1):
        x = conv1d_bn_relu(x, sd, "%s.block%d_conv1" % (prefix, i), 2, 1)
        x = conv1d_bn_relu(x, sd, "%s.block%d_conv2" % (prefix, i), 1, 1)
        x = torch.cat([x, feats[i - 1]], 1)
        x = conv1d_bn_relu(x, sd, "%s.block%d_merge" % (prefix, i), 1, 0)
        branches.append(x)
    ups = [deconv1d_bn_relu(b, sd, "%s.block%d_deconv" % (prefix, i + 2), 2 ** i)
           for i, b in enumerate(branches)]
    L = ups[0].shape[-1]
    return torch.cat([u[:, :, :L] for u in ups], 1)


def heads(x, sd):
    """det_base.py:367-374: returns logits rows (B*T2, 2) and (B*T2, out)."""
    cls = F.conv1d(x, sd["cls_out.weight"], sd["cls_out.bias"])
    reg = F.conv1d(x, sd["reg_out.weight"], sd["reg_out.bias"])
    cls = cls.permute(0, 2, 1).contiguous().view(-1, 2)
    reg = reg.permute(0, 2, 1).contiguous().view(-1, reg.shape[1])
    return cls, reg


def decode_eval(cls_scores, outputs, center_ref2, mean_size, B, num_bins=12):
    """det_base.py:376-412 + box_transform.py:5-12,28-41. Returns the eval 6-tuple."""
    S = mean_size.shape[0]
    ref2 = center_ref2.permute(0, 2, 1).contiguous().view(-1, 3)
    cls_probs = F.softmax(cls_scores, -1)
    center = outputs[:, 0:3]
    h_scores = outputs[:, 3:3 + num_bins]
    h_res = outputs[:, 3 + num_bins:3 + 2 * num_bins]
    s_scores = outputs[:, 3 + 2 * num_bins:3 + 2 * num_bins + S]
    s_res = outputs[:, 3 + 2 * num_bins + S:].contiguous().view(-1, S, 3)
    h_probs = F.softmax(h_scores, -1)
    s_probs = F.softmax(s_scores, -1)
    h_lab = torch.argmax(h_probs, -1)
    s_lab = torch.argmax(s_probs, -1)
    center_preds = center + ref2
    apc = 2 * np.pi / float(num_bins)
    ang = h_lab.float() * apc + torch.gather(h_res, 1, h_lab.unsqueeze(1)).squeeze(1) * (apc / 2)
    ang = torch.where(ang > np.pi, ang - 2 * np.pi, ang)
    ms = mean_size.type_as(outputs)
    ex = ms[s_lab]
    off = torch.gather(s_res, 1, s_lab.view(-1, 1, 1).expand(-1, -1, 3)).squeeze(1)
    size = off * ex + ex
    return (cls_probs.view(B, -1, 2), center_preds.view(B, -1, 3), ang.view(B, -1),
            size.view(B, -1, 3), h_probs.view(B, -1, num_bins), s_probs.view(B, -1, S))


def pointnet_det_eval(data, sd, dists, nsamples, mean_size, num_bins=12, return_all=False):
    """Whole eval forward. ``data``: dict of numpy/torch arrays keyed as det_base.py:336-347."""
    with torch.no_grad():
        pc = _t(data["point_cloud"])[:, :3, :].contiguous()
        S = len(dists)
        centers = [_t(data["center_ref%d" % (i + 1)]) for i in range(S)]
        one_hot = _t(data["one_hot"]) if data.get("one_hot") is not None else None
        feats, groups = pointnet_feat(pc, centers, one_hot, sd, dists, nsamples)
        x = conv_feat_net(feats, sd)
        cls, reg = heads(x, sd)
        out = decode_eval(cls, reg, centers[1], _t(np.asarray(mean_size)), pc.shape[0], num_bins)
        if return_all:
            return dict(groups=groups, feats=feats, x=x, cls=cls, reg=reg, out=out)
        return out
