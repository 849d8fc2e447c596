"""Pairwise rotated-box IoU on the device - drop-in for the reference's `rbbox_iou_3d_pair`
(ops/pybind11/box_ops.h:173-260; imported by models/det_base.py:26 and called at :495 on
`corner_preds.detach().cpu().numpy()`).  Here the corners stay on the GPU.

    overlap = rbbox_iou_3d_pair(corner_preds, corner_gts)        # (M, 2) CUDA tensor: [:, 0] BEV, [:, 1] 3-D
    overlap, stats = rbbox_iou_3d_pair(corner_preds, corner_gts, iou_thresh=cfg.IOU_THRESH)
    iou2d_mean, iou3d_mean, iou3d_gt_mean = stats               # 0-dim views, no host sync

No CPU fallback: CPU tensors raise, as every other entry of this package does.
"""
from __future__ import annotations

import torch

from . import _lib


def rbbox_iou_3d_pair(box_corners: torch.Tensor, qbox_corners: torch.Tensor, iou_thresh=None):
    if not (box_corners.is_cuda and qbox_corners.is_cuda):
        raise RuntimeError("rbbox_iou_3d_pair: CUDA tensors required (there is no CPU fallback)")
    if box_corners.dim() != 3 or tuple(box_corners.shape[1:]) != (8, 3) or \
            qbox_corners.dim() != 3 or tuple(qbox_corners.shape[1:]) != (8, 3):
        raise ValueError("rbbox_iou_3d_pair: corners must be (M, 8, 3)")
    M = box_corners.shape[0]
    dev = box_corners.device
    out = torch.zeros((M, 2), dtype=torch.float32, device=dev)
    stats = torch.zeros(3, dtype=torch.float32, device=dev) if iou_thresh is not None else None
    if M == qbox_corners.shape[0] and M > 0:     # N != K or N == 0 returns zeros (box_ops.h:201-203)
        c = box_corners.detach().to(torch.float32).contiguous()
        q = qbox_corners.detach().to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            _lib.call("fcn_rbbox_iou_3d_pair", M, c.data_ptr(), q.data_ptr(), out.data_ptr(),
                      float(iou_thresh if iou_thresh is not None else 0.0),
                      stats.data_ptr() if stats is not None else None,
                      torch.cuda.current_stream(dev).cuda_stream)
    if iou_thresh is None:
        return out
    return out, (stats[0], stats[1], stats[2])
