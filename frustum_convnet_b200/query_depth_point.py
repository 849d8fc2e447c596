"""Drop-in for /root/reference/ops/query_depth_point/query_depth_point.py.

Same public surface: ``QueryDepthPoint(dis_z, nsample)(xyz1 (B,3,N), xyz2 (B,3,M)) ->
(idx int64 (B,M,nsample), pts_cnt int32 (B,M))`` with the same assertions
(query_depth_point.py:23-27) and the same non-differentiable contract (:42-44), but backed by
``fcn_query_depth_point_b3n`` of libfrustum_b200.so: channel-first input is consumed directly
(no permute().contiguous() copies, no zero-fill memsets).
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib


def query_depth_point(dis_z: float, nsample: int, xyz1: torch.Tensor, xyz2: torch.Tensor):
    assert xyz1.is_cuda and xyz1.size(1) == 3
    assert xyz2.is_cuda and xyz2.size(1) == 3
    assert xyz1.size(0) == xyz2.size(0)
    assert xyz1.is_contiguous()
    assert xyz2.is_contiguous()
    assert xyz1.dtype == torch.float32 and xyz2.dtype == torch.float32
    b, n, m = xyz1.size(0), xyz1.size(2), xyz2.size(2)
    with torch.cuda.device_of(xyz1):
        idx = torch.empty((b, m, nsample), dtype=torch.int64, device=xyz1.device)
        pts_cnt = torch.empty((b, m), dtype=torch.int32, device=xyz1.device)
        stream = torch.cuda.current_stream().cuda_stream
        _lib.call("fcn_query_depth_point_b3n", b, n, m, float(dis_z), int(nsample),
                  xyz1.data_ptr(), xyz2.data_ptr(), idx.data_ptr(), pts_cnt.data_ptr(), stream)
    return idx, pts_cnt


def query_depth_point_bn3(dis_z, nsample, xyz1_bn3, xyz2_bm3, idx, pts_cnt):
    """The reference's *native* entry (query_depth_point_cuda.forward, cpp:25-45): caller-owned
    outputs, (b,n,3)/(b,m,3) inputs."""
    b, n, m = xyz1_bn3.size(0), xyz1_bn3.size(1), xyz2_bm3.size(1)
    assert xyz1_bn3.is_cuda and xyz1_bn3.is_contiguous() and xyz2_bm3.is_contiguous()
    assert idx.dtype == torch.int64 and pts_cnt.dtype == torch.int32
    with torch.cuda.device_of(xyz1_bn3):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.call("fcn_query_depth_point_bn3", b, n, m, float(dis_z), int(nsample),
                  xyz1_bn3.data_ptr(), xyz2_bm3.data_ptr(), idx.data_ptr(), pts_cnt.data_ptr(),
                  stream)


class QueryDepthPoint(nn.Module):
    def __init__(self, dis_z, nsample):
        super().__init__()
        self.dis_z = dis_z
        self.nsample = nsample

    @torch.no_grad()
    def forward(self, xyz1, xyz2):
        return query_depth_point(self.dis_z, self.nsample, xyz1.detach(), xyz2.detach())
