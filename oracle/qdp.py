"""ORACLE (test infrastructure) — grouping op ``query_depth_point`` on the CPU.

Three independent restatements of
/root/reference/ops/query_depth_point/query_depth_point_cuda_kernel.cu:16-65
behind the Python calling convention of
/root/reference/ops/query_depth_point/query_depth_point.py:12-40:

* ``qdp_c``      — the plain-C loop (oracle/query_depth_point_ref.c) via ctypes;
* ``qdp_loops``  — a literal pure-Python triple loop (small cases only);
* ``qdp_numpy``  — a vectorised numpy formulation (stable sort of hit flags).

All take channel-first float32 arrays xyz1 (B,3,N), xyz2 (B,3,M) and return
(idx int64 (B,M,K), cnt int32 (B,M)).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle_qdp.so")
    src = os.path.join(_HERE, "query_depth_point_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
        for name in ("oracle_query_depth_point_bn3", "oracle_query_depth_point_b3n"):
            fn = getattr(_LIB, name)
            fn.restype = None
            fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return _LIB


def qdp_c(xyz1: np.ndarray, xyz2: np.ndarray, dis_z: float, nsample: int, transposed_call=False):
    xyz1 = np.ascontiguousarray(xyz1, dtype=np.float32)
    xyz2 = np.ascontiguousarray(xyz2, dtype=np.float32)
    B, _, N = xyz1.shape
    M = xyz2.shape[2]
    idx = np.empty((B, M, nsample), dtype=np.int64)
    cnt = np.empty((B, M), dtype=np.int32)
    if transposed_call:  # exactly what query_depth_point.py:29-39 hands to the kernel
        a = np.ascontiguousarray(xyz1.transpose(0, 2, 1))
        b = np.ascontiguousarray(xyz2.transpose(0, 2, 1))
        _lib().oracle_query_depth_point_bn3(B, N, M, dis_z, nsample, a.ctypes.data, b.ctypes.data,
                                            idx.ctypes.data, cnt.ctypes.data)
    else:
        _lib().oracle_query_depth_point_b3n(B, N, M, dis_z, nsample, xyz1.ctypes.data,
                                            xyz2.ctypes.data, idx.ctypes.data, cnt.ctypes.data)
    return idx, cnt


def qdp_loops(xyz1, xyz2, dis_z, nsample):
    xyz1 = np.asarray(xyz1, dtype=np.float32)
    xyz2 = np.asarray(xyz2, dtype=np.float32)
    B, _, N = xyz1.shape
    M = xyz2.shape[2]
    dz = np.float32(dis_z)
    idx = np.zeros((B, M, nsample), dtype=np.int64)
    cnt = np.zeros((B, M), dtype=np.int32)
    for b in range(B):
        for m in range(M):
            c = 0
            z2 = xyz2[b, 2, m]
            for k in range(N):
                if c == nsample:
                    break
                d3 = np.abs(np.float32(z2 - xyz1[b, 2, k]))
                if d3 < dz:
                    if c == 0:
                        idx[b, m, :] = k
                    idx[b, m, c] = k
                    c += 1
            cnt[b, m] = c
    return idx, cnt


def qdp_numpy(xyz1, xyz2, dis_z, nsample):
    xyz1 = np.asarray(xyz1, dtype=np.float32)
    xyz2 = np.asarray(xyz2, dtype=np.float32)
    B, _, N = xyz1.shape
    z1 = xyz1[:, 2, :]                      # (B,N)
    z2 = xyz2[:, 2, :]                      # (B,M)
    d = np.abs((z2[:, :, None] - z1[:, None, :]).astype(np.float32))
    hit = d < np.float32(dis_z)             # (B,M,N)  NaN -> False
    total = hit.sum(-1)
    cnt = np.minimum(total, nsample).astype(np.int32)
    # stable order: hits first, in index order
    order = np.argsort(~hit, axis=-1, kind="stable")[:, :, :nsample]
    if order.shape[-1] < nsample:           # N < nsample
        pad = np.zeros(order.shape[:-1] + (nsample - order.shape[-1],), dtype=order.dtype)
        order = np.concatenate([order, pad], -1)
    first = order[:, :, :1]
    k = np.arange(nsample)[None, None, :]
    idx = np.where(k < cnt[:, :, None], order, first)
    idx = np.where(cnt[:, :, None] > 0, idx, 0).astype(np.int64)
    return idx, cnt
