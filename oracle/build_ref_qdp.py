"""TEST INFRASTRUCTURE - builds oracle/_ref/libqdp_ref.so: the reference's own CUDA grouping kernel.

Recipe ("a patched THC-free copy is a legitimate secondary GPU-side oracle"):
  1. copy lines 14-65 (the DIVUP macro and the `query_depth_point_gpu<T>` kernel, nothing else) of
     /root/reference/ops/query_depth_point/query_depth_point_cuda_kernel.cu into
     oracle/_ref/qdp_ref_kernel.cuh  - git-ignored, never committed, travels to the GPU box with the snapshot;
  2. nvcc -gencode arch=compute_100a,code=sm_100a oracle/qdp_ref_wrap.cu -> oracle/_ref/libqdp_ref.so.
The kernel body is byte-identical to the reference; the un-buildable THC/ATen host wrapper (cu:68-86) is
replaced by the C launcher in qdp_ref_wrap.cu.  Runs only where /root/reference exists (the authoring
container); on the GPU box the prebuilt library is used as-is.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CU = "/root/reference/ops/query_depth_point/query_depth_point_cuda_kernel.cu"
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libqdp_ref.so")
FIRST, LAST = 14, 65


def build(force: bool = False):
    """Returns the library path, or None when the reference tree is absent and nothing was prebuilt."""
    if not os.path.exists(REF_CU):
        return LIB if os.path.exists(LIB) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    hdr = os.path.join(OUT_DIR, "qdp_ref_kernel.cuh")
    lines = open(REF_CU).read().split("\n")[FIRST - 1: LAST]
    text = "\n".join(lines) + "\n"
    assert "__global__ void query_depth_point_gpu" in text and text.rstrip().endswith("}"), \
        "reference kernel moved: check the line range"
    if not os.path.exists(hdr) or open(hdr).read() != text:
        open(hdr, "w").write(text)
        force = True
    wrap = os.path.join(HERE, "qdp_ref_wrap.cu")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(wrap):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-lineinfo", "-shared",
                               "-Xcompiler", "-fPIC", "-I", HERE, "-o", LIB, wrap])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
