// TEST INFRASTRUCTURE - secondary GPU-side oracle for the grouping op.
//
// Wraps the REFERENCE'S OWN kernel `query_depth_point_gpu<T>`
// (/root/reference/ops/query_depth_point/query_depth_point_cuda_kernel.cu:14-65), which oracle/build_ref_qdp.py
// extracts verbatim - at build time, from where it lies under /root/reference - into oracle/_ref/ (git-ignored;
// no reference source is committed).  The reference's host wrapper (cu:68-86) cannot be built any more
// (THC/THC.h was removed from PyTorch >= 2), so only its launch geometry is restated here:
// grid (DIVUP(m,256), b), 256 threads (cu:75-76).  Outputs must be pre-zeroed by the caller, as
// query_depth_point.py:36-39 does.  Only tests/ load the resulting library.
#include <cuda_runtime.h>
#include <cmath>

#include "_ref/qdp_ref_kernel.cuh"

extern "C" __attribute__((visibility("default"))) int qdp_ref_forward(int b, int n, int m, float dis_z, int nsample,
                                                                      const float *xyz1, const float *xyz2, long *idx,
                                                                      int *pts_cnt, void *stream) {
    if (b <= 0 || m <= 0) return 0;
    dim3 blocks(DIVUP(m, 256), b), threads(256);
    query_depth_point_gpu<float><<<blocks, threads, 0, (cudaStream_t)stream>>>(b, n, m, dis_z, nsample, xyz1, xyz2, idx,
                                                                                pts_cnt);
    return (int)cudaGetLastError();
}
