// Train-metric kernels for the pairwise rotated-box IoU: one thread per box pair, then a
// single-CTA, fixed-order reduction to the three scalars models/det_base.py:497-500 logs
// (mean BEV IoU, mean 3-D IoU, fraction of pairs with 3-D IoU >= threshold).  Everything stays on the
// device: no device->host copy of the boxes, no host synchronisation.
#include "box_iou.cuh"
#include "common.cuh"

namespace fcn {

constexpr int IOU_THREADS = 128;

__global__ void __launch_bounds__(IOU_THREADS)
rbbox_iou_pair_kernel(int M, const float *__restrict__ corners, const float *__restrict__ qcorners,
                      float *__restrict__ iou) {
    const int n = blockIdx.x * IOU_THREADS + threadIdx.x;
    if (n >= M) return;
    float c[24], q[24], o[2];
    const float4 *cs = (const float4 *)(corners + (size_t)n * 24);    // 96 B per box: 6 aligned float4
    const float4 *qs = (const float4 *)(qcorners + (size_t)n * 24);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float4 a = __ldg(cs + i), b = __ldg(qs + i);
        c[4 * i] = a.x; c[4 * i + 1] = a.y; c[4 * i + 2] = a.z; c[4 * i + 3] = a.w;
        q[4 * i] = b.x; q[4 * i + 1] = b.y; q[4 * i + 2] = b.z; q[4 * i + 3] = b.w;
    }
    rbbox_iou_pair(c, q, o);
    *(float2 *)(iou + 2 * (size_t)n) = make_float2(o[0], o[1]);
}

// stats[0] = mean iou2d, stats[1] = mean iou3d, stats[2] = mean(iou3d >= thresh); deterministic order
__global__ void __launch_bounds__(256)
rbbox_iou_stats_kernel(int M, const float *__restrict__ iou, float thresh, float *__restrict__ stats) {
    __shared__ float s[3][256];
    float a = 0.f, b = 0.f, c = 0.f;
    for (int n = threadIdx.x; n < M; n += 256) {
        const float2 v = *(const float2 *)(iou + 2 * (size_t)n);
        a += v.x;
        b += v.y;
        c += v.y >= thresh ? 1.f : 0.f;
    }
    s[0][threadIdx.x] = a;
    s[1][threadIdx.x] = b;
    s[2][threadIdx.x] = c;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w)
            for (int k = 0; k < 3; ++k) s[k][threadIdx.x] += s[k][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x < 3) stats[threadIdx.x] = M > 0 ? s[threadIdx.x][0] / (float)M : 0.f;
}

}  // namespace fcn

using namespace fcn;

extern "C" int fcn_rbbox_iou_3d_pair(int M, const float *corners, const float *qcorners, float *iou,
                                     float iou_thresh, float *stats, fcn_stream_t stream) {
    FCN_REQUIRE(M >= 0, "negative pair count");
    if (M > 0) {
        FCN_REQUIRE(corners && qcorners && iou, "NULL pointer");
        FCN_REQUIRE(((uintptr_t)corners & 15) == 0 && ((uintptr_t)qcorners & 15) == 0 && ((uintptr_t)iou & 7) == 0,
                    "corner arrays must be 16-byte aligned, iou 8-byte aligned");
        rbbox_iou_pair_kernel<<<ceil_div(M, IOU_THREADS), IOU_THREADS, 0, (cudaStream_t)stream>>>(M, corners,
                                                                                                   qcorners, iou);
        FCN_LAUNCH_CHECK();
    }
    if (stats != nullptr) {
        rbbox_iou_stats_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(M, iou, iou_thresh, stats);
        FCN_LAUNCH_CHECK();
    }
    return FCN_OK;
}
