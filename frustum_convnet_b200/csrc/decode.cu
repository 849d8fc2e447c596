// Eval-branch decode of the head logits and channel-first <-> position-major layout helpers.
//
// decode_eval_kernel replaces /root/reference/models/det_base.py:376-411 (softmax of the class,
// heading-bin and size-cluster scores, argmax, centre = offset + center_ref2) together with
// angle_decode / size_decode of models/box_transform.py:28-41,5-12.  One thread per (b, t) row.
#include "common.cuh"
#include "decode.cuh"

namespace fcn {

__global__ void decode_eval_kernel(int B, int T, int pitch, int ld, int NH, int NS,
                                   const float *__restrict__ logits,
                                   const float *__restrict__ center_ref,
                                   const float *__restrict__ mean_size, DecodeOut out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();
    pdl_launch_dependents();
    if (r >= B * T) return;
    const int b = r / T, t = r - b * T;
    decode_row(logits + ((size_t)b * pitch + t) * ld, r, b, t, T, NH, NS, center_ref, mean_size, &out, 1);
}

// (B,C,T) -> (B,T,ld): pad channels [C,ld) are written as zero.
__global__ void bct_to_btc_kernel(int C, int T, int pitch, int ld, const float *__restrict__ src,
                                  float *__restrict__ dst) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    const float *s = src + (size_t)b * C * T;
    float *d = dst + (size_t)b * pitch * ld;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, t = t0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && t < T) ? s[(size_t)c * T + t] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        if (t < T && c < ld) d[(size_t)t * ld + c] = tile[threadIdx.x][i];
    }
}

// (B,T,ld) -> (B,C,T)
__global__ void btc_to_bct_kernel(int C, int T, int pitch, int ld, const float *__restrict__ src,
                                  float *__restrict__ dst) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    const float *s = src + (size_t)b * pitch * ld;
    float *d = dst + (size_t)b * C * T;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (t < T && c < C) ? s[(size_t)t * ld + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, t = t0 + threadIdx.x;
        if (c < C && t < T) d[(size_t)c * T + t] = tile[threadIdx.x][i];
    }
}

}  // namespace fcn

using namespace fcn;

extern "C" int fcn_decode_eval(int B, int T, int pitch, int ld, int num_heading_bin, int num_size,
                               const float *logits, const float *center_ref,
                               const float *mean_size, float *cls_probs, float *center,
                               float *heading, float *size, float *heading_probs,
                               float *size_probs, fcn_stream_t stream) {
    FCN_REQUIRE(B >= 0 && T >= 0 && pitch >= T, "bad size / pitch");
    FCN_REQUIRE(num_heading_bin >= 1 && num_heading_bin <= DEC_MAX_BINS, "num_heading_bin out of range");
    FCN_REQUIRE(num_size >= 1 && num_size <= DEC_MAX_BINS, "num_size out of range");
    FCN_REQUIRE(ld >= 5 + 2 * num_heading_bin + 4 * num_size, "logit rows too short");
    if (B * T == 0) return FCN_OK;
    FCN_REQUIRE(logits && center_ref && mean_size && cls_probs && center && heading && size &&
                    heading_probs && size_probs, "NULL pointer");
    const int n = B * T;
    DecodeOut out = {cls_probs, center, heading, size, heading_probs, size_probs};
    FCN_CUDA(launch_pdl(decode_eval_kernel, dim3(ceil_div(n, 128)), dim3(128), (size_t)0, (cudaStream_t)stream,
                        B, T, pitch, ld, num_heading_bin, num_size, logits, center_ref, mean_size, out));
    return FCN_OK;
}

extern "C" int fcn_bct_to_btc(int B, int C, int T, int pitch, int ld, const float *src, float *dst,
                              fcn_stream_t stream) {
    FCN_REQUIRE(B >= 0 && C >= 1 && T >= 1 && ld >= C && pitch >= T, "bad shape");
    if (B == 0) return FCN_OK;
    FCN_REQUIRE(src && dst, "NULL pointer");
    dim3 grid(ceil_div(T, 32), ceil_div(ld, 32), B), block(32, 8);
    bct_to_btc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(C, T, pitch, ld, src, dst);
    FCN_LAUNCH_CHECK();
    return FCN_OK;
}

extern "C" int fcn_btc_to_bct(int B, int C, int T, int pitch, int ld, const float *src, float *dst,
                              fcn_stream_t stream) {
    FCN_REQUIRE(B >= 0 && C >= 1 && T >= 1 && ld >= C && pitch >= T, "bad shape");
    if (B == 0) return FCN_OK;
    FCN_REQUIRE(src && dst, "NULL pointer");
    dim3 grid(ceil_div(T, 32), ceil_div(C, 32), B), block(32, 8);
    btc_to_bct_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(C, T, pitch, ld, src, dst);
    FCN_LAUNCH_CHECK();
    return FCN_OK;
}
