// Eval-branch decode of the head logits and channel-first <-> position-major layout helpers.
//
// decode_eval_kernel replaces /root/reference/models/det_base.py:376-411 (softmax of the class,
// heading-bin and size-cluster scores, argmax, centre = offset + center_ref2) together with
// angle_decode / size_decode of models/box_transform.py:28-41,5-12.  One thread per (b, t) row.
#include "common.cuh"

namespace fcn {

constexpr int DEC_MAX_BINS = 64;

__global__ void decode_eval_kernel(int B, int T, int pitch, int ld, int NH, int NS,
                                   const float *__restrict__ logits,
                                   const float *__restrict__ center_ref,
                                   const float *__restrict__ mean_size, float *__restrict__ cls_probs,
                                   float *__restrict__ center, float *__restrict__ heading,
                                   float *__restrict__ size, float *__restrict__ heading_probs,
                                   float *__restrict__ size_probs) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();
    pdl_launch_dependents();
    if (r >= B * T) return;
    const int b = r / T, t = r - b * T;
    const float *row = logits + ((size_t)b * pitch + t) * ld;
    // class softmax (det_base.py:378)
    {
        const float a = row[0], c = row[1];
        const float m = fmaxf(a, c);
        const float ea = expf(a - m), ec = expf(c - m);
        const float s = ea + ec;
        cls_probs[(size_t)r * 2 + 0] = ea / s;
        cls_probs[(size_t)r * 2 + 1] = ec / s;
    }
    // centre = regressed offset + section centre (det_base.py:394)
#pragma unroll
    for (int c = 0; c < 3; ++c)
        center[(size_t)r * 3 + c] = __fadd_rn(row[2 + c], __ldg(center_ref + ((size_t)b * 3 + c) * T + t));
    const float *hs = row + 5, *hr = hs + NH, *ss = hr + NH, *sr = ss + NS;
    // heading: softmax, argmax (first maximum), angle_decode (box_transform.py:28-41)
    int hl = 0;
    {
        float m = hs[0];
        for (int i = 1; i < NH; ++i) m = fmaxf(m, hs[i]);
        float e[DEC_MAX_BINS], s = 0.f;
        for (int i = 0; i < NH; ++i) { e[i] = expf(hs[i] - m); s += e[i]; }
        float best = -1.f;
        for (int i = 0; i < NH; ++i) {
            const float pr = e[i] / s;
            heading_probs[(size_t)r * NH + i] = pr;
            if (pr > best) { best = pr; hl = i; }
        }
        const float apc = (float)(2.0 * 3.14159265358979323846 / (double)NH);
        const float half = (float)(2.0 * 3.14159265358979323846 / (double)NH / 2.0);
        float ang = __fadd_rn(__fmul_rn((float)hl, apc), __fmul_rn(hr[hl], half));
        if (ang > (float)3.14159265358979323846) ang = __fsub_rn(ang, (float)(2.0 * 3.14159265358979323846));
        heading[r] = ang;
    }
    // size: softmax, argmax, size_decode (box_transform.py:5-12)
    {
        float m = ss[0];
        for (int i = 1; i < NS; ++i) m = fmaxf(m, ss[i]);
        float e[DEC_MAX_BINS], s = 0.f;
        for (int i = 0; i < NS; ++i) { e[i] = expf(ss[i] - m); s += e[i]; }
        float best = -1.f;
        int sl = 0;
        for (int i = 0; i < NS; ++i) {
            const float pr = e[i] / s;
            size_probs[(size_t)r * NS + i] = pr;
            if (pr > best) { best = pr; sl = i; }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float ex = __ldg(mean_size + sl * 3 + c);
            size[(size_t)r * 3 + c] = __fadd_rn(__fmul_rn(sr[sl * 3 + c], ex), ex);
        }
    }
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
    FCN_CUDA(launch_pdl(decode_eval_kernel, dim3(ceil_div(n, 128)), dim3(128), (size_t)0, (cudaStream_t)stream,
                        B, T, pitch, ld, num_heading_bin, num_size, logits, center_ref, mean_size, cls_probs,
                        center, heading, size, heading_probs, size_probs));
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
