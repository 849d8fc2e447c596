// Eval-branch decode of ONE head-logit row, shared by decode_eval_kernel (stand-alone launch) and the heads
// epilogue of the persistent FCN kernel (fcn_mega.cu), so both produce bit-identical results.
// Replaces /root/reference/models/det_base.py:376-411 (softmax of the class, heading-bin and size-cluster
// scores, argmax, centre = offset + center_ref2) + angle_decode / size_decode of
// models/box_transform.py:28-41,5-12.  Op order follows the reference (__fmul_rn/__fadd_rn: no FMA contraction).
#pragma once
#include <cuda_runtime.h>

namespace fcn {

constexpr int DEC_MAX_BINS = 64;
constexpr int DEC_MAX_PEERS = 8;

// the six outputs of det_base.py:411 as flat row-major blocks
struct DecodeOut {
    float *cls_probs, *center, *heading, *size, *heading_probs, *size_probs;
};

// row: [cls0, cls1, center(3), heading scores(NH), heading res(NH), size scores(NS), size res(NS*3)];
// r = b*T + t is the output row; the results are stored to `n_out` output sets (the local block and, in the
// multi-GPU path, the peers' gather buffers over NVLink).
__device__ __forceinline__ void decode_row(const float *row, int r, int b, int t, int T, int NH, int NS,
                                           const float *__restrict__ center_ref,
                                           const float *__restrict__ mean_size, const DecodeOut *outs, int n_out) {
    // class softmax (det_base.py:378)
    {
        const float a = row[0], c = row[1];
        const float m = fmaxf(a, c);
        const float ea = expf(a - m), ec = expf(c - m);
        const float s = ea + ec;
        const float p0 = ea / s, p1 = ec / s;
        for (int o = 0; o < n_out; ++o) {
            outs[o].cls_probs[(size_t)r * 2 + 0] = p0;
            outs[o].cls_probs[(size_t)r * 2 + 1] = p1;
        }
    }
    // centre = regressed offset + section centre (det_base.py:394)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = __fadd_rn(row[2 + c], __ldg(center_ref + ((size_t)b * 3 + c) * T + t));
        for (int o = 0; o < n_out; ++o) outs[o].center[(size_t)r * 3 + c] = v;
    }
    const float *hs = row + 5, *hr = hs + NH, *ss = hr + NH, *sr = ss + NS;
    // heading: softmax, argmax (first maximum), angle_decode (box_transform.py:28-41)
    {
        int hl = 0;
        float m = hs[0];
        for (int i = 1; i < NH; ++i) m = fmaxf(m, hs[i]);
        float s = 0.f;
        for (int i = 0; i < NH; ++i) s += expf(hs[i] - m);
        float best = -1.f;
        for (int i = 0; i < NH; ++i) {
            const float pr = expf(hs[i] - m) / s;
            for (int o = 0; o < n_out; ++o) outs[o].heading_probs[(size_t)r * NH + i] = pr;
            if (pr > best) { best = pr; hl = i; }
        }
        const float apc = (float)(2.0 * 3.14159265358979323846 / (double)NH);
        const float half = (float)(2.0 * 3.14159265358979323846 / (double)NH / 2.0);
        float ang = __fadd_rn(__fmul_rn((float)hl, apc), __fmul_rn(hr[hl], half));
        if (ang > (float)3.14159265358979323846) ang = __fsub_rn(ang, (float)(2.0 * 3.14159265358979323846));
        for (int o = 0; o < n_out; ++o) outs[o].heading[r] = ang;
    }
    // size: softmax, argmax, size_decode (box_transform.py:5-12)
    {
        float m = ss[0];
        for (int i = 1; i < NS; ++i) m = fmaxf(m, ss[i]);
        float s = 0.f;
        for (int i = 0; i < NS; ++i) s += expf(ss[i] - m);
        float best = -1.f;
        int sl = 0;
        for (int i = 0; i < NS; ++i) {
            const float pr = expf(ss[i] - m) / s;
            for (int o = 0; o < n_out; ++o) outs[o].size_probs[(size_t)r * NS + i] = pr;
            if (pr > best) { best = pr; sl = i; }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float ex = __ldg(mean_size + sl * 3 + c);
            const float v = __fadd_rn(__fmul_rn(sr[sl * 3 + c], ex), ex);
            for (int o = 0; o < n_out; ++o) outs[o].size[(size_t)r * 3 + c] = v;
        }
    }
}

}  // namespace fcn
