// Fused detection losses + their gradient w.r.t. the head logits (train half).
//
// Replaces the ~300 PyTorch elementwise / gather / reduction ops of the reference's train branch
// (/root/reference/models/det_base.py:414-503: fg selection, focal loss models/common.py:217-232, Huber losses
// models/model_util.py:9-19, box coding models/box_transform.py:5-65, corner loss models/model_util.py:48-72 +
// det_base.py:449-461, accuracies utils/utils.py:28-50, IoU metrics det_base.py:488-500) AND their autograd backward
// by three tiny launches:
//   det_loss_count_kernel : nfg = #(label == 1), nkeep = #(label != -1)
//   det_loss_rows_kernel  : one thread per (frustum, position) row: every loss term of the row, its hand-derived
//                           gradient w.r.t. the 2 + (3 + 2*NH + 4*NS) logits of the row (scaled by 1/nfg and the
//                           loss weights), accuracies, rotated IoU of the predicted box (box_iou.cuh);
//                           block reduction -> fp32 atomics into 16 accumulators
//   det_loss_final_kernel : the eight losses (det_base.py:505-514) and six metrics (:516-523)
// Means over the foreground rows are sum(w * term) / nfg, exactly the reference's `nonzero()` row selection
// (tests/test_train_host_cpu.py proves the weighted form equal to the selected form incl. gradients).
#include "box_iou.cuh"
#include "common.cuh"

namespace fcn {

constexpr int LOSS_MAX_BINS = 32;
constexpr float LOSS_PI = 3.14159265358979323846f;

enum { A_CLS = 0, A_CENTER, A_HCLS, A_HRES, A_SCLS, A_SRES, A_CORNER, A_CLSACC, A_HACC, A_SACC, A_IOU2, A_IOU3, A_IOUGT, A_N };

__global__ void det_loss_count_kernel(int N, const long long *__restrict__ lab, float *__restrict__ counts) {
    __shared__ int s[2];
    if (threadIdx.x < 2) s[threadIdx.x] = 0;
    __syncthreads();
    int fg = 0, keep = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) {
        const long long l = lab[r];
        fg += l == 1;
        keep += l != -1;
    }
    atomicAdd(&s[0], fg);
    atomicAdd(&s[1], keep);
    __syncthreads();
    if (threadIdx.x == 0) { atomicAdd(&counts[0], (float)s[0]); atomicAdd(&counts[1], (float)s[1]); }
}

__device__ __forceinline__ float huber(float a_abs, float delta) {
    const float q = fminf(a_abs, delta);
    return 0.5f * q * q + delta * (a_abs - q);
}

// corners of models/model_util.py:48-72 (x: +-l/2, y: +-h/2, z: +-w/2; rotation about y)
__device__ __forceinline__ void corners_of(const float *ctr, float heading, const float *size, float *c) {
    const float l = size[0], w = size[1], h = size[2];
    const float cs = cosf(heading), sn = sinf(heading);
    const float sx[8] = {1, 1, -1, -1, 1, 1, -1, -1}, sy[8] = {1, 1, 1, 1, -1, -1, -1, -1}, sz[8] = {1, -1, -1, 1, 1, -1, -1, 1};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float x = sx[j] * l * 0.5f, y = sy[j] * h * 0.5f, z = sz[j] * w * 0.5f;
        c[3 * j + 0] = cs * x + sn * z + ctr[0];
        c[3 * j + 1] = y + ctr[1];
        c[3 * j + 2] = -sn * x + cs * z + ctr[2];
    }
}

__global__ void __launch_bounds__(128)
det_loss_rows_kernel(const __grid_constant__ fcn_loss_args a, const float *__restrict__ counts, float *__restrict__ acc) {
    __shared__ float s_acc[A_N];
    if (threadIdx.x < A_N) s_acc[threadIdx.x] = 0.f;
    __syncthreads();
    const int N = a.B * a.T2, NH = a.NH, NS = a.NS, W = 3 + 2 * NH + 4 * NS;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    float v[A_N];
#pragma unroll
    for (int i = 0; i < A_N; ++i) v[i] = 0.f;
    if (r < N) {
        const int b = r / a.T2, t = r - b * a.T2;
        const long long lab = a.cls_label[r];
        const float nfg = fmaxf(counts[0], 1.f), inv_fg = 1.f / nfg;
        const float *cl = a.cls + (size_t)r * 2, *o = a.reg + (size_t)r * W;
        float *dc = a.dcls + (size_t)r * 2, *dr = a.dreg + (size_t)r * W;
        for (int i = 0; i < W; ++i) dr[i] = 0.f;
        // ---------------- focal classification loss (all rows with label != -1), common.py:217-232
        const float m = fmaxf(cl[0], cl[1]);
        const float e0 = expf(cl[0] - m), e1 = expf(cl[1] - m);
        const float p0 = e0 / (e0 + e1), p1 = e1 / (e0 + e1);
        float d0 = 0.f, d1 = 0.f;
        if (lab != -1) {
            const int tl = lab >= 1 ? 1 : 0;
            const float alpha = tl == 0 ? 0.75f : 0.25f;
            const float pt = tl == 0 ? p0 : p1;
            const float lg = logf(pt + 1e-14f);
            v[A_CLS] = -alpha * (1.f - pt) * (1.f - pt) * lg;
            // d/dpt [ -alpha (1-pt)^2 log(pt+eps) ]
            const float dpt = -alpha * (-2.f * (1.f - pt) * lg + (1.f - pt) * (1.f - pt) / (pt + 1e-14f));
            const float scale = dpt / (counts[0] + 1e-14f);           // loss.sum() / (num_fg + 1e-14)
            // dpt/dc_k = pt (delta_kt - p_k)
            d0 = scale * pt * ((tl == 0 ? 1.f : 0.f) - p0);
            d1 = scale * pt * ((tl == 1 ? 1.f : 0.f) - p1);
            v[A_CLSACC] = ((p1 > p0 ? 1 : 0) == (int)lab) ? 1.f : 0.f;   // argmax: first maximum on ties
        }
        dc[0] = d0; dc[1] = d1;
        if (lab == 1) {
            // ---------------- labels of this frustum, box_transform.py:15-25
            const float cgt[3] = {a.box3d_center[b * 3], a.box3d_center[b * 3 + 1], a.box3d_center[b * 3 + 2]};
            const float ref[3] = {a.center_ref2[((size_t)b * 3 + 0) * a.T2 + t], a.center_ref2[((size_t)b * 3 + 1) * a.T2 + t],
                                  a.center_ref2[((size_t)b * 3 + 2) * a.T2 + t]};
            const float hgt = a.box3d_heading[b];
            const float sgt[3] = {a.box3d_size[b * 3], a.box3d_size[b * 3 + 1], a.box3d_size[b * 3 + 2]};
            const int sc = (int)a.size_class[b];
            const float two_pi = 2.f * LOSS_PI, apc = two_pi / (float)NH;
            float g = fmodf(hgt, two_pi);
            if (g < 0.f) g += two_pi;                                  // python-style modulo
            float shifted = fmodf(g + apc * 0.5f, two_pi);
            if (shifted < 0.f) shifted += two_pi;
            const int hc = min(max((int)floorf(shifted / apc), 0), NH - 1);
            const float hres_lab = (shifted - ((float)hc * apc + apc * 0.5f)) / (apc * 0.5f);
            const float *ctr = o, *hs = o + 3, *hr = hs + NH, *ss = hr + NH, *sr = ss + NS;
            float ex[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) ex[i] = a.mean_size[sc * 3 + i];
            const float wb = a.w_box * inv_fg;                         // weight of a box term of ONE fg row in the total
            // ---------------- centre Huber loss (delta 3) on the distance
            {
                float dv[3], d2 = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) { dv[i] = (cgt[i] - ref[i]) - ctr[i]; d2 += dv[i] * dv[i]; }
                const float d = sqrtf(d2);
                v[A_CENTER] = huber(d, 3.f);
                const float gd = d > 0.f ? fminf(d, 3.f) / d : 0.f;   // dh/dd * 1/d
#pragma unroll
                for (int i = 0; i < 3; ++i) dr[i] += wb * gd * (-dv[i]);
            }
            // ---------------- heading class CE + residual Huber (delta 1)
            int h_arg = 0;
            {
                float mx = hs[0];
                for (int i = 1; i < NH; ++i) mx = fmaxf(mx, hs[i]);
                float sum = 0.f;
                for (int i = 0; i < NH; ++i) sum += expf(hs[i] - mx);
                v[A_HCLS] = -(hs[hc] - mx - logf(sum));
                float best = -1.f;
                for (int i = 0; i < NH; ++i) {
                    const float p = expf(hs[i] - mx) / sum;
                    dr[3 + i] += wb * (p - (i == hc ? 1.f : 0.f));
                    if (p > best) { best = p; h_arg = i; }
                }
                v[A_HACC] = h_arg == hc ? 1.f : 0.f;
                const float e = hr[hc] - hres_lab;
                v[A_HRES] = huber(fabsf(e), 1.f);
                dr[3 + NH + hc] += wb * a.w_head_reg * fmaxf(-1.f, fminf(1.f, e));
            }
            // ---------------- size class CE + residual Huber (delta 1) on the norm
            int s_arg = 0;
            {
                float mx = ss[0];
                for (int i = 1; i < NS; ++i) mx = fmaxf(mx, ss[i]);
                float sum = 0.f;
                for (int i = 0; i < NS; ++i) sum += expf(ss[i] - mx);
                v[A_SCLS] = -(ss[sc] - mx - logf(sum));
                float best = -1.f;
                for (int i = 0; i < NS; ++i) {
                    const float p = expf(ss[i] - mx) / sum;
                    dr[3 + 2 * NH + i] += wb * (p - (i == sc ? 1.f : 0.f));
                    if (p > best) { best = p; s_arg = i; }
                }
                v[A_SACC] = s_arg == sc ? 1.f : 0.f;
                float dv[3], d2 = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) { dv[i] = (sgt[i] - ex[i]) / ex[i] - sr[sc * 3 + i]; d2 += dv[i] * dv[i]; }
                const float d = sqrtf(d2);
                v[A_SRES] = huber(d, 1.f);
                const float gd = d > 0.f ? fminf(d, 1.f) / d : 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) dr[3 + 2 * NH + NS + sc * 3 + i] += wb * a.w_size_reg * gd * (-dv[i]);
            }
            // ---------------- corner loss (det_base.py:449-461): predicted box with the GT class labels
            {
                const float cp[3] = {ref[0] + ctr[0], ref[1] + ctr[1], ref[2] + ctr[2]};
                float ang = (float)hc * apc + hr[hc] * (apc * 0.5f);
                if (ang > LOSS_PI) ang -= two_pi;
                float sz[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) sz[i] = sr[sc * 3 + i] * ex[i] + ex[i];
                float P[24], G[24], F[24];
                corners_of(cp, ang, sz, P);
                corners_of(cgt, hgt, sgt, G);
                corners_of(cgt, hgt + LOSS_PI, sgt, F);
                float dg = 0.f, df = 0.f, ng[8], nf[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float a2 = 0.f, b2 = 0.f;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float x = P[3 * j + k] - G[3 * j + k], y = P[3 * j + k] - F[3 * j + k];
                        a2 += x * x; b2 += y * y;
                    }
                    ng[j] = sqrtf(a2); nf[j] = sqrtf(b2);
                    dg += ng[j]; df += nf[j];
                }
                dg *= 0.125f; df *= 0.125f;
                const bool use_g = dg <= df;
                const float D = use_g ? dg : df;
                v[A_CORNER] = huber(D, 1.f);
                const float gD = wb * a.w_corner * fminf(D, 1.f) * 0.125f;   // dL/dD * dD/d||.||_j
                const float cs = cosf(ang), sn = sinf(ang);
                const float sx[8] = {1, 1, -1, -1, 1, 1, -1, -1}, sy[8] = {1, 1, 1, 1, -1, -1, -1, -1},
                            szz[8] = {1, -1, -1, 1, 1, -1, -1, 1};
                float gc[3] = {0.f, 0.f, 0.f}, gth = 0.f, gl = 0.f, gw = 0.f, gh = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float nrm = use_g ? ng[j] : nf[j];
                    if (!(nrm > 0.f)) continue;
                    const float *Tg = use_g ? G : F;
                    const float gx = gD * (P[3 * j] - Tg[3 * j]) / nrm, gy = gD * (P[3 * j + 1] - Tg[3 * j + 1]) / nrm,
                                gz = gD * (P[3 * j + 2] - Tg[3 * j + 2]) / nrm;
                    gc[0] += gx; gc[1] += gy; gc[2] += gz;
                    const float x = sx[j] * sz[0] * 0.5f, z = szz[j] * sz[1] * 0.5f;
                    gth += gx * (-sn * x + cs * z) + gz * (-cs * x - sn * z);
                    gl += (gx * cs - gz * sn) * sx[j] * 0.5f;
                    gw += (gx * sn + gz * cs) * szz[j] * 0.5f;
                    gh += gy * sy[j] * 0.5f;
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) dr[i] += gc[i];
                dr[3 + NH + hc] += gth * (apc * 0.5f);
                dr[3 + 2 * NH + NS + sc * 3 + 0] += gl * ex[0];
                dr[3 + 2 * NH + NS + sc * 3 + 1] += gw * ex[1];
                dr[3 + 2 * NH + NS + sc * 3 + 2] += gh * ex[2];
                // ---------------- IoU metrics: predicted box with the PREDICTED class labels (det_base.py:488-500)
                if (a.with_iou) {
                    float angp = (float)h_arg * apc + hr[h_arg] * (apc * 0.5f);
                    if (angp > LOSS_PI) angp -= two_pi;
                    float szp[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) szp[i] = sr[s_arg * 3 + i] * a.mean_size[s_arg * 3 + i] + a.mean_size[s_arg * 3 + i];
                    corners_of(cp, angp, szp, P);
                    float iou[2];
                    rbbox_iou_pair(P, G, iou);
                    v[A_IOU2] = iou[0];
                    v[A_IOU3] = iou[1];
                    v[A_IOUGT] = iou[1] >= a.iou_thresh ? 1.f : 0.f;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < A_N; ++i) {
        float x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x != 0.f) atomicAdd(&s_acc[i], x);
    }
    __syncthreads();
    if (threadIdx.x < A_N && s_acc[threadIdx.x] != 0.f) atomicAdd(&acc[threadIdx.x], s_acc[threadIdx.x]);
}

// out[0..7]  = total, cls, center, head_cls, head_res, size_cls, size_res, corners   (det_base.py:505-514)
// out[8..13] = cls_acc, head_acc, size_acc, IoU_2D, IoU_3D, IoU_>=thresh            (:516-523)
__global__ void det_loss_final_kernel(const __grid_constant__ fcn_loss_args a, const float *__restrict__ counts,
                                      const float *__restrict__ acc, float *__restrict__ out) {
    if (threadIdx.x != 0) return;
    const float nfg = fmaxf(counts[0], 1.f), nkeep = fmaxf(counts[1], 1.f);
    const float cls = acc[A_CLS] / (counts[0] + 1e-14f);
    const float center = acc[A_CENTER] / nfg, hcls = acc[A_HCLS] / nfg, hres = acc[A_HRES] / nfg;
    const float scls = acc[A_SCLS] / nfg, sres = acc[A_SRES] / nfg, corner = acc[A_CORNER] / nfg;
    out[0] = cls + a.w_box * (center + hcls + scls + a.w_head_reg * hres + a.w_size_reg * sres + a.w_corner * corner);
    out[1] = cls; out[2] = center; out[3] = hcls; out[4] = hres; out[5] = scls; out[6] = sres; out[7] = corner;
    out[8] = acc[A_CLSACC] / nkeep;
    out[9] = acc[A_HACC] / nfg;
    out[10] = acc[A_SACC] / nfg;
    const float nan = __int_as_float(0x7fc00000);
    out[11] = a.with_iou ? acc[A_IOU2] / nfg : nan;
    out[12] = a.with_iou ? acc[A_IOU3] / nfg : nan;
    out[13] = a.with_iou ? acc[A_IOUGT] / nfg : nan;
    out[14] = counts[0];
    out[15] = counts[1];
}

}  // namespace fcn

using namespace fcn;

extern "C" int fcn_det_loss(const fcn_loss_args *args, fcn_stream_t stream) {
    FCN_REQUIRE(args != nullptr, "args is NULL");
    const fcn_loss_args &a = *args;
    FCN_REQUIRE(a.B >= 1 && a.T2 >= 1 && a.NH >= 1 && a.NH <= LOSS_MAX_BINS && a.NS >= 1 && a.NS <= LOSS_MAX_BINS, "bad sizes");
    FCN_REQUIRE(a.cls && a.reg && a.center_ref2 && a.cls_label && a.box3d_center && a.box3d_heading && a.box3d_size &&
                    a.size_class && a.mean_size && a.dcls && a.dreg && a.out && a.scratch, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int N = a.B * a.T2;
    float *counts = a.scratch, *acc = a.scratch + 2;
    FCN_CUDA(cudaMemsetAsync(a.scratch, 0, sizeof(float) * (2 + A_N), st));
    det_loss_count_kernel<<<min(ceil_div(N, 256), 64), 256, 0, st>>>(N, (const long long *)a.cls_label, counts);
    FCN_LAUNCH_CHECK();
    det_loss_rows_kernel<<<ceil_div(N, 128), 128, 0, st>>>(a, counts, acc);
    FCN_LAUNCH_CHECK();
    det_loss_final_kernel<<<1, 32, 0, st>>>(a, counts, acc, a.out);
    FCN_LAUNCH_CHECK();
    return FCN_OK;
}
