// Training-mode kernels of the frustum hot path (config 5: cfgs/refine_car.yaml, fwd + bwd + optimizer).
//
// Replaces, for PointNetDet in train() mode, what the reference runs through PyTorch/cuDNN autograd:
//   conv (1x1 / k3 / strided / transposed) + BatchNorm with BATCH statistics + ReLU      models/common.py:38-63
//   PointNetModule / PointNetFeat (mask, max over K)                                      models/det_base.py:62-159
//   ConvFeatNet (concats, crops) and the two heads                                        models/det_base.py:163-224,367-368
//   the optimizer step                                                                    train/train_net_det.py:121-128,322-323
// Batch-statistics BN breaks the eval fusions (the statistics of layer l must be complete before layer l+1 can
// normalise its input), so the train step is a sequence of table-driven kernels over dense position-major fp32
// tensors [rows, channels]:
//   * no post-activation tensor is ever materialised: every consumer applies "scale*y + shift, ReLU" of its
//     producer (scale/shift derived on the fly from the producer's Sum(y), Sum(y^2)) while loading its operand;
//     kernel taps / strides / concats / the pixel shuffle of transposed convs / the crop are index maps of the
//     operand load (fcn_train_src / fcn_train_seg);
//   * forward GEMM  : Y = act(X) * W (+bias), epilogue accumulates Sum(y), Sum(y^2) per channel (fp64 atomics);
//   * backward       : with dz = dA * 1[pre > 0], xh = (y - mean) * invstd, the BN backward
//                       dY = scale * (dz - mean(dz) - xh * mean(dz * xh))
//                      is also applied on the fly: a column reduction produces Sum(dz), Sum(dz*xh) (= dbeta,
//                      dgamma), then  dW += act(X)^T * dY  and  dX += dY * W^T  read Y, dA and the sums directly;
//   * weights and gradients are addressed IN THE PARAMETER LAYOUT (Conv1d (Co,Ci,k), ConvTranspose1d (Ci,Co,k),
//     Conv2d (Co,Ci,1,1)) through strides, so gradients land directly in one flat bucket (one all-reduce);
//   * fp32 FMA arithmetic (parity with the fp32 reference fixture: losses 2e-4, gradients 2e-3).
#include "common.cuh"

namespace fcn {

constexpr float TR_EPS = 1e-5f;
constexpr int TG_BM = 64, TG_BN = 64, TG_BK = 16, TG_THREADS = 256;

struct BnCoef {
    float scale, shift, mean, invstd;
};

// scale/shift/mean/invstd of channel c of a BN layer from its Sum(y), Sum(y^2) (biased variance: training forward)
__device__ __forceinline__ BnCoef bn_coef(const double *sums, const float *gamma, const float *beta, int C, int c,
                                          double inv_count) {
    const double mean = sums[c] * inv_count;
    double var = sums[C + c] * inv_count - mean * mean;
    var = var > 0.0 ? var : 0.0;
    BnCoef r;
    r.mean = (float)mean;
    r.invstd = (float)(1.0 / sqrt(var + (double)TR_EPS));
    r.scale = gamma[c] * r.invstd;
    r.shift = beta[c] - r.mean * r.scale;
    return r;
}

// operand element: post-activation value of source `s` at (frustum b, position p, channel c); 0 outside [0, T*up)
__device__ __forceinline__ float src_act(const fcn_train_src &s, int b, int p, int c, float scale, float shift) {
    if (p < 0 || p >= s.T * s.up) return 0.f;
    const int row = b * s.T + p / s.up, col = s.c0 + (p % s.up) * s.cup + c;
    float x = s.raw[(size_t)row * s.ld + col];
    if (s.sums != nullptr) {
        x = fmaf(x, scale, shift);
        if (s.relu) x = fmaxf(x, 0.f);
    }
    return x;
}

// gradient element dY(r, n) of a layer (BN backward applied on the fly; plain layers: dY = dA)
struct DyCtx {
    float scale, m1, m2, mean, invstd, shift;
};
__device__ __forceinline__ DyCtx dy_ctx(const fcn_train_layer &L, int n) {
    DyCtx c;
    if (!L.has_bn) { c.scale = 1.f; c.m1 = c.m2 = c.mean = c.shift = 0.f; c.invstd = 1.f; return c; }
    const int co = n % L.Cout;
    const double inv = 1.0 / ((double)L.B * L.T_out * L.up);
    const BnCoef b = bn_coef(L.sums, L.gamma, L.beta, L.Cout, co, inv);
    c.scale = b.scale; c.shift = b.shift; c.mean = b.mean; c.invstd = b.invstd;
    c.m1 = (float)(L.dsums[co] * inv);
    c.m2 = (float)(L.dsums[L.Cout + co] * inv);
    return c;
}
__device__ __forceinline__ float dy_elem(const fcn_train_layer &L, const DyCtx &c, int r, int n) {
    const float da = L.dA[(size_t)r * L.N + n];
    if (!L.has_bn) return da;
    const float y = L.Y[(size_t)r * L.N + n];
    const float dz = (L.relu && fmaf(y, c.scale, c.shift) <= 0.f) ? 0.f : da;
    const float xh = (y - c.mean) * c.invstd;
    return c.scale * (dz - c.m1 - xh * c.m2);
}

// weight element (segment channel c, output column n) in the parameter layout
__device__ __forceinline__ size_t w_index(const fcn_train_layer &L, const fcn_train_seg &g, int c, int n) {
    const int j = n / L.Cout, co = n - j * L.Cout;
    return (size_t)g.w_off + (size_t)c * g.s_ci + (size_t)co * L.s_co + (size_t)j * L.s_j;
}

// ------------------------------------------------------------------ forward GEMM + statistics
// The FCN layers of the refinement stage are SKINNY (M = B*T = 96..640 rows, K up to 1536): a (M/64, N/64) grid
// is 4-32 CTAs, each with a long serial K loop (measured 100-270 us per layer).  The K tiles of all segments
// are therefore linearised and split over gridDim.z; slices write partial sums to a workspace and
// train_fwd_finish_kernel adds them in a fixed order (deterministic), adds the bias, stores Y and accumulates
// the batch statistics.  gridDim.z == 1 stores Y / statistics directly.
__global__ void __launch_bounds__(TG_THREADS)
train_fwd_kernel(const __grid_constant__ fcn_train_layer L, float *__restrict__ partial, int tiles_per_slice) {
    __shared__ float As[TG_BK][TG_BM + 4], Bs[TG_BK][TG_BN + 4];
    __shared__ float s_scale[TG_BK], s_shift[TG_BK];
    __shared__ float s_sum[TG_BN], s_sq[TG_BN];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int r0 = blockIdx.x * TG_BM, n0 = blockIdx.y * TG_BN;
    const int M = L.B * L.T_out;
    const int q_lo = blockIdx.z * tiles_per_slice, q_hi = q_lo + tiles_per_slice;
    float acc[4][4] = {};
    int q0 = 0;                                              // first linear K tile of the current segment
    for (int sg = 0; sg < L.n_seg; ++sg) {
        const fcn_train_seg &g = L.seg[sg];
        const fcn_train_src &s = g.src;
        const double inv = s.sums != nullptr ? 1.0 / s.count : 0.0;
        const int nt = (g.C + TG_BK - 1) / TG_BK;
        const int t_lo = max(q_lo - q0, 0), t_hi = min(q_hi - q0, nt);
        q0 += nt;
        for (int t = t_lo; t < t_hi; ++t) {
            const int k0 = t * TG_BK;
            __syncthreads();
            if (tid < TG_BK) {
                const int c = k0 + tid;
                float sc = 1.f, sh = 0.f;
                if (s.sums != nullptr && c < g.C) {
                    const BnCoef b = bn_coef(s.sums, s.gamma, s.beta, s.Cstat, (s.coff + c) % s.Cstat, inv);
                    sc = b.scale; sh = b.shift;
                }
                s_scale[tid] = sc; s_shift[tid] = sh;
            }
            __syncthreads();
            // A tile: 64 rows x 16 channels (consecutive threads -> consecutive channels)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rr = ty + 16 * i, r = r0 + rr, c = k0 + tx;
                float v = 0.f;
                if (r < M && c < g.C) {
                    const int b = r / L.T_out, tt = r - b * L.T_out;
                    v = src_act(s, b, tt * g.stride + g.tap, c, s_scale[tx], s_shift[tx]);
                }
                As[tx][rr] = v;
            }
            // B tile: 16 channels x 64 columns
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int kk = tid / 64 + 4 * i, nn = tid % 64, c = k0 + kk, n = n0 + nn;
                Bs[kk][nn] = (c < g.C && n < L.N) ? L.W[w_index(L, g, c, n)] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < TG_BK; ++kk) {
                float a[4], bq[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) bq[j] = Bs[kk][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bq[j], acc[i][j]);
            }
        }
    }
    if (gridDim.z > 1) {                                     // partial sums of this K slice
        float *dst = partial + (size_t)blockIdx.z * M * L.N;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + ty * 4 + i;
            if (r >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tx * 4 + j;
                if (n < L.N) dst[(size_t)r * L.N + n] = acc[i][j];
            }
        }
        return;
    }
    // epilogue: store raw output (+bias), per-channel Sum(y), Sum(y^2)
    if (tid < TG_BN) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        float ps = 0.f, pq = 0.f;
        if (n < L.N) {
            const float bias = L.bias != nullptr ? L.bias[n] : 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r0 + ty * 4 + i;
                if (r < M) {
                    const float y = acc[i][j] + bias;
                    L.Y[(size_t)r * L.N + n] = y;
                    ps += y; pq += y * y;
                }
            }
            if (L.has_bn) { atomicAdd(&s_sum[tx * 4 + j], ps); atomicAdd(&s_sq[tx * 4 + j], pq); }
        }
    }
    if (L.has_bn) {
        __syncthreads();
        if (tid < TG_BN && n0 + tid < L.N) {
            const int co = (n0 + tid) % L.Cout;
            atomicAdd(&L.sums[co], (double)s_sum[tid]);
            atomicAdd(&L.sums[L.Cout + co], (double)s_sq[tid]);
        }
    }
}

// Y = sum over K slices (fixed order) + bias; batch statistics.  Block = 32 columns x 8 row lanes.
__global__ void __launch_bounds__(256)
train_fwd_finish_kernel(const __grid_constant__ fcn_train_layer L, const float *__restrict__ partial, int nslice,
                        int rows_per_block) {
    const int nl = threadIdx.x % 32, rl = threadIdx.x / 32;
    const int n = blockIdx.x * 32 + nl;
    const int M = L.B * L.T_out;
    const int rb = blockIdx.y * rows_per_block, re = min(M, rb + rows_per_block);
    __shared__ float s1[8][33], s2[8][33];
    float a1 = 0.f, a2 = 0.f;
    if (n < L.N) {
        const float bias = L.bias != nullptr ? L.bias[n] : 0.f;
        for (int r = rb + rl; r < re; r += 8) {
            float y = bias;
            for (int z = 0; z < nslice; ++z) y += partial[((size_t)z * M + r) * L.N + n];
            L.Y[(size_t)r * L.N + n] = y;
            a1 += y; a2 += y * y;
        }
    }
    if (!L.has_bn) return;
    s1[rl][nl] = a1; s2[rl][nl] = a2;
    __syncthreads();
    if (rl == 0 && n < L.N) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { t1 += s1[i][nl]; t2 += s2[i][nl]; }
        const int co = n % L.Cout;
        atomicAdd(&L.sums[co], (double)t1);
        atomicAdd(&L.sums[L.Cout + co], (double)t2);
    }
}

// ------------------------------------------------------------------ backward: column reduction Sum(dz), Sum(dz*xh)
__global__ void __launch_bounds__(256)
train_reduce_kernel(const __grid_constant__ fcn_train_layer L, int rows_per_block) {
    // block = 32 columns x 8 row lanes
    const int nl = threadIdx.x % 32, rl = threadIdx.x / 32;
    const int n = blockIdx.x * 32 + nl;
    const int M = L.B * L.T_out;
    const int rb = blockIdx.y * rows_per_block, re = min(M, rb + rows_per_block);
    __shared__ float s1[8][33], s2[8][33];
    float a1 = 0.f, a2 = 0.f;
    if (n < L.N) {
        BnCoef b = {1.f, 0.f, 0.f, 1.f};
        if (L.has_bn) b = bn_coef(L.sums, L.gamma, L.beta, L.Cout, n % L.Cout, 1.0 / ((double)M * L.up));
        for (int r = rb + rl; r < re; r += 8) {
            const float da = L.dA[(size_t)r * L.N + n];
            if (!L.has_bn) { a1 += da; continue; }
            const float y = L.Y[(size_t)r * L.N + n];
            const float dz = (L.relu && fmaf(y, b.scale, b.shift) <= 0.f) ? 0.f : da;
            a1 += dz;
            a2 += dz * ((y - b.mean) * b.invstd);
        }
    }
    s1[rl][nl] = a1; s2[rl][nl] = a2;
    __syncthreads();
    if (rl == 0 && n < L.N) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { t1 += s1[i][nl]; t2 += s2[i][nl]; }
        const int co = L.has_bn ? n % L.Cout : n;
        const int C = L.has_bn ? L.Cout : L.N;
        atomicAdd(&L.dsums[co], (double)t1);
        if (L.has_bn) atomicAdd(&L.dsums[C + co], (double)t2);
    }
}

// ------------------------------------------------------------------ backward: dW += act(X)^T * dY
__global__ void __launch_bounds__(TG_THREADS)
train_dw_kernel(const __grid_constant__ fcn_train_layer L, int seg_idx, int rows_per_block) {
    __shared__ float As[TG_BK][TG_BM + 4], Bs[TG_BK][TG_BN + 4];   // As[row][k-channel], Bs[row][n]
    __shared__ float s_scale[TG_BM], s_shift[TG_BM];
    __shared__ DyCtx s_dy[TG_BN];
    const fcn_train_seg &g = L.seg[seg_idx];
    const fcn_train_src &s = g.src;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int c0 = blockIdx.x * TG_BM, n0 = blockIdx.y * TG_BN;
    const int M = L.B * L.T_out;
    const int rb = blockIdx.z * rows_per_block, re = min(M, rb + rows_per_block);
    if (tid < TG_BM) {
        const int c = c0 + tid;
        float sc = 1.f, sh = 0.f;
        if (s.sums != nullptr && c < g.C) {
            const BnCoef b = bn_coef(s.sums, s.gamma, s.beta, s.Cstat, (s.coff + c) % s.Cstat, 1.0 / s.count);
            sc = b.scale; sh = b.shift;
        }
        s_scale[tid] = sc; s_shift[tid] = sh;
    }
    if (tid < TG_BN) s_dy[tid] = dy_ctx(L, min(n0 + tid, L.N - 1));
    float acc[4][4] = {};
    for (int rr0 = rb; rr0 < re; rr0 += TG_BK) {
        __syncthreads();
        // A tile: 16 rows x 64 channels of this segment
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = tid / 64 + 4 * i, cc = tid % 64, r = rr0 + rr, c = c0 + cc;
            float v = 0.f;
            if (r < re && c < g.C) {
                const int b = r / L.T_out, t = r - b * L.T_out;
                v = src_act(s, b, t * g.stride + g.tap, c, s_scale[cc], s_shift[cc]);
            }
            As[rr][cc] = v;
        }
        // dY tile: 16 rows x 64 columns
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = tid / 64 + 4 * i, nn = tid % 64, r = rr0 + rr, n = n0 + nn;
            Bs[rr][nn] = (r < re && n < L.N) ? dy_elem(L, s_dy[nn], r, n) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TG_BK; ++kk) {
            float a[4], bq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bq[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bq[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty * 4 + i;
        if (c >= g.C) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < L.N) atomicAdd(&L.dW[w_index(L, g, c, n)], acc[i][j]);
        }
    }
}

// ------------------------------------------------------------------ backward: dX(seg) += dY * W^T
__global__ void __launch_bounds__(TG_THREADS)
train_dx_kernel(const __grid_constant__ fcn_train_layer L, int seg_idx, int n_per_slice) {
    __shared__ float As[TG_BK][TG_BM + 4], Bs[TG_BK][TG_BN + 4];   // As[n][row], Bs[n][channel]
    __shared__ DyCtx s_dy[TG_BK];
    const fcn_train_seg &g = L.seg[seg_idx];
    const fcn_train_src &s = g.src;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int r0 = blockIdx.x * TG_BM, c0 = blockIdx.y * TG_BN;
    const int M = L.B * L.T_out;
    float acc[4][4] = {};
    const int n_lo = blockIdx.z * n_per_slice, n_hi = min(L.N, n_lo + n_per_slice);
    for (int nn0 = n_lo; nn0 < n_hi; nn0 += TG_BK) {
        __syncthreads();
        if (tid < TG_BK) s_dy[tid] = dy_ctx(L, min(nn0 + tid, L.N - 1));
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {       // dY tile: 64 rows x 16 columns (consecutive threads -> consecutive n)
            const int rr = ty + 16 * i, r = r0 + rr, n = nn0 + tx;
            As[tx][rr] = (r < M && n < n_hi) ? dy_elem(L, s_dy[tx], r, n) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {       // W tile: 16 columns x 64 channels
            const int kk = tid / 64 + 4 * i, cc = tid % 64, n = nn0 + kk, c = c0 + cc;
            Bs[kk][cc] = (n < n_hi && c < g.C) ? L.W[w_index(L, g, c, n)] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TG_BK; ++kk) {
            float a[4], bq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bq[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bq[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty * 4 + i;
        if (r >= M) continue;
        const int b = r / L.T_out, t = r - b * L.T_out, p = t * g.stride + g.tap;
        if (p < 0 || p >= s.T * s.up) continue;
        const int row = b * s.T + p / s.up, colb = s.c0 + (p % s.up) * s.cup;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + tx * 4 + j;
            if (c < g.C) atomicAdd(&s.grad[(size_t)row * s.ld + colb + c], acc[i][j]);
        }
    }
}

// ------------------------------------------------------------------ PointNet pooling (mask, max over K) fwd / bwd
// feat[(b,t), c] = cnt>0 ? max_k relu(scale*y + shift) : 0 ; argmax kept for the backward; one-hot columns appended
__global__ void train_pool_fwd_kernel(fcn_train_pool_args P) {
    const int bt = blockIdx.x, c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= P.C + P.V) return;
    float *out = P.feat + (size_t)bt * P.ld_feat;
    if (c >= P.C) { out[c] = P.one_hot[(size_t)(bt / P.T) * P.V + (c - P.C)]; return; }
    const BnCoef b = bn_coef(P.sums, P.gamma, P.beta, P.C, c, 1.0 / ((double)P.B * P.T * P.K));
    float best = -INFINITY;
    int arg = 0;
    const float *y = P.Y + (size_t)bt * P.K * P.C + c;
    const bool valid = P.cnt[bt] > 0;
    for (int k = 0; k < P.K; ++k) {
        float a = fmaxf(fmaf(y[(size_t)k * P.C], b.scale, b.shift), 0.f);
        if (!valid) a = 0.f;                 // x * (cnt > 0)  (det_base.py:100-101): all-zero rows, argmax = 0
        if (a > best) { best = a; arg = k; }
    }
    out[c] = best;
    P.argmax[(size_t)bt * P.C + c] = arg;
}
// dA3[(b,t,argmax), c] = cnt>0 ? dfeat[(b,t), c] : 0   (dA3 is zero-filled by the caller)
__global__ void train_pool_bwd_kernel(fcn_train_pool_args P) {
    const int bt = blockIdx.x, c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= P.C || P.cnt[bt] <= 0) return;
    const int arg = P.argmax[(size_t)bt * P.C + c];
    P.dA[((size_t)bt * P.K + arg) * P.C + c] = P.dfeat[(size_t)bt * P.ld_feat + c];
}

// ------------------------------------------------------------------ end of step: BN bookkeeping for ALL layers
// running statistics (momentum 0.1, unbiased variance; models/common.py BatchNorm defaults), dgamma / dbeta / dbias
__global__ void train_finalize_kernel(const fcn_train_layer *layers, int n_layers, int update_running) {
    const fcn_train_layer &L = layers[blockIdx.x];
    const double cnt = (double)L.B * L.T_out * L.up;
    for (int c = threadIdx.x; c < (L.has_bn ? L.Cout : L.N); c += blockDim.x) {
        if (!L.has_bn) {
            if (L.dbias != nullptr) L.dbias[c] += (float)L.dsums[c];
            continue;
        }
        L.dbeta[c] += (float)L.dsums[c];
        L.dgamma[c] += (float)L.dsums[L.Cout + c];
        if (update_running) {
            const double mean = L.sums[c] / cnt;
            double var = L.sums[L.Cout + c] / cnt - mean * mean;
            var = var > 0.0 ? var : 0.0;
            const double unb = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
            L.run_mean[c] = (float)(0.9 * (double)L.run_mean[c] + 0.1 * mean);
            L.run_var[c] = (float)(0.9 * (double)L.run_var[c] + 0.1 * unb);
        }
    }
}

// ------------------------------------------------------------------ fused Adam over the flat parameter bucket
// torch.optim.Adam semantics (train_net_det.py:322-323): g += wd * p; m, v EMAs; bias correction; p -= lr * mhat/(sqrt(vhat)+eps)
__global__ void train_adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                  float *__restrict__ v, size_t n, float lr, float b1, float b2, float eps, float wd,
                                  float bc1, float bc2, float gscale) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = g[i] * gscale + wd * p[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] -= (lr / bc1) * (mi / denom);
}


// =====================================================================================================
// Fast path for the PointNet layers 2/3 (97 % of the step's FLOPs): single source, 1x1 conv, rows map 1:1
// (tap 0, stride 1, no pixel shuffle), channels and columns multiples of 16 / 64.  128 x 64 x 16 tiles,
// 8 x 4 micro-tiles, float4 global loads along the contiguous axis, register prefetch of the next tile.
// Same arithmetic (fp32 FMA, same on-the-fly BN / ReLU forward and backward) as the generic kernels above.
// =====================================================================================================
constexpr int FP_BM = 128, FP_BN = 64, FP_BK = 16, FP_THREADS = 256;

#define FP_COMPUTE(As_, Bs_)                                                                        \
    _Pragma("unroll") for (int kk = 0; kk < FP_BK; ++kk) {                                          \
        const float4 a0 = *(const float4 *)&As_[kk][ty * 8], a1 = *(const float4 *)&As_[kk][ty * 8 + 4]; \
        const float4 b0 = *(const float4 *)&Bs_[kk][tx * 4];                                        \
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};                       \
        const float b[4] = {b0.x, b0.y, b0.z, b0.w};                                                \
        _Pragma("unroll") for (int i = 0; i < 8; ++i)                                               \
            _Pragma("unroll") for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);  \
    }

__device__ __forceinline__ bool plain_layer(const fcn_train_layer &L) {
    const fcn_train_seg &g = L.seg[0];
    return L.n_seg == 1 && g.tap == 0 && g.stride == 1 && g.src.up == 1 && L.up == 1 && g.src.T == L.T_out &&
           g.s_ci == 1 && g.C % FP_BK == 0 && L.N % FP_BN == 0 && g.src.c0 == 0 && g.src.ld % 4 == 0;
}

// Y[M,N] = act(X[M,K]) * W^T   (W in parameter layout (N,K): both operands contiguous along K)
__global__ void __launch_bounds__(FP_THREADS)
train_fwd_plain_kernel(const __grid_constant__ fcn_train_layer L) {
    __shared__ __align__(16) float As[FP_BK][FP_BM + 4], Bs[FP_BK][FP_BN + 4];
    __shared__ float s_scale[512], s_shift[512];
    __shared__ float s_sum[FP_BN], s_sq[FP_BN];
    const fcn_train_seg &g = L.seg[0];
    const fcn_train_src &s = g.src;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int r0 = blockIdx.x * FP_BM, n0 = blockIdx.y * FP_BN;
    const int M = L.B * L.T_out, K = g.C;
    for (int c = tid; c < K; c += FP_THREADS) {
        float sc = 1.f, sh = 0.f;
        if (s.sums != nullptr) {
            const BnCoef b = bn_coef(s.sums, s.gamma, s.beta, s.Cstat, (s.coff + c) % s.Cstat, 1.0 / s.count);
            sc = b.scale; sh = b.shift;
        }
        s_scale[c] = sc; s_shift[c] = sh;
    }
    if (tid < FP_BN) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
    __syncthreads();
    const bool act = s.sums != nullptr, relu = s.relu != 0;
    // A: 128 rows x 16 k = 512 float4 (2 per thread); B: 64 n x 16 k = 256 float4 (1 per thread)
    const int arow = tid / 4, akq = (tid % 4) * 4;
    const int bn = tid / 4, bkq = (tid % 4) * 4;
    float4 pa[2], pb;
    auto gload = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = r0 + arow + 64 * h;
            pa[h] = r < M ? *(const float4 *)(s.raw + (size_t)r * s.ld + k0 + akq) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        pb = *(const float4 *)(L.W + g.w_off + (size_t)(n0 + bn) * L.s_co + k0 + bkq);
    };
    auto sstore = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float v[4] = {pa[h].x, pa[h].y, pa[h].z, pa[h].w};
            const bool ok = r0 + arow + 64 * h < M;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float x = v[i];
                if (act) {
                    x = fmaf(x, s_scale[k0 + akq + i], s_shift[k0 + akq + i]);
                    if (relu) x = fmaxf(x, 0.f);
                }
                As[akq + i][arow + 64 * h] = ok ? x : 0.f;
            }
        }
        Bs[bkq + 0][bn] = pb.x; Bs[bkq + 1][bn] = pb.y; Bs[bkq + 2][bn] = pb.z; Bs[bkq + 3][bn] = pb.w;
    };
    float acc[8][4] = {};
    gload(0);
    for (int k0 = 0; k0 < K; k0 += FP_BK) {
        __syncthreads();
        sstore(k0);
        __syncthreads();
        if (k0 + FP_BK < K) gload(k0 + FP_BK);
        FP_COMPUTE(As, Bs)
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        float ps = 0.f, pq = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = r0 + ty * 8 + i;
            if (r < M) {
                const float y = acc[i][j];
                L.Y[(size_t)r * L.N + n] = y;
                ps += y; pq += y * y;
            }
        }
        atomicAdd(&s_sum[tx * 4 + j], ps);
        atomicAdd(&s_sq[tx * 4 + j], pq);
    }
    __syncthreads();
    if (tid < FP_BN) {
        atomicAdd(&L.sums[n0 + tid], (double)s_sum[tid]);
        atomicAdd(&L.sums[L.Cout + n0 + tid], (double)s_sq[tid]);
    }
}

// dY(m, n) for 4 consecutive n (BN/ReLU backward on the fly), coefficient arrays in shared memory
struct DyCoef {
    float scale, shift, mean, invstd, m1, m2;
};
__device__ __forceinline__ float dy_apply(const DyCoef &c, float da, float y, bool relu) {
    const float dz = (relu && fmaf(y, c.scale, c.shift) <= 0.f) ? 0.f : da;
    return c.scale * (dz - c.m1 - (y - c.mean) * c.invstd * c.m2);
}

// dX[M,K] = dY[M,N] * W   (W (N,K): contiguous along K);  plain store (single consumer)
__global__ void __launch_bounds__(FP_THREADS)
train_dx_plain_kernel(const __grid_constant__ fcn_train_layer L) {
    __shared__ __align__(16) float As[FP_BK][FP_BM + 4], Bs[FP_BK][FP_BN + 4];
    __shared__ DyCoef s_c[512];
    const fcn_train_seg &g = L.seg[0];
    const fcn_train_src &s = g.src;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int r0 = blockIdx.x * FP_BM, c0 = blockIdx.y * FP_BN;
    const int M = L.B * L.T_out, N = L.N;
    for (int n = tid; n < N; n += FP_THREADS) {
        const DyCtx c = dy_ctx(L, n);
        s_c[n] = {c.scale, c.shift, c.mean, c.invstd, c.m1, c.m2};
    }
    __syncthreads();
    const bool relu = L.relu != 0;
    const int arow = tid / 4, anq = (tid % 4) * 4;          // dY tile: 128 rows x 16 n
    const int bnr = tid / 16, bkq = (tid % 16) * 4;         // W tile: 16 n x 64 k
    float4 pda[2], py[2], pb;
    auto gload = [&](int n0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = r0 + arow + 64 * h;
            if (r < M) {
                pda[h] = *(const float4 *)(L.dA + (size_t)r * N + n0 + anq);
                py[h] = *(const float4 *)(L.Y + (size_t)r * N + n0 + anq);
            } else {
                pda[h] = py[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        pb = *(const float4 *)(L.W + g.w_off + (size_t)(n0 + bnr) * L.s_co + c0 + bkq);
    };
    auto sstore = [&](int n0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float da[4] = {pda[h].x, pda[h].y, pda[h].z, pda[h].w}, y[4] = {py[h].x, py[h].y, py[h].z, py[h].w};
            const bool ok = r0 + arow + 64 * h < M;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                As[anq + i][arow + 64 * h] = ok ? dy_apply(s_c[n0 + anq + i], da[i], y[i], relu) : 0.f;
        }
        *(float4 *)&Bs[bnr][bkq] = pb;
    };
    float acc[8][4] = {};
    gload(0);
    for (int n0 = 0; n0 < N; n0 += FP_BK) {
        __syncthreads();
        sstore(n0);
        __syncthreads();
        if (n0 + FP_BK < N) gload(n0 + FP_BK);
        FP_COMPUTE(As, Bs)
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = r0 + ty * 8 + i;
        if (r < M)
            *(float4 *)(s.grad + (size_t)r * s.ld + c0 + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

// dW[N,K] += dY[M,N]^T * act(X[M,K])   over the row slice [z*rows_per_block, ...): atomicAdd into the bucket
__global__ void __launch_bounds__(FP_THREADS)
train_dw_plain_kernel(const __grid_constant__ fcn_train_layer L, int rows_per_block) {
    __shared__ __align__(16) float As[FP_BK][FP_BM + 4], Bs[FP_BK][FP_BN + 4];   // As[row][n], Bs[row][k]
    __shared__ DyCoef s_c[FP_BM];
    __shared__ float s_scale[FP_BN], s_shift[FP_BN];
    const fcn_train_seg &g = L.seg[0];
    const fcn_train_src &s = g.src;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int n0 = blockIdx.x * FP_BM, c0 = blockIdx.y * FP_BN;
    const int M = L.B * L.T_out, N = L.N;
    const int rb = blockIdx.z * rows_per_block, re = min(M, rb + rows_per_block);
    if (tid < FP_BM) {
        const int n = min(n0 + tid, N - 1);
        const DyCtx c = dy_ctx(L, n);
        s_c[tid] = {c.scale, c.shift, c.mean, c.invstd, c.m1, c.m2};
    }
    if (tid < FP_BN) {
        float sc = 1.f, sh = 0.f;
        if (s.sums != nullptr) {
            const BnCoef b = bn_coef(s.sums, s.gamma, s.beta, s.Cstat, (s.coff + c0 + tid) % s.Cstat, 1.0 / s.count);
            sc = b.scale; sh = b.shift;
        }
        s_scale[tid] = sc; s_shift[tid] = sh;
    }
    __syncthreads();
    const bool relu = L.relu != 0, act = s.sums != nullptr, srelu = s.relu != 0;
    // dY tile: 16 rows x 128 n = 512 float4 (2 per thread); X tile: 16 rows x 64 k = 256 float4 (1 per thread)
    const int ar = tid / 32, anq = (tid % 32) * 4;
    const int br = tid / 16, bkq = (tid % 16) * 4;
    float4 pda[2], py[2], px;
    const bool n_ok = n0 + anq < N;      // N % 64 == 0: a 128-wide tile may hang over by 64
    auto gload = [&](int m0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = m0 + ar + 8 * h;
            if (r < re && n_ok) {
                pda[h] = *(const float4 *)(L.dA + (size_t)r * N + n0 + anq);
                py[h] = *(const float4 *)(L.Y + (size_t)r * N + n0 + anq);
            } else {
                pda[h] = py[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        const int r = m0 + br;
        px = r < re ? *(const float4 *)(s.raw + (size_t)r * s.ld + c0 + bkq) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto sstore = [&](int m0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const bool ok = m0 + ar + 8 * h < re && n_ok;
            float4 o;
            o.x = ok ? dy_apply(s_c[anq + 0], pda[h].x, py[h].x, relu) : 0.f;
            o.y = ok ? dy_apply(s_c[anq + 1], pda[h].y, py[h].y, relu) : 0.f;
            o.z = ok ? dy_apply(s_c[anq + 2], pda[h].z, py[h].z, relu) : 0.f;
            o.w = ok ? dy_apply(s_c[anq + 3], pda[h].w, py[h].w, relu) : 0.f;
            *(float4 *)&As[ar + 8 * h][anq] = o;
        }
        const bool ok = m0 + br < re;
        float x[4] = {px.x, px.y, px.z, px.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (act) {
                x[i] = fmaf(x[i], s_scale[bkq + i], s_shift[bkq + i]);
                if (srelu) x[i] = fmaxf(x[i], 0.f);
            }
            if (!ok) x[i] = 0.f;
        }
        *(float4 *)&Bs[br][bkq] = make_float4(x[0], x[1], x[2], x[3]);
    };
    float acc[8][4] = {};
    gload(rb);
    for (int m0 = rb; m0 < re; m0 += FP_BK) {
        __syncthreads();
        sstore(m0);
        __syncthreads();
        if (m0 + FP_BK < re) gload(m0 + FP_BK);
        FP_COMPUTE(As, Bs)
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = n0 + ty * 8 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            atomicAdd(&L.dW[g.w_off + (size_t)n * L.s_co + c0 + tx * 4 + j], acc[i][j]);
    }
}

static int check_layer(const fcn_train_layer &L) {
    FCN_REQUIRE(L.B >= 1 && L.T_out >= 1 && L.N >= 1 && L.Cout >= 1 && L.up >= 1, "bad layer shape");
    FCN_REQUIRE(L.n_seg >= 1 && L.n_seg <= FCN_MAX_SEGS, "n_seg out of range");
    FCN_REQUIRE(L.W && L.Y, "NULL weight / output");
    FCN_REQUIRE(!L.has_bn || (L.sums && L.gamma && L.beta), "BN layer without statistics buffers");
    for (int i = 0; i < L.n_seg; ++i) {
        const fcn_train_src &s = L.seg[i].src;
        FCN_REQUIRE(s.raw && s.T >= 1 && s.up >= 1 && s.ld >= 1 && L.seg[i].C >= 1 && L.seg[i].stride >= 1, "bad segment");
        FCN_REQUIRE(s.sums == nullptr || (s.gamma && s.beta && s.Cstat >= 1 && s.count > 0), "bad source statistics");
    }
    return FCN_OK;
}

static bool host_plain(const fcn_train_layer &L, bool need_src_grad_layout) {
    const fcn_train_seg &g = L.seg[0];
    const bool ok = L.n_seg == 1 && g.tap == 0 && g.stride == 1 && g.src.up == 1 && L.up == 1 && g.src.T == L.T_out &&
                    g.s_ci == 1 && g.C % 64 == 0 && g.C <= 512 && L.N % 64 == 0 && L.N <= 512 && g.src.c0 == 0 &&
                    g.src.ld % 4 == 0 && g.w_off % 4 == 0 && L.s_co % 4 == 0 && L.has_bn && L.bias == nullptr &&
                    L.Cout == L.N;
    (void)need_src_grad_layout;
    return ok;
}

}  // namespace fcn

using namespace fcn;

// K slices of the generic forward GEMM for this layer (1 = no split); `workspace_floats` bounds the partial sums
static int fwd_slices(const fcn_train_layer &L, long long workspace_floats, int *tiles_per_slice) {
    const int M = L.B * L.T_out;
    int tiles = 0;
    for (int i = 0; i < L.n_seg; ++i) tiles += ceil_div(L.seg[i].C, TG_BK);
    const int ctas = ceil_div(M, TG_BM) * ceil_div(L.N, TG_BN);
    int want = ceil_div(2 * 148, ctas);                      // ~2 CTAs per SM in total
    if (want > tiles / 2) want = tiles / 2;                  // at least two K tiles per slice
    const long long fit = workspace_floats / ((long long)M * L.N);
    if (want > fit) want = (int)fit;
    if (want < 2) { *tiles_per_slice = tiles; return 1; }
    *tiles_per_slice = ceil_div(tiles, want);
    return ceil_div(tiles, *tiles_per_slice);
}

extern "C" long long fcn_train_workspace_floats(const fcn_train_layer *L) {
    // enough for the split the launcher would like to use: 2*148 CTAs' worth of partial tiles, at most K/32 slices
    if (L == nullptr) return 0;
    int tps = 0;
    const int n = fwd_slices(*L, (long long)1 << 40, &tps);
    return n > 1 ? (long long)n * L->B * L->T_out * L->N : 0;
}

extern "C" int fcn_train_forward(const fcn_train_layer *L, float *workspace, long long workspace_floats,
                                 fcn_stream_t stream) {
    FCN_REQUIRE(L != nullptr, "NULL layer");
    if (int rc = check_layer(*L)) return rc;
    const int M = L->B * L->T_out;
    cudaStream_t st = (cudaStream_t)stream;
    if (host_plain(*L, false)) {
        dim3 grid(ceil_div(M, FP_BM), L->N / FP_BN);
        train_fwd_plain_kernel<<<grid, FP_THREADS, 0, st>>>(*L);
        FCN_LAUNCH_CHECK();
        return FCN_OK;
    }
    int tps = 0;
    const int nslice = workspace != nullptr ? fwd_slices(*L, workspace_floats, &tps) : (fwd_slices(*L, 0, &tps), 1);
    dim3 grid(ceil_div(M, TG_BM), ceil_div(L->N, TG_BN), nslice);
    train_fwd_kernel<<<grid, TG_THREADS, 0, st>>>(*L, workspace, tps);
    FCN_LAUNCH_CHECK();
    if (nslice > 1) {
        const int rpb = 32;
        dim3 gf(ceil_div(L->N, 32), ceil_div(M, rpb));
        train_fwd_finish_kernel<<<gf, 256, 0, st>>>(*L, workspace, nslice, rpb);
        FCN_LAUNCH_CHECK();
    }
    return FCN_OK;
}

extern "C" int fcn_train_backward(const fcn_train_layer *L, int need_dx_mask, fcn_stream_t stream) {
    FCN_REQUIRE(L != nullptr, "NULL layer");
    if (int rc = check_layer(*L)) return rc;
    FCN_REQUIRE(L->dA && L->dW && L->dsums, "NULL gradient buffers");
    const int M = L->B * L->T_out;
    cudaStream_t st = (cudaStream_t)stream;
    {   // Sum(dz), Sum(dz * xh) per channel
        const int rpb = M >= 4096 ? 256 : 32;
        dim3 grid(ceil_div(L->N, 32), ceil_div(M, rpb));
        train_reduce_kernel<<<grid, 256, 0, st>>>(*L, rpb);
        FCN_LAUNCH_CHECK();
    }
    if (host_plain(*L, true)) {
        const int rpb = 1024;
        dim3 grid(ceil_div(L->N, FP_BM), L->seg[0].C / FP_BN, ceil_div(M, rpb));
        train_dw_plain_kernel<<<grid, FP_THREADS, 0, st>>>(*L, rpb);
        FCN_LAUNCH_CHECK();
        if ((need_dx_mask & 1) && L->seg[0].src.grad != nullptr) {
            dim3 gx(ceil_div(M, FP_BM), L->seg[0].C / FP_BN);
            train_dx_plain_kernel<<<gx, FP_THREADS, 0, st>>>(*L);
            FCN_LAUNCH_CHECK();
        }
        return FCN_OK;
    }
    for (int sg = 0; sg < L->n_seg; ++sg) {
        // split the reduction (rows for dW, columns for dX) until ~2 CTAs per SM exist: the layers are skinny and
        // a handful of CTAs with long serial loops was the measured bottleneck (outputs are atomic accumulations)
        const int cw = ceil_div(L->seg[sg].C, TG_BM) * ceil_div(L->N, TG_BN);
        int zs = ceil_div(2 * 148, cw);
        int rpb = ceil_div(ceil_div(M, zs), TG_BK) * TG_BK;
        if (rpb < 2 * TG_BK) rpb = 2 * TG_BK;
        dim3 grid(ceil_div(L->seg[sg].C, TG_BM), ceil_div(L->N, TG_BN), ceil_div(M, rpb));
        train_dw_kernel<<<grid, TG_THREADS, 0, st>>>(*L, sg, rpb);
        FCN_LAUNCH_CHECK();
        if (((need_dx_mask >> sg) & 1) && L->seg[sg].src.grad != nullptr) {
            const int cx = ceil_div(M, TG_BM) * ceil_div(L->seg[sg].C, TG_BN);
            int zx = ceil_div(2 * 148, cx);
            int nps = ceil_div(ceil_div(L->N, zx), TG_BK) * TG_BK;
            if (nps < 2 * TG_BK) nps = 2 * TG_BK;
            dim3 gx(ceil_div(M, TG_BM), ceil_div(L->seg[sg].C, TG_BN), ceil_div(L->N, nps));
            train_dx_kernel<<<gx, TG_THREADS, 0, st>>>(*L, sg, nps);
            FCN_LAUNCH_CHECK();
        }
    }
    return FCN_OK;
}

extern "C" int fcn_train_pool(const fcn_train_pool_args *P, int backward, fcn_stream_t stream) {
    FCN_REQUIRE(P != nullptr, "NULL args");
    FCN_REQUIRE(P->B >= 1 && P->T >= 1 && P->K >= 1 && P->C >= 1 && P->V >= 0, "bad shape");
    FCN_REQUIRE(P->Y && P->cnt && P->argmax && P->sums && P->gamma && P->beta, "NULL pointer");
    dim3 grid(P->B * P->T, ceil_div(P->C + P->V, 128));
    if (!backward) {
        FCN_REQUIRE(P->feat && (P->V == 0 || P->one_hot), "NULL feature / one-hot pointer");
        train_pool_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*P);
    } else {
        FCN_REQUIRE(P->dA && P->dfeat, "NULL gradient pointer");
        train_pool_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*P);
    }
    FCN_LAUNCH_CHECK();
    return FCN_OK;
}

extern "C" int fcn_train_finalize(const fcn_train_layer *layers_dev, int n_layers, int update_running,
                                  fcn_stream_t stream) {
    FCN_REQUIRE(layers_dev != nullptr && n_layers >= 1, "bad layer table");
    train_finalize_kernel<<<n_layers, 128, 0, (cudaStream_t)stream>>>(layers_dev, n_layers, update_running);
    FCN_LAUNCH_CHECK();
    return FCN_OK;
}

extern "C" int fcn_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                             fcn_stream_t stream) {
    FCN_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "bad arguments");
    if (n == 0) return FCN_OK;
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    const int threads = 256;
    train_adam_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        param, grad, exp_avg, exp_avg_sq, (size_t)n, lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale);
    FCN_LAUNCH_CHECK();
    return FCN_OK;
}
