// 1-D conv / transposed conv / 1x1 conv (+ folded BN, optional ReLU) as an implicit GEMM on
// position-major activations, fp32 SIMT variant ("precision = 0").
//
// Replaces the Conv1d / DeConv1d blocks of ConvFeatNet (/root/reference/models/det_base.py:167-224,
// models/common.py:38-63) and the two 1x1 heads (det_base.py:367-368).  torch.cat along channels
// (det_base.py:202,208,214,222) is expressed as extra K segments of the A operand; the kernel taps
// (k=3, pad=1, stride 1|2) are K segments with a row shift, so no im2col buffer exists in HBM.
// A transposed conv with kernel == stride (det_base.py:181-183) is a GEMM with `up*Cout` columns
// whose column group j lands on output position t*up + j.
#include "common.cuh"

namespace fcn {

constexpr int CG_TM = 64, CG_TN = 64, CG_KC = 32, CG_THREADS = 256;
constexpr int CG_LDA = CG_KC + 4;

__device__ __forceinline__ void cg_cp_async16(void *smem, const void *gmem, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = pred ? 16 : 0;  // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}

__global__ void __launch_bounds__(CG_THREADS)
conv_gemm_simt_kernel(const __grid_constant__ fcn_conv_args p) {
    __shared__ __align__(16) float sA[2][CG_TM * CG_LDA];
    __shared__ __align__(16) float sW[2][CG_KC * CG_TN];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int M = p.B * p.P_m;
    const int m0 = blockIdx.x * CG_TM, n0 = blockIdx.y * CG_TN;
    const int nchunks = p.K_pad / CG_KC;
    pdl_wait();
    pdl_launch_dependents();

    // loader coordinates: A: rows (tid>>3) and +32, 4 channels at (tid&7)*4;  W: rows (tid>>4), +16
    const int arow = tid >> 3, acol = (tid & 7) * 4;
    int ab[2], at[2];
    bool arow_ok[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = m0 + arow + h * 32;
        const int rr = r < M ? r : 0;
        ab[h] = rr / p.P_m;
        at[h] = rr - ab[h] * p.P_m;
        arow_ok[h] = r < M && at[h] < p.T_out;
    }
    const int wrow = tid >> 4, wcol = (tid & 15) * 4;

    auto load_chunk = [&](int kc, int stage) {
        // which segment does chunk kc belong to?
        int seg = 0, c0 = kc * CG_KC;
#pragma unroll
        for (int s = 0; s < FCN_MAX_SEGS; ++s) {
            if (s < p.n_seg - 1) {
                const int span = ((p.seg[s].C + CG_KC - 1) / CG_KC) * CG_KC;
                if (seg == s && c0 >= span) { c0 -= span; seg = s + 1; }
            }
        }
        const fcn_conv_seg sg = p.seg[seg];
        const int c = c0 + acol;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ts = at[h] * sg.stride + sg.tap;
            const bool ok = arow_ok[h] && ts >= 0 && ts < sg.T_src && c < sg.ld;
            const float *src = ok ? sg.src + ((size_t)ab[h] * sg.pitch + ts) * sg.ld + c : sg.src;
            cg_cp_async16(&sA[stage][(arow + h * 32) * CG_LDA + acol], src, ok);
        }
        const float *w = p.wt + (size_t)(kc * CG_KC) * p.n_cols + n0;
        cg_cp_async16(&sW[stage][wrow * CG_TN + wcol], w + (size_t)wrow * p.n_cols + wcol, true);
        cg_cp_async16(&sW[stage][(wrow + 16) * CG_TN + wcol], w + (size_t)(wrow + 16) * p.n_cols + wcol, true);
        asm volatile("cp.async.commit_group;\n");
    };

    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;

    load_chunk(0, 0);
    for (int kc = 0; kc < nchunks; ++kc) {
        const int cur = kc & 1;
        if (kc + 1 < nchunks) {
            load_chunk(kc + 1, cur ^ 1);
            asm volatile("cp.async.wait_group 1;\n");
        } else {
            asm volatile("cp.async.wait_group 0;\n");
        }
        __syncthreads();
        const float *a_base = &sA[cur][(ty * 4) * CG_LDA];
        const float *w_base = &sW[cur][tx * 4];
#pragma unroll
        for (int kk = 0; kk < CG_KC; kk += 4) {
            float4 a[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = *(const float4 *)(a_base + r * CG_LDA + kk);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 w = *(const float4 *)(w_base + (kk + j) * CG_TN);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float av = j == 0 ? a[r].x : j == 1 ? a[r].y : j == 2 ? a[r].z : a[r].w;
                    acc[r][0] = fmaf(av, w.x, acc[r][0]);
                    acc[r][1] = fmaf(av, w.y, acc[r][1]);
                    acc[r][2] = fmaf(av, w.z, acc[r][2]);
                    acc[r][3] = fmaf(av, w.w, acc[r][3]);
                }
            }
        }
        __syncthreads();
    }

    // epilogue: bias (+ReLU), scatter column group j to output position t*up + j
    const int n = n0 + tx * 4;
    if (n >= p.up * p.Cout) return;
    const float4 bb = __ldg((const float4 *)(p.bias + n));
    const int jj = n / p.Cout, co = n - jj * p.Cout;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int row = m0 + ty * 4 + r;
        if (row >= M) continue;
        const int b = row / p.P_m, t = row - b * p.P_m;
        const int tt = t * p.up + jj;
        if (t >= p.T_out || tt >= p.T_store) continue;
        float4 o = make_float4(acc[r][0] + bb.x, acc[r][1] + bb.y, acc[r][2] + bb.z, acc[r][3] + bb.w);
        if (p.relu) {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
        }
        *(float4 *)(p.out + ((size_t)b * p.P_store + tt) * p.ld_out + p.c_off + co) = o;
    }
}

int conv_gemm_simt(const fcn_conv_args &a, cudaStream_t stream) {
    const int M = a.B * a.P_m;
    if (M == 0) return FCN_OK;
    dim3 grid(ceil_div(M, CG_TM), a.n_cols / CG_TN);
    FCN_CUDA(launch_pdl(conv_gemm_simt_kernel, grid, dim3(CG_THREADS), (size_t)0, stream, a));
    return FCN_OK;
}

}  // namespace fcn
