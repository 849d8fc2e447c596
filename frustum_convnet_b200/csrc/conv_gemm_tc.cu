// 1-D conv / transposed conv / 1x1 conv (+ folded BN, optional ReLU) as an implicit GEMM on the
// 5th-gen tensor cores ("precision = 1|2"): tcgen05.mma kind::tf32, accumulator in TMEM, weights
// staged by the TMA engine (bulk copies of pre-swizzled [N_TILE x 32] images), activations gathered
// from the position-major maps (taps / concat segments / zero padding) with coalesced 16-byte
// cp.async (LDGSTS, zero-fill) straight into the K-major 128B-swizzled operand layout.
// Activations in HBM are already TF32-rounded by their producers (`round_out`), so the tensor
// core's operand truncation is exact.
//
// Pipeline (measured design points): one stage = 64 K elements (two 128-byte swizzle
// atoms) so that the single-thread costs — one mbarrier wait (~170 clk even when already complete),
// one proxy fence, one commit — are paid once per 8 MMAs; A and W of a stage complete on ONE mbarrier
// (128 cp.async arrivals + the bulk copy's transaction bytes); the MMA warp runs warp-uniform and
// issues through an elected lane; producers keep per-segment row pointers in registers so a K block
// costs ~16 instructions per thread.
// Same math as conv_gemm_simt.cu; replaces the Conv1d/DeConv1d/cat/head calls of
// /root/reference/models/det_base.py:196-224,367-368.
#include "common.cuh"
#include "umma.cuh"

namespace fcn {
using namespace umma;

constexpr int GT_ROWS = 128;
constexpr int GT_PROD_WARPS = 4;                 // A producers, then epilogue (thread = output row)
constexpr int GT_THREADS = (GT_PROD_WARPS + 2) * 32;

template <int NT>
struct GtCfg {
    static constexpr int NSTAGE = NT > 64 ? 3 : 4;
    static constexpr int A_ATOM = GT_ROWS * 128, W_ATOM = NT * 128;
    static constexpr int A_STAGE = 2 * A_ATOM, W_STAGE = 2 * W_ATOM;
    static constexpr int OFF_W = NSTAGE * A_STAGE;
    static constexpr int OFF_BAR = OFF_W + NSTAGE * W_STAGE;
    static constexpr int NBAR = 2 * NSTAGE + 1;
    static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
    static constexpr int BYTES = OFF_TMEM + 16 + 1024;
    static_assert(BYTES <= 232448, "exceeds the 227 KB shared-memory limit per CTA");
};

__device__ __forceinline__ bool gt_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

template <int NT>
__global__ void __launch_bounds__(GT_THREADS)
conv_gemm_tc_kernel(const __grid_constant__ fcn_conv_args p) {
    using Cfg = GtCfg<NT>;
    constexpr int NSTAGE = Cfg::NSTAGE;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t *smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint8_t *sA = smem, *sW = smem + Cfg::OFF_W;
    uint64_t *bars = (uint64_t *)(smem + Cfg::OFF_BAR);
    uint64_t *full = bars, *empty = bars + NSTAGE, *acc_full = bars + 2 * NSTAGE;
    uint32_t *tmem_slot = (uint32_t *)(smem + Cfg::OFF_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = p.B * p.P_m;
    const int m0 = blockIdx.x * GT_ROWS, n_tile = blockIdx.y;
    const int NS = p.K_pad / 64;                 // pipeline stages of 64 K elements

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&full[i], GT_PROD_WARPS * 32 + 1);   // 128 cp.async arrivals + the W loader
            mbar_init(&empty[i], 1);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == GT_PROD_WARPS) tmem_alloc<(NT < 32 ? 32 : NT)>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                 // prologue above overlapped the previous layer's tail
    pdl_launch_dependents();

    if (warp == GT_PROD_WARPS + 1) {
        // ================= weight loader (one lane) =================
        if (lane == 0) {
            const uint8_t *src = (const uint8_t *)p.w_tc + (size_t)n_tile * NS * Cfg::W_STAGE;
            for (int s = 0; s < NS; ++s) {
                const int st = s % NSTAGE, ph = (s / NSTAGE) & 1;
                mbar_wait(&empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&full[st], Cfg::W_STAGE);
                bulk_g2s(sW + st * Cfg::W_STAGE, src + (size_t)s * Cfg::W_STAGE, Cfg::W_STAGE, &full[st]);
            }
        }
    } else if (warp == GT_PROD_WARPS) {
        // ================= MMA issuer: warp-uniform control flow, elected lane issues =================
        constexpr uint32_t idesc = make_idesc_tf32(128, NT);
        const uint64_t adesc0 = make_desc_sw128(smem_u32(sA));
        const uint64_t bdesc0 = make_desc_sw128(smem_u32(sW));
        const bool dbg = p.dbg_clocks != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
        for (int s = 0; s < NS; ++s) {
            const int st = s % NSTAGE, ph = (s / NSTAGE) & 1;
            long long t0 = 0, t1 = 0;
            if (dbg) t0 = clock64();
            mbar_wait(&full[st], ph);
            if (dbg) t1 = clock64();
            tc_fence_after();
            if (gt_elect_one()) {
                fence_proxy_async();   // cp.async (generic proxy) writes -> tcgen05 (async proxy) reads
                const uint64_t ad = adesc0 + (uint64_t)(st * (Cfg::A_STAGE >> 4));
                const uint64_t bd = bdesc0 + (uint64_t)(st * (Cfg::W_STAGE >> 4));
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_tf32(tmem_base, ad + (uint64_t)(a * (Cfg::A_ATOM >> 4) + 2 * k),
                                 bd + (uint64_t)(a * (Cfg::W_ATOM >> 4) + 2 * k), idesc, (s | a | k) != 0);
                mma_commit(&empty[st]);
            }
            __syncwarp();
            if (dbg) {
                long long *d = p.dbg_clocks + (size_t)s * 8;
                d[0] = t0; d[1] = t1; d[2] = t1; d[3] = clock64();
            }
        }
        if (gt_elect_one()) mma_commit(acc_full);
        __syncwarp();
    } else {
        // ================= A producers, then epilogue =================
        // Coalesced gather: one warp instruction covers 4 rows x 128 B (lane -> row j*4 + lane/8,
        // 16-byte chunk lane%8), i.e. 4 full cache lines instead of 32 partial ones.
        const int chunk = lane & 7;
        int gb[8], gt[8];               // (frustum, position) of the 8 rows this lane feeds
        uint32_t gdst[8];
        unsigned gok = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int lr = warp * 32 + j * 4 + (lane >> 3);
            const int gr = m0 + lr;
            const int grr = gr < M ? gr : 0;
            gb[j] = grr / p.P_m;
            gt[j] = grr - gb[j] * p.P_m;
            const bool okr = gr < M && gt[j] < p.T_out;
            gok |= (okr ? 1u : 0u) << j;
            gdst[j] = smem_u32(sA) + (uint32_t)((lr >> 3) * 1024 + (lr & 7) * 128 + ((chunk ^ (lr & 7)) << 4));
        }
        const bool dbgp = p.dbg_clocks != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;
        int kbg = 0;                    // global 32-wide K-block index
        for (int seg = 0; seg < p.n_seg; ++seg) {
            const fcn_conv_seg sg = p.seg[seg];
            const float *rp[8];         // channel-0 address of the 8 source rows for this segment
            unsigned okm = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int ts = gt[j] * sg.stride + sg.tap;
                const bool ok = ((gok >> j) & 1u) && ts >= 0 && ts < sg.T_src;
                okm |= (ok ? 1u : 0u) << j;
                rp[j] = sg.src + ((size_t)gb[j] * sg.pitch + (ok ? ts : 0)) * sg.ld + chunk * 4;
            }
            const int nkb = (sg.C + 31) >> 5;
            for (int kbi = 0; kbi < nkb; ++kbi, ++kbg) {
                const int s = kbg >> 1, st = s % NSTAGE, ph = (s / NSTAGE) & 1;
                if ((kbg & 1) == 0) {
                    long long tp0 = 0;
                    if (dbgp) tp0 = clock64();
                    mbar_wait(&empty[st], ph ^ 1);
                    if (dbgp) { p.dbg_clocks[(size_t)s * 8 + 4] = tp0; p.dbg_clocks[(size_t)s * 8 + 5] = clock64(); }
                }
                const unsigned m = (kbi * 32 + chunk * 4 < sg.ld) ? okm : 0u;   // pad columns / past ld: zeros
                const uint32_t dofs = st * Cfg::A_STAGE + (kbg & 1) * Cfg::A_ATOM;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(gdst[j] + dofs),
                                 "l"(rp[j] + kbi * 32), "r"(((m >> j) & 1u) ? 16 : 0)
                                 : "memory");
                if (kbg & 1) {
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&full[st]))
                                 : "memory");
                    if (dbgp) p.dbg_clocks[(size_t)s * 8 + 6] = clock64();
                }
            }
        }
        if (kbg & 1) {   // odd number of K blocks: the second atom of the last stage is all zeros
            const int s = kbg >> 1, st = s % NSTAGE;
            const uint32_t dofs = st * Cfg::A_STAGE + Cfg::A_ATOM;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(gdst[j] + dofs), "l"(p.wt), "r"(0)
                             : "memory");
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&full[st]))
                         : "memory");
        }
        // ---- epilogue: TMEM -> +bias (+ReLU, TF32 rounding) -> position-major global store
        const int row = warp * 32 + lane;
        const int r = m0 + row;
        const int rb = r < M ? r / p.P_m : 0;
        const int rt = r < M ? r - rb * p.P_m : 0;
        const bool row_ok = r < M && rt < p.T_out;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(lane_taddr + c0, v);
            tmem_wait_ld();
            const int n = n_tile * NT + c0;
            if (!row_ok || n >= p.up * p.Cout) continue;
            const int jj = n / p.Cout, co = n - jj * p.Cout;
            const int tt = rt * p.up + jj;
            if (tt >= p.T_store) continue;
            float *out = p.out + ((size_t)rb * p.P_store + tt) * p.ld_out + p.c_off + co;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 bb = __ldg((const float4 *)(p.bias + n + c * 4));
                float4 o = make_float4(__uint_as_float(v[c * 4]) + bb.x, __uint_as_float(v[c * 4 + 1]) + bb.y,
                                       __uint_as_float(v[c * 4 + 2]) + bb.z, __uint_as_float(v[c * 4 + 3]) + bb.w);
                if (p.relu) {
                    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                }
                if (p.round_out) { o.x = to_tf32(o.x); o.y = to_tf32(o.y); o.z = to_tf32(o.z); o.w = to_tf32(o.w); }
                *(float4 *)(out + c * 4) = o;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == GT_PROD_WARPS) {
        tc_fence_after();
        tmem_dealloc<(NT < 32 ? 32 : NT)>(tmem_base);
    }
}

template <int NT>
static int launch_gt(const fcn_conv_args &a, cudaStream_t stream) {
    using Cfg = GtCfg<NT>;
    auto kern = conv_gemm_tc_kernel<NT>;
    FCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::BYTES));
    const int M = a.B * a.P_m;
    dim3 grid(ceil_div(M, GT_ROWS), a.n_cols / NT);
    FCN_CUDA(launch_pdl(kern, grid, dim3(GT_THREADS), (size_t)Cfg::BYTES, stream, a));
    return FCN_OK;
}

// precision 1: N tile 128; precision 2: N tile 64 (more CTAs for the short, wide layers).
int conv_gemm_tc(const fcn_conv_args &a, cudaStream_t stream) {
    FCN_REQUIRE(a.w_tc != nullptr, "NULL tensor-core weight image");
    FCN_REQUIRE(a.Cout % 32 == 0, "tensor-core variant needs Cout % 32 == 0");
    FCN_REQUIRE(a.K_pad % 64 == 0, "tensor-core variant needs K_pad % 64 == 0");
    if (a.B * a.T_out == 0) return FCN_OK;
    if (a.precision == 2) return launch_gt<64>(a, stream);
    FCN_REQUIRE(a.n_cols % 128 == 0, "n_cols must be a multiple of 128 for the 128-wide N tile");
    return launch_gt<128>(a, stream);
}

}  // namespace fcn
