// PointNet tile kernel, TF32 tensor-core variant ("precision = 1"): tcgen05.mma (kind::tf32) with
// TMEM accumulators, weight tiles staged by the TMA engine (1-D bulk copies of pre-swizzled
// images, mbarrier ring), warp-specialised roles.
//
// Per 128-row tile (rows = unique grouped points {dx,dy,dz,section} from group_rows_kernel):
//   compute warps : layer 1 (3->C1, fp32 FMA, folded BN, ReLU) -> A operand in shared memory
//                   (K-major, 128B swizzle, values rounded to TF32 with cvt.rna)
//   MMA warp      : layer 2  D2[128 x C2] = A1 * W2^T          (accumulators in TMEM cols [0,C2))
//   compute warps : TMEM -> registers, +bias, ReLU, cvt.rna -> A operand (same buffer)
//   MMA warp      : layer 3  D3[128 x 128-col chunk] = A2 * W3^T, double-buffered TMEM chunks
//   compute warps : TMEM -> registers, +bias, ReLU, segmented max over the rows of each section
//                   (redux.sync per column), integer atomicMax into the position-major feature map
//   loader warp   : streams (or keeps resident) the pre-swizzled W2/W3 stage images.
// Replaces /root/reference/models/det_base.py:95-101 + the torch.max of :134-143 for one scale.
#include "common.cuh"
#include "umma.cuh"

namespace fcn {
using namespace umma;

constexpr int TC_ROWS = 128;
constexpr int TC_COMPUTE_WARPS = 8;
constexpr int TC_THREADS = (TC_COMPUTE_WARPS + 2) * 32;
constexpr int TC_STAGE_BYTES = 16384;

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

template <int C1, int C2, int C3>
struct TcCfg {
    static constexpr int KB1 = C1 / 32, KB2 = C2 / 32;
    static constexpr int KBMAX = KB1 > KB2 ? KB1 : KB2;
    static constexpr int N2 = C2 < 128 ? C2 : 128;
    static constexpr int NCH2 = C2 / N2;
    static constexpr int N3 = 128;
    static constexpr int NCH3 = C3 / N3;
    static constexpr int A_BYTES = TC_ROWS * (C1 > C2 ? C1 : C2) * 4;
    static constexpr int JOBS2 = NCH2 * KB1, JOBS3 = NCH3 * KB2, JOBS = JOBS2 + JOBS3;
    static constexpr int NSTAGE = (C1 >= 256) ? 3 : (C1 >= 128 ? 6 : JOBS);
    static constexpr bool RESIDENT = JOBS <= NSTAGE;
    // Small-channel scales are latency-bound per tile (tiny MMAs, long SIMT phases): run TWO CTAs per SM
    // so that one CTA's epilogue overlaps the other's loads/MMAs.  That needs <= 113 KB of shared memory
    // and <= 256 TMEM columns (single acc3 buffer: with one 128-column chunk per tile the second buffer
    // never overlapped anything).
    static constexpr int CTAS_PER_SM = (C1 <= 64) ? 2 : 1;
    static constexpr int TMEM_COLS = (CTAS_PER_SM == 2) ? 256 : 512;
    static constexpr int ACC3_COL = (CTAS_PER_SM == 2) ? 128 : 256;
    static constexpr int ACC3_BUFS = (CTAS_PER_SM == 2) ? 1 : 2;
    static_assert(CTAS_PER_SM == 1 || (NCH3 == 1 && C2 <= 128), "2 CTAs/SM layout assumes one chunk per tile");
    static constexpr int OFF_W = A_BYTES;
    // resident weights are packed tightly (layer-2 stages are only N2 x 128 B), streamed stages use 16 KB slots
    static constexpr int W_BYTES = RESIDENT ? (JOBS2 * N2 * 128 + JOBS3 * TC_STAGE_BYTES) : NSTAGE * TC_STAGE_BYTES;
    __host__ __device__ static constexpr int stage_off(int st) {
        return !RESIDENT ? st * TC_STAGE_BYTES
                         : (st < JOBS2 ? st * N2 * 128 : JOBS2 * N2 * 128 + (st - JOBS2) * TC_STAGE_BYTES);
    }
    static constexpr int OFF_RECS = OFF_W + W_BYTES;
    static constexpr int OFF_W1 = OFF_RECS + 2 * TC_ROWS * 16;       // recs are double-buffered by tile parity
    static constexpr int OFF_B2 = OFF_W1 + C1 * 16;
    static constexpr int OFF_B3 = OFF_B2 + C2 * 4;
    static constexpr int OFF_SECT = OFF_B3 + C3 * 4;                 // int sect[128]
    static constexpr int OFF_BAR = OFF_SECT + 2 * TC_ROWS * 4;
    static constexpr int NBAR = 2 * NSTAGE + KBMAX + 1 + 4;
    static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
    static constexpr int BYTES = OFF_TMEM + 16 + 1024;  // + alignment slack
    static_assert(C1 % 32 == 0 && C2 % 32 == 0 && C3 % 128 == 0, "channel counts");
    static_assert(C2 <= 256, "acc2 occupies TMEM columns [0,256)");
    static_assert(C1 == C2, "a_ready phase bookkeeping assumes the same K-block count for A1 and A2");
    static_assert(BYTES <= 232448, "exceeds the 227 KB shared-memory limit per CTA");
    static_assert(CTAS_PER_SM == 1 || 2 * (BYTES + 1024) <= 233472, "two CTAs must fit in 228 KB per SM");
};

template <int C1, int C2, int C3>
__global__ void __launch_bounds__(TC_THREADS, (TcCfg<C1, C2, C3>::CTAS_PER_SM))
pointnet_tc_kernel(const __grid_constant__ fcn_pointnet_args p) {
    using Cfg = TcCfg<C1, C2, C3>;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t *smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint8_t *sA = smem;
    uint8_t *sW = smem + Cfg::OFF_W;
    float4 *recs_all = (float4 *)(smem + Cfg::OFF_RECS);
    float4 *w1s = (float4 *)(smem + Cfg::OFF_W1);
    float *b2s = (float *)(smem + Cfg::OFF_B2);
    float *b3s = (float *)(smem + Cfg::OFF_B3);
    int *sect_all = (int *)(smem + Cfg::OFF_SECT);
    uint64_t *bars = (uint64_t *)(smem + Cfg::OFF_BAR);
    uint64_t *w_full = bars, *w_empty = bars + Cfg::NSTAGE;
    uint64_t *a_ready = bars + 2 * Cfg::NSTAGE;
    uint64_t *acc2_full = a_ready + Cfg::KBMAX;
    uint64_t *acc3_full = acc2_full + 1, *acc3_empty = acc3_full + 2;
    uint32_t *tmem_slot = (uint32_t *)(smem + Cfg::OFF_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_wait();                 // tile table / rows / feature map come from the grouping kernels
    pdl_launch_dependents();
    const int ntiles = min(*p.ntiles, p.max_tiles);
    // tile stride of the persistent loop: the grid, or (FCN_BALANCE_ROUNDS) only as many CTAs as the round
    // count needs - e.g. 2.4 rounds become 3 rounds on 80 % of the CTAs, the rest exit at once and leave
    // their SMs to the kernels of the other in-flight forwards
    const int tstride = balanced_stride(ntiles, (int)gridDim.x);
    if ((int)blockIdx.x >= ntiles || (int)blockIdx.x >= tstride) return;  // whole CTA exits: nothing was started
    const int my_tiles = (ntiles - (int)blockIdx.x + tstride - 1) / tstride;

    // ---- one-time setup
    for (int i = tid; i < C1; i += TC_THREADS)
        w1s[i] = make_float4(__ldg(p.w1t + i), __ldg(p.w1t + C1 + i), __ldg(p.w1t + 2 * C1 + i), __ldg(p.b1 + i));
    for (int i = tid; i < C2; i += TC_THREADS) b2s[i] = __ldg(p.b2 + i);
    for (int i = tid; i < C3; i += TC_THREADS) b3s[i] = __ldg(p.b3 + i);
    if (tid == 0) {
        for (int i = 0; i < Cfg::NSTAGE; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < Cfg::KBMAX; ++i) mbar_init(&a_ready[i], 4);
        mbar_init(acc2_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc3_full[i], 1); mbar_init(&acc3_empty[i], TC_COMPUTE_WARPS); }
        fence_barrier_init();
    }
    if (warp == TC_COMPUTE_WARPS) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t sA_addr = smem_u32(sA), sW_addr = smem_u32(sW);
    const int4 *tiles = (const int4 *)p.tiles;

    if (warp == TC_COMPUTE_WARPS + 1) {
        // ================= weight loader (one elected lane) =================
        if (lane == 0) {
            const int ntile_loads = Cfg::RESIDENT ? 1 : my_tiles;
            uint32_t job = 0;
            for (int it = 0; it < ntile_loads; ++it) {
                for (int j = 0; j < Cfg::JOBS; ++j, ++job) {
                    const uint32_t st = job % Cfg::NSTAGE, ph = (job / Cfg::NSTAGE) & 1;
                    mbar_wait(&w_empty[st], ph ^ 1);
                    const bool l2 = j < Cfg::JOBS2;
                    const uint32_t bytes = l2 ? Cfg::N2 * 128 : TC_STAGE_BYTES;
                    const uint8_t *src = l2 ? (const uint8_t *)p.w2_tc + (size_t)j * (Cfg::N2 * 128)
                                            : (const uint8_t *)p.w3_tc + (size_t)(j - Cfg::JOBS2) * TC_STAGE_BYTES;
                    mbar_arrive_expect_tx(&w_full[st], bytes);
                    bulk_g2s(sW + Cfg::stage_off(st), src, bytes, &w_full[st]);
                }
            }
        }
    } else if (warp == TC_COMPUTE_WARPS) {
        // ================= MMA issuer: warp-uniform control flow, elected lane issues =================
        constexpr uint32_t idesc2 = make_idesc_tf32(128, Cfg::N2);
        constexpr uint32_t idesc3 = make_idesc_tf32(Cfg::N3, TC_ROWS);   // D3^T: [N3 channels] x [128 rows]
        const uint64_t adesc0 = make_desc_sw128(sA_addr), bdesc0 = make_desc_sw128(sW_addr);
        uint32_t job = 0, chunk = 0;
        bool w_ready = false;
        // non-blocking look-ahead at the weight stage of job j (resident weights: always there after tile 0)
        auto probe_next = [&](uint32_t j) -> bool {
            if (Cfg::RESIDENT) return j >= (uint32_t)Cfg::JOBS || mbar_test_wait(&w_full[j % Cfg::NSTAGE], 0);
            return mbar_test_wait(&w_full[j % Cfg::NSTAGE], (j / Cfg::NSTAGE) & 1);
        };
        const bool dbg = p.dbg_clocks != nullptr && blockIdx.x == 0 && lane == 0;
        uint32_t dj = 0;   // debug job counter (never reset)
        for (int it = 0; it < my_tiles; ++it) {
            if (Cfg::RESIDENT) { job = 0; w_ready = it > 0; }
            // ---- layer 2
            for (int nc = 0; nc < Cfg::NCH2; ++nc) {
                for (int kb = 0; kb < Cfg::KB1; ++kb, ++job) {
                    const uint32_t st = job % Cfg::NSTAGE, ph = Cfg::RESIDENT ? 0 : (job / Cfg::NSTAGE) & 1;
                    long long t0 = 0, t1 = 0, t2 = 0;
                    if (dbg) t0 = clock64();
                    if (nc == 0) mbar_wait(&a_ready[kb], 0);
                    if (dbg) t1 = clock64();
                    if (!w_ready) mbar_wait(&w_full[st], ph);
                    w_ready = probe_next(job + 1);   // issued before the MMAs: its latency hides behind them
                    if (dbg) t2 = clock64();
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ad = adesc0 + (uint64_t)(kb * ((TC_ROWS * 128) >> 4));
                        const uint64_t bd = bdesc0 + (uint64_t)(Cfg::stage_off(st) >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            mma_tf32(tmem_base + nc * Cfg::N2, ad + 2 * k, bd + 2 * k, idesc2, (kb | k) != 0);
                        if (!Cfg::RESIDENT) mma_commit(&w_empty[st]);
                    }
                    __syncwarp();
                    if (dbg && dj < 1000) { long long *d = p.dbg_clocks + 4 * dj; d[0] = t0; d[1] = t1; d[2] = t2; d[3] = clock64(); }
                    ++dj;
                }
            }
            if (elect_one()) mma_commit(acc2_full);
            __syncwarp();
            // ---- layer 3
            for (int nc = 0; nc < Cfg::NCH3; ++nc, ++chunk) {
                const uint32_t buf = chunk % Cfg::ACC3_BUFS;
                mbar_wait(&acc3_empty[buf], ((chunk / Cfg::ACC3_BUFS) & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < Cfg::KB2; ++kb, ++job) {
                    const uint32_t st = job % Cfg::NSTAGE, ph = Cfg::RESIDENT ? 0 : (job / Cfg::NSTAGE) & 1;
                    long long t0 = 0, t1 = 0, t2 = 0;
                    if (dbg) t0 = clock64();
                    if (nc == 0) mbar_wait(&a_ready[kb], 1);
                    if (dbg) t1 = clock64();
                    if (!w_ready) mbar_wait(&w_full[st], ph);
                    w_ready = probe_next(job + 1);
                    if (dbg) t2 = clock64();
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ad = adesc0 + (uint64_t)(kb * ((TC_ROWS * 128) >> 4));
                        const uint64_t bd = bdesc0 + (uint64_t)(Cfg::stage_off(st) >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k)   // transposed product: M = channels (W3 stage), N = rows (A2)
                            mma_tf32(tmem_base + Cfg::ACC3_COL + buf * 128, bd + 2 * k, ad + 2 * k, idesc3, (kb | k) != 0);
                        if (!Cfg::RESIDENT) mma_commit(&w_empty[st]);
                    }
                    __syncwarp();
                    if (dbg && dj < 1000) { long long *d = p.dbg_clocks + 4 * dj; d[0] = t0; d[1] = t1; d[2] = t2; d[3] = clock64(); }
                    ++dj;
                }
                if (elect_one()) mma_commit(&acc3_full[buf]);
                __syncwarp();
            }
        }
    } else {
        // ================= compute / epilogue warps =================
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        const int rx = row & 7;
        uint32_t chunk = 0;
        int4 td_next = tiles[blockIdx.x];
        float4 rec_next = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h == 0 && row < td_next.z)
            rec_next = ((const float4 *)p.rows + (size_t)td_next.x * p.row_cap + td_next.y)[row];
        const bool dbgc = p.dbg_clocks != nullptr && blockIdx.x == 0 && tid == 0;
        for (int it = 0; it < my_tiles; ++it) {
            long long *dc = p.dbg_clocks + 4096 + 16 * it;
            if (dbgc) dc[0] = clock64();
            const int4 td = td_next;           // tile descriptor + this thread's record were prefetched
            const int b = td.x, nrows = td.z;  // during the previous tile (two dependent global loads)
            // staging buffers alternate with the tile parity: a warp that runs ahead into tile it+1
            // must not overwrite what slower warps still read for tile it (they meet at this barrier)
            float4 *recs = recs_all + (it & 1) * TC_ROWS;
            int *sect_s = sect_all + (it & 1) * TC_ROWS;
            float4 rec = rec_next;
            if (h == 0) {
                recs[row] = rec;
                sect_s[row] = __float_as_int(rec.w) & 0x7fffffff;
            }
            if (it + 1 < my_tiles) {
                td_next = tiles[blockIdx.x + (it + 1) * tstride];
                rec_next = make_float4(0.f, 0.f, 0.f, 0.f);
                if (h == 0 && row < td_next.z)
                    rec_next = ((const float4 *)p.rows + (size_t)td_next.x * p.row_cap + td_next.y)[row];
            }
            asm volatile("bar.sync 1, %0;\n" ::"n"(TC_COMPUTE_WARPS * 32));
            rec = recs[row];
            if (dbgc) dc[1] = clock64();
            // ---- layer 1 (fp32 FMA) -> A1, K-blocks kb = h, h+2, ...
            for (int kb = h; kb < Cfg::KB1; kb += 2) {
                uint8_t *dst = sA + kb * (TC_ROWS * 128) + row_off;
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 w = w1s[kb * 32 + c4 * 4 + j];
                        o[j] = to_tf32(fmaxf(fmaf(rec.z, w.z, fmaf(rec.y, w.y, fmaf(rec.x, w.x, w.w))), 0.f));
                    }
                    *(float4 *)(dst + ((c4 ^ rx) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[kb]);
            }
            // ---- epilogue 2: TMEM -> +bias, ReLU, TF32 -> A2 (same buffer; all layer-2 MMAs are done)
            if (dbgc) dc[2] = clock64();
            mbar_wait(acc2_full, it & 1);
            if (dbgc) dc[3] = clock64();
            tc_fence_after();
            for (int kb = h; kb < Cfg::KB2; kb += 2) {
                uint32_t v[32];
                tmem_ld32(lane_taddr + kb * 32, v);
                tmem_wait_ld();
                uint8_t *dst = sA + kb * (TC_ROWS * 128) + row_off;
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        o[j] = to_tf32(fmaxf(__uint_as_float(v[c4 * 4 + j]) + b2s[kb * 32 + c4 * 4 + j], 0.f));
                    *(float4 *)(dst + ((c4 ^ rx) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
                }
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[kb]);
            }
            // ---- epilogue 3: layer 3 is computed TRANSPOSED (D3^T = W3 * A2^T), so TMEM lane = output channel
            //      and TMEM column = tile row: after tcgen05.ld a thread holds 32 consecutive rows of ITS channel
            //      and the max over a section's rows is a register-only running max (no shared-memory
            //      transposition competing with the MMA's operand reads).  Section ends are warp-uniform;
            //      +bias, ReLU, TF32 rounding commute with the max (monotone), values >= 0 so the integer
            //      atomicMax on the float bits is exact; lanes = consecutive channels -> coalesced atomics.
            if (dbgc) dc[4] = clock64();
            int *feat = (int *)(p.out + (size_t)b * p.feat_pitch * p.ld_feat);
            for (int nc = 0; nc < Cfg::NCH3; ++nc, ++chunk) {
                const uint32_t buf = chunk % Cfg::ACC3_BUFS;
                mbar_wait(&acc3_full[buf], (chunk / Cfg::ACC3_BUFS) & 1);
                if (dbgc && nc < 5) dc[5 + 2 * nc] = clock64();
                tc_fence_after();
                const int c = nc * Cfg::N3 + q * 32 + lane;     // this thread's output channel
                const float bias = b3s[c];
                float run = -INFINITY;                           // carried across the warp's two row groups
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    const int g0 = (h * 2 + half) * 32;          // first tile row of this group
                    uint32_t v[32];
                    tmem_ld32(lane_taddr + Cfg::ACC3_COL + buf * 128 + g0, v);
                    const int rg = g0 + lane;                    // section ends inside rows [64h, 64h+64)
                    const int sg = sect_s[rg], sn = sect_s[(rg + 1) & (TC_ROWS - 1)];
                    const unsigned em = __ballot_sync(
                        0xffffffffu, rg < nrows && (rg == h * 64 + 63 || rg + 1 >= nrows || sn != sg));
                    if (dbgc && nc == 0) dc[9 + 3 * half] = clock64();
                    tmem_wait_ld();
                    if (dbgc && nc == 0) dc[10 + 3 * half] = clock64();
                    section_max32<(C1 <= 128)>(v, em, run, [&](float m, int end) {
                        const float o = to_tf32(m + bias);
                        if (o > 0.f)
                            atomicMax(feat + (size_t)sect_s[g0 + end] * p.ld_feat + c, __float_as_int(o));
                    });
                    if (dbgc && nc == 0) dc[11 + 3 * half] = clock64();
                }
                tc_fence_before();
                __syncwarp();
                if (dbgc && nc < 5) dc[6 + 2 * nc] = clock64();
                if (lane == 0) mbar_arrive(&acc3_empty[buf]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_COMPUTE_WARPS) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

template <int C1, int C2, int C3>
static int launch_tc(const fcn_pointnet_args &a, cudaStream_t stream) {
    using Cfg = TcCfg<C1, C2, C3>;
    auto kern = pointnet_tc_kernel<C1, C2, C3>;
    FCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::BYTES));
    int grid = sm_count() * Cfg::CTAS_PER_SM;
    if (grid > a.max_tiles) grid = a.max_tiles;
    if (grid < 1) return FCN_OK;
    static const int prio = env_priority("FCN_PRIO_PN");
    FCN_CUDA(launch_pdl_prio(prio, kern, dim3(grid), dim3(TC_THREADS), (size_t)Cfg::BYTES, stream, a));
    return FCN_OK;
}

int pointnet_tiles_tc(const fcn_pointnet_args &a, cudaStream_t stream) {
    FCN_REQUIRE(a.tile_rows == TC_ROWS, "the TF32 tensor-core variant needs tile_rows == 128");
    FCN_REQUIRE(!a.unpooled, "the tensor-core variant only produces the pooled feature map");
    FCN_REQUIRE(a.w2_tc && a.w3_tc, "NULL tensor-core weight image");
    if (a.C1 == 64 && a.C2 == 64 && a.C3 == 128) return launch_tc<64, 64, 128>(a, stream);
    if (a.C1 == 128 && a.C2 == 128 && a.C3 == 256) return launch_tc<128, 128, 256>(a, stream);
    if (a.C1 == 256 && a.C2 == 256 && a.C3 == 512) return launch_tc<256, 256, 512>(a, stream);
    return invalid("fcn_pointnet_tiles",
                   "unsupported (C1,C2,C3); built: (64,64,128) (128,128,256) (256,256,512)");
}

}  // namespace fcn

