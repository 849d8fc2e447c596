// PointNet tile kernel, 2-CTA (cta_group::2) TF32 tensor-core variant for the 128/256-channel scales
// ("precision = 1" when the host enables it).
//
// Why: with one CTA per SM the single MMA-issuing thread is the limit — a tcgen05.mma costs ~60 clk to
// issue and an mbarrier wait ~100-170 clk, while an M128 x N128 x K8 TF32 MMA is only 64 clk of math
// (profiles/: pointnet_s4 reaches ~38 % tensor-pipe activity).  Pairing two SMs doubles the rows per
// instruction (M = 256: each CTA keeps its own 128-row tile in its own shared memory and TMEM) and halves
// the per-CTA weight bytes (each CTA stages half of the N rows of every weight tile), which makes room for
// N = 256 instructions (128 clk of math each) in the same 16 KB stage slots.
//
// Per tile pair:   L1 (fp32 FMA, both CTAs) -> A1 | L2: D2[256 x C2] (TMEM region Ra) | epilogue 2 -> A2 |
//                  L3: chunk 0 -> Rb, chunk 1 -> Ra (aliasing the drained layer-2 accumulator) | epilogue 3.
// With two layer-3 chunks the regions swap roles every pair (Ra = R0, R1, R0, ...): the next pair's layer 2 goes to
// the region whose chunk-0 result was drained long ago, so the compute warps run  L1(next pair) BEFORE the chunk-1
// epilogue of this pair and that epilogue (and its atomics) hides behind the next pair's layer-2 MMAs.
// Cross-CTA protocol: the leader CTA (cluster rank 0) issues every MMA; both CTAs' compute warps arrive on
// the leader's A-ready / region-empty mbarriers (remote arrive through mapa), the peer relays "my half of
// the weight stage landed" to the leader's stage barrier, and the leader's tcgen05.commit multicasts
// stage-empty / accumulator-full arrivals to the same barrier offsets in both CTAs.
#include "common.cuh"
#include "umma.cuh"

namespace fcn {
using namespace umma;

constexpr int T2_ROWS = 128;
constexpr int T2_COMPUTE_WARPS = 16;         // 4 per TMEM lane quadrant: the epilogues are latency-bound, not issue-bound
constexpr int T2_THREADS = (T2_COMPUTE_WARPS + 2) * 32;
constexpr int T2_STAGE_BYTES = 16384;          // per CTA: half of a [256 x 128 B] weight tile

template <int C1, int C2, int C3>
struct Tc2Cfg {
    static constexpr int KB1 = C1 / 32, KB2 = C2 / 32, KBMAX = KB1 > KB2 ? KB1 : KB2;
    static constexpr int N2 = C2;                       // one layer-2 chunk (C2 <= 256)
    static constexpr int N3 = 256, NCH3 = C3 / N3;      // layer-3 chunks of 256 columns
    static constexpr int JOBS2 = KB1, JOBS3 = NCH3 * KB2, JOBS = JOBS2 + JOBS3;
    static constexpr int HALF2 = (N2 / 2) * 128, HALF3 = (N3 / 2) * 128;   // bytes per CTA per job
    static constexpr int A_BYTES = T2_ROWS * (C1 > C2 ? C1 : C2) * 4;
    // weight-stage ring depth (5 stages at 256 channels measured +-1 %: 3 kept)
    static constexpr int NSTAGE = (C1 >= 256) ? 3 : 6;
    static constexpr int OFF_W = A_BYTES;
    static constexpr int OFF_RECS = OFF_W + NSTAGE * T2_STAGE_BYTES;
    static constexpr int OFF_W1 = OFF_RECS + 2 * T2_ROWS * 16;
    static constexpr int OFF_B2 = OFF_W1 + C1 * 16;
    static constexpr int OFF_B3 = OFF_B2 + C2 * 4;
    static constexpr int OFF_SECT = OFF_B3 + C3 * 4;
    static constexpr int OFF_BAR = OFF_SECT + 2 * 2 * T2_ROWS * 4;   // int sect[parity][tile of the pair][128]
    static constexpr int NBAR = 2 * NSTAGE + 3 * KBMAX + 1 + 2 + 2;
    static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
    static constexpr int BYTES = OFF_TMEM + 16 + 1024;
    static_assert(C1 == C2 && C2 <= 256 && C3 % 256 == 0 && NCH3 <= 2, "supported shapes");
    static_assert(BYTES <= 232448, "exceeds the 227 KB shared-memory limit per CTA");
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar, uint32_t rank) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mma_tf32_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (after all prior MMAs of this thread) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma_commit_2cta(uint64_t *bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ bool t2_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

template <int C1, int C2, int C3>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
pointnet_tc2_kernel(const __grid_constant__ fcn_pointnet_args p) {
    using Cfg = Tc2Cfg<C1, C2, C3>;
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
    uint64_t *a1_ready = bars + 2 * Cfg::NSTAGE;          // leader only, one arrival per compute warp of both CTAs
    uint64_t *a2_ready = a1_ready + Cfg::KBMAX;
    uint64_t *a_free = a2_ready + Cfg::KBMAX;             // both CTAs, multicast commit: last chunk's MMAs read A2[kb]
    uint64_t *acc2_full = a_free + Cfg::KBMAX;            // both CTAs, multicast commit
    uint64_t *acc3_full = acc2_full + 1;                  // [2] per chunk, both CTAs
    uint64_t *r_empty = acc3_full + 2;                    // [2] TMEM regions, leader only, 16 arrivals
    uint32_t *tmem_slot = (uint32_t *)(smem + Cfg::OFF_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;
    pdl_wait();
    pdl_launch_dependents();
    const int ntiles = min(*p.ntiles, p.max_tiles);
    const int npairs = (ntiles + 1) >> 1;
    const int nclusters = balanced_stride(npairs, (int)(gridDim.x >> 1));   // pair stride (see umma.cuh)
    if (cluster_id >= npairs || cluster_id >= nclusters) return;            // both CTAs of the pair agree
    const int my_pairs = (npairs - cluster_id + nclusters - 1) / nclusters;

    for (int i = tid; i < C1; i += T2_THREADS)
        w1s[i] = make_float4(__ldg(p.w1t + i), __ldg(p.w1t + C1 + i), __ldg(p.w1t + 2 * C1 + i), __ldg(p.b1 + i));
    for (int i = tid; i < C2; i += T2_THREADS) b2s[i] = __ldg(p.b2 + i);
    for (int i = tid; i < C3; i += T2_THREADS) b3s[i] = __ldg(p.b3 + i);
    if (tid == 0) {
        for (int i = 0; i < Cfg::NSTAGE; ++i) {
            mbar_init(&w_full[i], rank == 0 ? 2 : 1);      // own loader (+tx bytes) [+ the peer's relay]
            mbar_init(&w_empty[i], 1);
        }
        for (int i = 0; i < Cfg::KBMAX; ++i) {
            mbar_init(&a1_ready[i], 2 * T2_COMPUTE_WARPS);
            mbar_init(&a2_ready[i], 2 * T2_COMPUTE_WARPS);
            mbar_init(&a_free[i], 1);
        }
        mbar_init(acc2_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc3_full[i], 1); mbar_init(&r_empty[i], 2 * T2_COMPUTE_WARPS); }
        fence_barrier_init();
    }
    cluster_sync_all();          // both CTAs of the pair are resident before the paired TMEM allocation
    if (warp == T2_COMPUTE_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                     "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // barriers of both CTAs are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t sA_addr = smem_u32(sA), sW_addr = smem_u32(sW);
    const int4 *tiles = (const int4 *)p.tiles;

    if (warp == T2_COMPUTE_WARPS + 1) {
        // ================= weight loader: this CTA's half of every stage =================
        if (lane == 0) {
            uint32_t job = 0;
            for (int it = 0; it < my_pairs; ++it) {
                for (int j = 0; j < Cfg::JOBS; ++j, ++job) {
                    const uint32_t st = job % Cfg::NSTAGE, ph = (job / Cfg::NSTAGE) & 1;
                    mbar_wait(&w_empty[st], ph ^ 1);
                    const bool l2 = j < Cfg::JOBS2;
                    const uint32_t bytes = l2 ? Cfg::HALF2 : Cfg::HALF3;
                    const uint8_t *src = l2 ? (const uint8_t *)p.w2_tc + (size_t)j * (2 * Cfg::HALF2) + rank * Cfg::HALF2
                                            : (const uint8_t *)p.w3_tc + (size_t)(j - Cfg::JOBS2) * (2 * Cfg::HALF3) +
                                                  rank * Cfg::HALF3;
                    mbar_arrive_expect_tx(&w_full[st], bytes);
                    bulk_g2s(sW + st * T2_STAGE_BYTES, src, bytes, &w_full[st]);
                }
            }
        }
    } else if (warp == T2_COMPUTE_WARPS) {
        if (rank != 0) {
            // ================= peer: relay "my half of stage st landed" to the leader =================
            if (lane == 0) {
                uint32_t job = 0;
                for (int it = 0; it < my_pairs; ++it)
                    for (int j = 0; j < Cfg::JOBS; ++j, ++job) {
                        const uint32_t st = job % Cfg::NSTAGE, ph = (job / Cfg::NSTAGE) & 1;
                        mbar_wait(&w_full[st], ph);
                        mbar_arrive_cluster(&w_full[st], 0);
                    }
            }
        } else {
            // ================= leader: MMA issuer for both CTAs (M = 256) =================
            constexpr uint32_t idesc2 = make_idesc_tf32(256, Cfg::N2);
            constexpr uint32_t idesc3 = make_idesc_tf32(Cfg::N3, 2 * T2_ROWS);   // D3^T: [256 channels] x [2 x 128 rows]
            const uint64_t adesc0 = make_desc_sw128(sA_addr), bdesc0 = make_desc_sw128(sW_addr);
            uint32_t job = 0;
            const bool dbg = p.dbg_clocks != nullptr && blockIdx.x == 0 && lane == 0;
            for (int it = 0; it < my_pairs; ++it) {
                const uint32_t par = it & 1;
                const uint32_t ra = Cfg::NCH3 == 2 ? par : 0u, rb = ra ^ 1u;   // TMEM regions of this pair (x 256 columns)
                // ---- layer 2 -> region Ra (drained by the chunk-0 epilogue of the previous pair)
                if (Cfg::NCH3 == 2) mbar_wait(&r_empty[ra], par ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < Cfg::KB1; ++kb, ++job) {
                    const uint32_t st = job % Cfg::NSTAGE, ph = (job / Cfg::NSTAGE) & 1;
                    long long t0 = 0, t1 = 0, t2 = 0;
                    if (dbg) t0 = clock64();
                    mbar_wait(&a1_ready[kb], par);
                    if (dbg) t1 = clock64();
                    mbar_wait(&w_full[st], ph);
                    if (dbg) t2 = clock64();
                    tc_fence_after();
                    if (t2_elect_one()) {
                        const uint64_t ad = adesc0 + (uint64_t)(kb * ((T2_ROWS * 128) >> 4));
                        const uint64_t bd = bdesc0 + (uint64_t)(st * (T2_STAGE_BYTES >> 4));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            mma_tf32_2cta(tmem_base + ra * 256u, ad + 2 * k, bd + 2 * k, idesc2, (kb | k) != 0);
                        mma_commit_2cta(&w_empty[st]);
                    }
                    __syncwarp();
                    if (dbg && job < 1000) { long long *d = p.dbg_clocks + 4 * job; d[0] = t0; d[1] = t1; d[2] = t2; d[3] = clock64(); }
                }
                if (t2_elect_one()) mma_commit_2cta(acc2_full);
                __syncwarp();
                // ---- layer 3: chunk 0 -> Rb (drained by the LAST epilogue of the previous pair), chunk 1 -> Ra (the
                //      layer-2 accumulator, drained by epilogue 2: every a2_ready was waited for during chunk 0)
                for (int nc = 0; nc < Cfg::NCH3; ++nc) {
                    if (nc == 0) mbar_wait(&r_empty[rb], par ^ 1);
                    tc_fence_after();
                    const uint32_t dcol = (nc == 0 ? rb : ra) * 256u;
                    for (int kb = 0; kb < Cfg::KB2; ++kb, ++job) {
                        const uint32_t st = job % Cfg::NSTAGE, ph = (job / Cfg::NSTAGE) & 1;
                        long long t0 = 0, t1 = 0, t2 = 0;
                        if (dbg) t0 = clock64();
                        if (nc == 0) mbar_wait(&a2_ready[kb], par);
                        if (dbg) t1 = clock64();
                        mbar_wait(&w_full[st], ph);
                        if (dbg) t2 = clock64();
                        tc_fence_after();
                        if (t2_elect_one()) {
                            const uint64_t ad = adesc0 + (uint64_t)(kb * ((T2_ROWS * 128) >> 4));
                            const uint64_t bd = bdesc0 + (uint64_t)(st * (T2_STAGE_BYTES >> 4));
#pragma unroll
                            for (int k = 0; k < 4; ++k)   // transposed: M = 256 channels (W3), N = 2 x 128 rows (A2)
                                mma_tf32_2cta(tmem_base + dcol, bd + 2 * k, ad + 2 * k, idesc3, (kb | k) != 0);
                            mma_commit_2cta(&w_empty[st]);
                            if (nc == Cfg::NCH3 - 1) mma_commit_2cta(&a_free[kb]);   // A2[kb] may be refilled
                        }
                        __syncwarp();
                        if (dbg && job < 1000) { long long *d = p.dbg_clocks + 4 * job; d[0] = t0; d[1] = t1; d[2] = t2; d[3] = clock64(); }
                    }
                    if (t2_elect_one()) mma_commit_2cta(&acc3_full[nc]);
                    __syncwarp();
                }
            }
        }
    } else {
        // ================= compute / epilogue warps (both CTAs, own 128-row tile) =================
        // warp = q + 4 * grp: q = TMEM lane quadrant (rows 32q.. of the tile / channels 32q.. of the transposed layer 3),
        // grp = 0..3 splits the K blocks (layer 1, epilogue 2) and, as (h, g) = (grp & 1, grp >> 1), the columns of the
        // transposed accumulator (epilogue 3: rows [64g, 64g + 64) of tile h)
        const int q = warp & 3, grp = warp >> 2, h = grp & 1, g = grp >> 1;
        const int row = q * 32 + lane;
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        const int rx = row & 7;
        auto tile_of = [&](int it, uint32_t r) -> int4 {
            const int tile = 2 * (cluster_id + it * nclusters) + (int)r;
            return (it < my_pairs && tile < ntiles) ? tiles[tile] : make_int4(0, 0, 0, 0);   // odd count: empty partner
        };
        // g == 0 warps stage: h == 0 the records of this CTA's tile (layer 1), h == 1 the section ids of the
        // partner's tile (epilogue 3 drains BOTH tiles' rows for this CTA's half of the channels)
        auto rec_of = [&](const int4 &t) -> float4 {
            return (g == 0 && row < t.z) ? ((const float4 *)p.rows + (size_t)t.x * p.row_cap + t.y)[row]
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        const bool dbgc = p.dbg_clocks != nullptr && blockIdx.x == 0 && tid == 0;
        int4 td = tile_of(0, rank), tdp = tile_of(0, rank ^ 1u);     // own tile / partner's tile of the staged pair
        float4 rec_pre = rec_of(h == 0 ? td : tdp);
        // records + section ids of pair `it` (prefetched in td / tdp / rec_pre) into the parity buffers, then
        // layer 1 (fp32 FMA) -> A1.  `chase`: A2[kb] of the previous pair is still being read by its last MMAs.
        auto stage_and_layer1 = [&](int it, bool chase) {
            float4 *recs = recs_all + (it & 1) * T2_ROWS;
            int *sect_w = sect_all + (it & 1) * (2 * T2_ROWS);   // [tile of the pair][row]
            // a_free[0] of the previous pair: its last chunk has started, i.e. EVERY warp has finished the epilogue-3
            // pass that read this parity's section ids
            if (chase) mbar_wait(&a_free[0], (it & 1) ^ 1);
            if (g == 0) {
                if (h == 0) recs[row] = rec_pre;
                sect_w[((h == 0) ? rank : (rank ^ 1u)) * T2_ROWS + row] = __float_as_int(rec_pre.w) & 0x7fffffff;
            }
            asm volatile("bar.sync 1, %0;\n" ::"n"(T2_COMPUTE_WARPS * 32));
            const float4 rec = recs[row];
            // every warp computes ITS 8 channels (two 16-byte units) of every K block, in the order the MMAs consume
            // them: a K block is complete ~1/4 of a block time after its buffer slice is free
            // (two K blocks per publication.  Measured: the cost of a publication hardly depends on how the loads, FMAs
            // and the proxy fence are arranged - the 128 KB of A1 stores compete with the MMA operand reads and the
            // weight fills for the SM's shared-memory data path, which this kernel keeps ~90 % busy)
            long long *dl = p.dbg_clocks + 6144 + 32 * it;
            if (dbgc) dl[0] = clock64();
#pragma unroll 1
            for (int kb0 = 0; kb0 < Cfg::KB1; kb0 += 2) {
                if (chase) mbar_wait(&a_free[kb0 + 1], (it & 1) ^ 1);     // in-order MMAs: implies a_free[kb0]
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int kb = kb0 + i;
                    uint8_t *dst = sA + kb * (T2_ROWS * 128) + row_off;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int c4 = 2 * grp + u;
                        float o[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 w = w1s[kb * 32 + c4 * 4 + j];
                            o[j] = to_tf32(fmaxf(fmaf(rec.z, w.z, fmaf(rec.y, w.y, fmaf(rec.x, w.x, w.w))), 0.f));
                        }
                        *(float4 *)(dst + ((c4 ^ rx) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                }
                fence_proxy_async_all();
                __syncwarp();
                if (lane < 2) mbar_arrive_cluster(&a1_ready[kb0 + lane], 0);
                if (dbgc) dl[1 + (kb0 >> 1)] = clock64();
            }
        };
        stage_and_layer1(0, false);
        for (int it = 0; it < my_pairs; ++it) {
            const uint32_t par = it & 1;
            const uint32_t ra = Cfg::NCH3 == 2 ? par : 0u, rb = ra ^ 1u;
            long long *dc = p.dbg_clocks + 4096 + 16 * it;
            if (dbgc) dc[0] = clock64();
            const int *sect_s = sect_all + (it & 1) * (2 * T2_ROWS);
            // epilogue 3 drains, per warp, rows of tile h of the pair: keep that tile's record, then prefetch the next
            // pair's (its global loads fly during epilogue 2)
            const int4 tdt = ((uint32_t)h == rank) ? td : tdp;
            td = tile_of(it + 1, rank);
            tdp = tile_of(it + 1, rank ^ 1u);
            rec_pre = rec_of(h == 0 ? td : tdp);
            // ---- epilogue 2: TMEM Ra -> +bias, ReLU, TF32 -> A2 (same buffer)
            mbar_wait(acc2_full, par);
            if (dbgc) dc[3] = clock64();
            tc_fence_after();
            // every warp converts ITS 8 columns of every K block (four blocks per TMEM load batch)
#pragma unroll 1
            for (int kb0 = 0; kb0 < Cfg::KB2; kb0 += 4) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 4; ++i) tmem_ld8(lane_taddr + ra * 256u + (kb0 + i) * 32 + grp * 8, v + 8 * i);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int kb = kb0 + i;
                    uint8_t *dst = sA + kb * (T2_ROWS * 128) + row_off;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int c4 = 2 * grp + u;
                        const float4 b = *(const float4 *)(b2s + kb * 32 + c4 * 4);
                        *(float4 *)(dst + ((c4 ^ rx) << 4)) =
                            make_float4(to_tf32(fmaxf(__uint_as_float(v[8 * i + 4 * u + 0]) + b.x, 0.f)),
                                        to_tf32(fmaxf(__uint_as_float(v[8 * i + 4 * u + 1]) + b.y, 0.f)),
                                        to_tf32(fmaxf(__uint_as_float(v[8 * i + 4 * u + 2]) + b.z, 0.f)),
                                        to_tf32(fmaxf(__uint_as_float(v[8 * i + 4 * u + 3]) + b.w, 0.f)));
                    }
                    if (i & 1) {                 // publish two K blocks per proxy fence
                        tc_fence_before();
                        fence_proxy_async_all();
                        __syncwarp();
                        if (lane < 2) mbar_arrive_cluster(&a2_ready[kb - 1 + lane], 0);
                    }
                }
            }
            // ---- epilogue 3: layer 3 is computed TRANSPOSED (D3^T = W3 * A2^T): TMEM lane = output channel
            //      (this CTA owns channels [128*rank, 128*rank+128) of the 256-wide chunk), TMEM column = row
            //      (columns [0,128) = tile of rank 0, [128,256) = tile of rank 1).  Warp (q,h,g) drains channels
            //      32q.. for rows [64g, 64g+64) of tile h: a thread holds 32 consecutive rows of ITS channel, the
            //      section max is a register-only running max (carried across the two 32-row groups; a section that
            //      crosses row 64 is finished by two atomics), the section ends are warp-uniform, and lanes =
            //      consecutive channels -> coalesced atomics.
            if (dbgc) dc[4] = clock64();
            const int nrows_t = tdt.z;
            int *feat = (int *)(p.out + (size_t)tdt.x * p.feat_pitch * p.ld_feat);
            const int *sect_t = sect_s + h * T2_ROWS;
            for (int nc = 0; nc < Cfg::NCH3; ++nc) {
                const uint32_t reg = nc == 0 ? rb : ra;
                if (nc == Cfg::NCH3 - 1 && it + 1 < my_pairs) {
                    // layer 1 of the NEXT pair refills the operand buffer K block by K block behind the last chunk's
                    // MMAs; its layer-2 MMAs (into the other TMEM region) then run under the drain below
                    stage_and_layer1(it + 1, true);
                    if (dbgc) dc[9] = clock64();
                }
                mbar_wait(&acc3_full[nc], par);
                if (dbgc) dc[5 + 2 * nc] = clock64();
                tc_fence_after();
                const int c = nc * Cfg::N3 + (int)rank * T2_ROWS + q * 32 + lane;   // this thread's output channel
                const float bias = b3s[c];
                float run = -INFINITY;
#pragma unroll 1
                for (int part = 2 * g; part < 2 * g + 2; ++part) {
                    const int g0 = part * 32;
                    uint32_t v[32];
                    tmem_ld32(lane_taddr + reg * 256u + h * T2_ROWS + g0, v);
                    const int rg = g0 + lane;
                    const int sg = sect_t[rg], sn = sect_t[(rg + 1) & (T2_ROWS - 1)];
                    const unsigned em = __ballot_sync(
                        0xffffffffu, rg < nrows_t && (rg == 64 * g + 63 || rg + 1 >= nrows_t || sn != sg));
                    tmem_wait_ld();
                    if (dbgc && nc == 0) dc[10 + 2 * (part & 1)] = clock64();
                    section_max32_blocks(v, em, run, [&](float m, int end) {
                        const float o = to_tf32(m + bias);
                        if (o > 0.f)
                            atomicMax(feat + (size_t)sect_t[g0 + end] * p.ld_feat + c, __float_as_int(o));
                    });
                    if (dbgc && nc == 0) dc[11 + 2 * (part & 1)] = clock64();
                }
                tc_fence_before();
                __syncwarp();
                if (dbgc) dc[6 + 2 * nc] = clock64();
                // region drained (waited for by the leader's MMA warp before it is written again)
                if (lane == 0) mbar_arrive_cluster(&r_empty[reg], 0);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // neither CTA may exit (or free TMEM) while the partner can still touch it
    if (warp == T2_COMPUTE_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

template <int C1, int C2, int C3>
static int launch_tc2(const fcn_pointnet_args &a, cudaStream_t stream) {
    using Cfg = Tc2Cfg<C1, C2, C3>;
    auto kern = pointnet_tc2_kernel<C1, C2, C3>;
    FCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::BYTES));
    int grid = sm_count() & ~1;
    const int max_pairs2 = 2 * ((a.max_tiles + 1) / 2);
    if (grid > max_pairs2) grid = max_pairs2;
    if (grid < 2) return FCN_OK;
    static const int prio = env_priority("FCN_PRIO_PN");
    FCN_CUDA(launch_pdl_prio(prio, kern, dim3(grid), dim3(T2_THREADS), (size_t)Cfg::BYTES, stream, a));
    return FCN_OK;
}

// precision 3 (host opt-in): 2-CTA clusters for the 128/256-channel scales.
int pointnet_tiles_tc2(const fcn_pointnet_args &a, cudaStream_t stream) {
    FCN_REQUIRE(a.tile_rows == T2_ROWS, "the TF32 tensor-core variant needs tile_rows == 128");
    FCN_REQUIRE(!a.unpooled, "the tensor-core variant only produces the pooled feature map");
    FCN_REQUIRE(a.w2_tc && a.w3_tc, "NULL tensor-core weight image");
    if (a.C1 == 128 && a.C2 == 128 && a.C3 == 256) return launch_tc2<128, 128, 256>(a, stream);
    if (a.C1 == 256 && a.C2 == 256 && a.C3 == 512) return launch_tc2<256, 256, 512>(a, stream);
    return invalid("fcn_pointnet_tiles", "2-CTA variant built for (128,128,256) and (256,256,512) only");
}

}  // namespace fcn
