// Persistent FCN kernel: ALL conv / transposed-conv / merge layers of ConvFeatNet, the two heads and the eval
// decode of one forward in ONE launch (replaces the 14 fcn_conv_gemm launches + fcn_decode_eval;
// /root/reference/models/det_base.py:196-224,367-411).
//
// Why: every per-layer launch paid ~8 us of fixed cost per CTA (launch/PDL wait, barrier + TMEM set-up, pipeline
// fill, drain) for 1-6 us of tensor-core work on grids of 9-72 CTAs - 0.14 of the 0.27 ms forward latency and
// ~44 SM-us per step for 10 % of the FLOPs.
//
// Design
//   * work unit ("job") = one [128 rows x NT columns] output tile of one layer; the host builds the job table in
//     TOPOLOGICAL order (every job depends only on jobs with a smaller index);
//   * persistent CTAs fetch the next job with ONE atomicAdd (dynamic scheduling).  Together with the topological
//     order this is deadlock-free under ANY co-residency: the smallest unfinished job is always owned by a running
//     CTA whose dependencies are complete - no grid-wide barrier, no co-residency requirement, late CTAs
//     (several forwards are in flight on other streams) simply pick up what is left;
//   * layer-to-layer dependencies are per row tile: the epilogue of a tile publishes a counter
//     (release, gpu scope); the TMA producer of a consumer tile acquires the counters of the (<= 4) source
//     tiles it reads (taps +-1, stride 2, pixel-shuffled transposed convs), then crosses to the async proxy;
//   * inside a CTA the same warp-specialised tcgen05 pipeline as conv_gemm_tma.cu, made continuous across
//     jobs: warp 0 = scheduler + TMA producer (A boxes by cp.async.bulk.tensor with the conv taps / stride in
//     the coordinates, weight stage images by bulk copies), warp 1 = MMA issuer (kind::tf32, M = 128,
//     N = 128 | 64), warps 2-5 = epilogue; TWO TMEM accumulators so the epilogue of job j overlaps the K loop of
//     job j+1; the smem stage ring runs across job boundaries (the next job's weights stream while the current
//     one still multiplies);
//   * the heads tile (NT = 64: all 2 + 39 logits of a position in one thread's TMEM lane) stores the logits row
//     and decodes it in place (decode.cuh - the same function the stand-alone decode kernel uses), writing the
//     6-tuple to the local result block AND to the peers' gather buffers over NVLink (multi-GPU: no NCCL
//     all-gather, no extra launch; one epoch flag per forward);
//   * self-cleaning: the last CTA to finish resets the job counter and the tile flags for the next forward.
#include <cuda.h>

#include "common.cuh"
#include "decode.cuh"
#include "umma.cuh"

namespace fcn {
using namespace umma;

constexpr int MG_ROWS = 128;
constexpr int MG_THREADS = 6 * 32;   // warp 0: scheduler + TMA producer, warp 1: MMA, warps 2-5: epilogue
constexpr int MG_NSTAGE = 3;         // 64-wide K stages: 2 x 16 KB of A + 32 KB of W each
constexpr int MG_A_ATOM = MG_ROWS * 128, MG_A_STAGE = 2 * MG_A_ATOM;
constexpr int MG_W_STAGE = 2 * 128 * 128;                     // sized for NT = 128
constexpr int MG_OFF_W = MG_NSTAGE * MG_A_STAGE;
constexpr int MG_OFF_BAR = MG_OFF_W + MG_NSTAGE * MG_W_STAGE;
constexpr int MG_JOBQ = 4;                                    // job descriptors in flight inside a CTA
constexpr int MG_NBAR = 2 * MG_NSTAGE + 4 + 2 * MG_JOBQ;      // full/empty, acc_full/acc_empty [2], job_full/job_empty
constexpr int MG_OFF_JOBQ = MG_OFF_BAR + MG_NBAR * 8;
constexpr int MG_OFF_TMEM = MG_OFF_JOBQ + MG_JOBQ * 4;
constexpr int MG_BYTES = MG_OFF_TMEM + 16 + 1024;
static_assert(MG_BYTES <= 232448, "exceeds the 227 KB shared-memory limit per CTA");

struct MegaParams {
    fcn_mega_args a;
};

__device__ __forceinline__ bool mg_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mg_tma_load_3d(uint32_t dst, const void *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
        "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(int *p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// generic-proxy writes of other SMs (observed through the acquire above) -> subsequent async-proxy (TMA) reads
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;\n" ::: "memory"); }

__global__ void __launch_bounds__(MG_THREADS, 1)
fcn_mega_kernel(const __grid_constant__ MegaParams P) {
    const fcn_mega_args &p = P.a;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t *smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint8_t *sA = smem, *sW = smem + MG_OFF_W;
    uint64_t *bars = (uint64_t *)(smem + MG_OFF_BAR);
    uint64_t *full = bars, *empty = bars + MG_NSTAGE;
    uint64_t *acc_full = bars + 2 * MG_NSTAGE, *acc_empty = acc_full + 2;
    uint64_t *job_full = acc_empty + 2, *job_empty = job_full + MG_JOBQ;
    volatile int *jobq = (volatile int *)(smem + MG_OFF_JOBQ);
    uint32_t *tmem_slot = (uint32_t *)(smem + MG_OFF_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < MG_NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        for (int i = 0; i < MG_JOBQ; ++i) { mbar_init(&job_full[i], 1); mbar_init(&job_empty[i], 5); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                 // feature maps come from the PointNet kernels; prologue above overlapped their tail
    pdl_launch_dependents();

    int *job_counter = p.sync, *done_ctas = p.sync + 1, *epoch = p.sync + 2, *flags = p.sync + 4;

    if (warp == 0) {
        // ================= scheduler + TMA producer (one lane) =================
        if (lane == 0) {
            uint32_t stage = 0;                       // running K-stage counter across jobs
            for (uint32_t q = 0;; ++q) {
                const uint32_t slot = q % MG_JOBQ;
                mbar_wait(&job_empty[slot], ((q / MG_JOBQ) & 1) ^ 1);
                int j = atomicAdd(job_counter, 1);
                if (j >= p.n_jobs) j = -1;
                jobq[slot] = j;
                mbar_arrive(&job_full[slot]);         // release: the queue entry is visible to the waiters
                if (j < 0) break;
                const fcn_mega_job &job = p.jobs[j];
                const fcn_mega_layer &L = p.layers[job.layer];
                // dependencies: the row tiles of the producing layers this tile reads
                for (int d = 0; d < job.n_dep; ++d) {
                    const int first = job.dep[d].first, cnt = job.dep[d].count, target = job.dep[d].target;
                    for (int i = 0; i < cnt; ++i) {
                        uint32_t spins = 0;
                        while (ld_acquire_gpu(flags + first + i) < target) {
                            if (++spins > (1u << 24)) __trap();   // protocol bug -> launch failure, not a hang
                        }
                    }
                }
                if (job.n_dep > 0) fence_proxy_async_global();
                const int r0 = job.m_tile * MG_ROWS;
                const int NS = L.n_stage;
                const uint32_t w_bytes = (uint32_t)L.NT * 256u;          // [NT rows x 64 tf32] per stage
                const uint8_t *wsrc = (const uint8_t *)L.w_tc + (size_t)job.n_tile * NS * w_bytes;
                int seg = 0, kbi = 0;                 // running (segment, 32-channel block inside the segment)
                for (int s = 0; s < NS; ++s, ++stage) {
                    const uint32_t st = stage % MG_NSTAGE, ph = (stage / MG_NSTAGE) & 1;
                    mbar_wait(&empty[st], ph ^ 1);
                    mbar_arrive_expect_tx(&full[st], MG_A_STAGE + w_bytes);
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        const uint32_t dst = smem_u32(sA) + st * MG_A_STAGE + a * MG_A_ATOM;
                        if (seg < L.n_seg) {
                            const fcn_mega_seg &sg = L.seg[seg];
                            mg_tma_load_3d(dst, (const uint8_t *)p.tmaps + 128 * (size_t)sg.map_idx, kbi * 32,
                                           r0 * sg.stride + sg.tap, 0, &full[st]);
                            if (++kbi >= sg.kblocks) { kbi = 0; ++seg; }
                        } else {   // K padding block: a box fully outside the channel range -> zeros
                            mg_tma_load_3d(dst, (const uint8_t *)p.tmaps + 128 * (size_t)L.seg[0].map_idx, 1 << 20, 0, 0,
                                           &full[st]);
                        }
                    }
                    bulk_g2s(sW + st * MG_W_STAGE, wsrc + (size_t)s * w_bytes, w_bytes, &full[st]);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: warp-uniform control flow, elected lane issues =================
        const uint64_t adesc0 = make_desc_sw128(smem_u32(sA));
        const uint64_t bdesc0 = make_desc_sw128(smem_u32(sW));
        uint32_t stage = 0;
        for (uint32_t q = 0;; ++q) {
            const uint32_t slot = q % MG_JOBQ;
            mbar_wait(&job_full[slot], (q / MG_JOBQ) & 1);
            const int j = jobq[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(&job_empty[slot]);
            if (j < 0) break;
            const fcn_mega_layer &L = p.layers[p.jobs[j].layer];
            const int NS = L.n_stage, NT = L.NT;
            const uint32_t idesc = make_idesc_tf32(128, NT);
            const uint32_t w_atom16 = (uint32_t)(NT * 128) >> 4;          // second 32-wide K atom of a W stage
            const uint32_t buf = q & 1;
            mbar_wait(&acc_empty[buf], ((q >> 1) & 1) ^ 1);               // epilogue of job q-2 drained this buffer
            tc_fence_after();
            const uint32_t dtmem = tmem_base + buf * 128;
            for (int s = 0; s < NS; ++s, ++stage) {
                const uint32_t st = stage % MG_NSTAGE, ph = (stage / MG_NSTAGE) & 1;
                mbar_wait(&full[st], ph);
                tc_fence_after();
                if (mg_elect_one()) {
                    const uint64_t ad = adesc0 + (uint64_t)(st * (MG_A_STAGE >> 4));
                    const uint64_t bd = bdesc0 + (uint64_t)(st * (MG_W_STAGE >> 4));
#pragma unroll
                    for (int a = 0; a < 2; ++a)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            mma_tf32(dtmem, ad + (uint64_t)(a * (MG_A_ATOM >> 4) + 2 * k),
                                     bd + (uint64_t)(a * w_atom16 + 2 * k), idesc, (s | a | k) != 0);
                    mma_commit(&empty[st]);
                }
                __syncwarp();
            }
            if (mg_elect_one()) mma_commit(&acc_full[buf]);
            __syncwarp();
        }
    } else {
        // ================= epilogue: TMEM -> +bias (+ReLU, TF32 rounding) -> position-major store ==========
        const int qd = warp & 3;                      // TMEM lane quadrant this warp may access
        for (uint32_t q = 0;; ++q) {
            const uint32_t slot = q % MG_JOBQ;
            mbar_wait(&job_full[slot], (q / MG_JOBQ) & 1);
            const int j = jobq[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(&job_empty[slot]);
            if (j < 0) break;
            const fcn_mega_job &job = p.jobs[j];
            const fcn_mega_layer &L = p.layers[job.layer];
            const int NT = L.NT;
            const uint32_t buf = q & 1;
            const int r = job.m_tile * MG_ROWS + qd * 32 + lane;    // flattened GEMM row of this thread
            const int b = r / L.P_m, rt = r - b * L.P_m;            // (frustum, position)
            const bool row_ok = r < L.n_rows && rt < L.T_out;
            mbar_wait(&acc_full[buf], (q >> 1) & 1);
            tc_fence_after();
            const uint32_t lane_taddr = tmem_base + buf * 128 + ((uint32_t)(qd * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < NT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(lane_taddr + c0, v);
                tmem_wait_ld();
                const int n = job.n_tile * NT + c0;
                if (!row_ok || n >= L.up * L.Cout) continue;
                const int jj = n / L.Cout, co = n - jj * L.Cout;
                const int tt = rt * L.up + jj;
                if (tt >= L.T_store) continue;
                float *out = L.out + ((size_t)b * L.P_store + tt) * L.ld_out + L.c_off + co;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 bb = __ldg((const float4 *)(L.bias + n + c * 4));
                    float4 o = make_float4(__uint_as_float(v[c * 4]) + bb.x, __uint_as_float(v[c * 4 + 1]) + bb.y,
                                           __uint_as_float(v[c * 4 + 2]) + bb.z, __uint_as_float(v[c * 4 + 3]) + bb.w);
                    if (L.relu) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    }
                    if (L.round_out) { o.x = to_tf32(o.x); o.y = to_tf32(o.y); o.z = to_tf32(o.z); o.w = to_tf32(o.w); }
                    *(float4 *)(out + c * 4) = o;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);   // TMEM buffer free for job q+2
            if (L.is_heads) {
                // fused eval decode: this thread just stored the complete logits row of (b, rt); decode it in place
                // (same thread -> program order) into the local block and the peers' gather buffers
                if (row_ok)
                    decode_row(L.out + ((size_t)b * L.P_store + rt) * L.ld_out, b * p.T + rt, b, rt, p.T, p.NH, p.NS,
                               p.center_ref, p.mean_size, (const DecodeOut *)p.outs, p.n_out);
                if (p.n_out > 1) __threadfence_system();   // remote (NVLink) stores before the completion count
            }
            // publish: all rows of this warp are stored (gpu scope) -> one count per epilogue warp
            __threadfence();
            __syncwarp();
            if (lane == 0) red_release_gpu_add(flags + L.flag_base + job.m_tile, 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
    // self-cleaning: the last CTA out resets the scheduler state for the next forward and, on the multi-GPU
    // path, raises this rank's epoch flag in every peer (all result rows were fenced at system scope above)
    __shared__ int s_last;
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(done_ctas, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int i = tid; i < p.n_flags; i += MG_THREADS) flags[i] = 0;
        __syncthreads();
        if (tid == 0) {
            *job_counter = 0;
            *done_ctas = 0;
            const int e = *epoch + 1;
            *epoch = e;
            __threadfence_system();
            for (int o = 0; o < p.n_flag_out; ++o)
                if (p.flag_out[o] != nullptr) *(volatile int *)p.flag_out[o] = e;
            __threadfence_system();
        }
    }
}

}  // namespace fcn

using namespace fcn;

extern "C" int fcn_mega_forward(const fcn_mega_args *args, fcn_stream_t stream) {
    FCN_REQUIRE(args != nullptr, "args is NULL");
    const fcn_mega_args &a = *args;
    FCN_REQUIRE(a.n_jobs >= 0 && a.n_layers >= 1 && a.n_flags >= 0, "bad table sizes");
    FCN_REQUIRE(a.layers && a.jobs && a.tmaps && a.sync, "NULL table pointer");
    FCN_REQUIRE(((uintptr_t)a.tmaps & 63) == 0, "tensor maps must be 64-byte aligned");
    FCN_REQUIRE(a.n_out >= 0 && a.n_out <= FCN_MAX_PEERS && a.n_flag_out >= 0 && a.n_flag_out <= FCN_MAX_PEERS,
                "too many output sets");
    FCN_REQUIRE(a.NH >= 1 && a.NH <= DEC_MAX_BINS && a.NS >= 1 && a.NS <= DEC_MAX_BINS, "bad decode sizes");
    if (a.n_jobs == 0) return FCN_OK;
    static_assert(sizeof(DecodeOut) == sizeof(fcn_decode_out), "decode output block layout");
    int grid = a.grid > 0 ? a.grid : sm_count();
    if (grid > a.n_jobs) grid = a.n_jobs;
    if (grid > sm_count()) grid = sm_count();
    FCN_CUDA(cudaFuncSetAttribute(fcn_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_BYTES));
    MegaParams P;
    P.a = a;
    static const int prio = env_priority("FCN_PRIO_CONV");
    FCN_CUDA(launch_pdl_prio(prio, fcn_mega_kernel, dim3(grid), dim3(MG_THREADS), (size_t)MG_BYTES, (cudaStream_t)stream, P));
    return FCN_OK;
}
