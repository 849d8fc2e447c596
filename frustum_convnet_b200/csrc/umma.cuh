// Hand-written sm_100a building blocks: tcgen05 (UMMA) TF32 MMA with TMEM accumulators,
// mbarrier pipelines and 1-D bulk (TMA engine, UBLKCP) global->shared copies.
//
// Shared-memory operand layout used everywhere in this library ("K-major, 128-byte swizzle"):
//   a tile is [rows][32 tf32] = rows x 128 B; 8-row groups are 1024 B apart (SBO) and inside a
//   group the 16-byte chunk c of row r is stored at chunk position (c ^ (r & 7)).
//   One tcgen05.mma (kind::tf32) consumes K = 8 elements = 32 B, so a 128-B row holds 4 k-steps;
//   stepping k advances the descriptor start address by 32 B.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Persistent PointNet kernels use only as many CTAs (clusters) as their number of rounds needs: the kernel takes the
// same ceil(n / slots) rounds, but the SMs it does not need go to the other kernels in flight (measured, car B=32,
// 8 streams: 283.0 k vs 279.3 k frustums/s; the round-1 candidates "hand-pipelined SIMT loads", "5 weight stages in
// the 2-CTA kernel" and "late PDL trigger" measured within +-1 % and were dropped).
#ifndef FCN_BALANCE_ROUNDS
#define FCN_BALANCE_ROUNDS 1
#endif

namespace fcn {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(
                     smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Non-blocking probe (mbarrier.test_wait never suspends the thread, unlike try_wait).
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a protocol bug traps (-> cudaErrorLaunchFailure) instead of hanging the GPU.
// (Keep this loop minimal: an out-of-line diagnostics call inside it slowed every kernel by 20-80 %.)
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

// ------------------------------------------------------------------ bulk copy (TMA engine, 1-D)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------ TMEM
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_result) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(smem_result)),
                 "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets lane (base_lane + i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
          "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
          "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
          "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// 32 lanes x 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

// ------------------------------------------------------------------ descriptors
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//  [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major) | [32,46) SBO>>4 = 1024>>4 |
//  [46,48) version = 1 | [49,52) base offset = 0 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major
// (cute::UMMA::InstrDescriptor: c_format[4,6)=1, a_format[7,10)=2, b_format[10,13)=2,
//  a_major[15]=0, b_major[16]=0, n_dim[17,23)=N>>3, m_dim[24,29)=M>>4).
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                     smem_u32(bar))
                 : "memory");
}

// Round fp32 to TF32 (10-bit mantissa), nearest with ties away from zero — bit-identical to
// cvt.rna.tf32.f32 for finite inputs, but done with two integer ALU ops: the F2F conversion runs on
// the quarter-rate XU pipe and was the measured bottleneck of the epilogues (61 % XU utilisation).
__device__ __forceinline__ float to_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// byte offset of element (row r, k) inside a [rows][32] K-major SW128 tile (k in [0,32))
__device__ __forceinline__ uint32_t sw128_offset(int r, int k) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + ((k & 3) << 2));
}

}  // namespace umma
// Segmented max over 32 consecutive rows held in registers (v[r] = row r of this thread's channel).
// `em` (warp-uniform) has bit r set when row r closes a section; `run` carries the open section's max in from
// the previous 32-row group and out to the next one.  Two code shapes (measured on B200, car workload):
//   PER_ROW = true : one running max, a (uniform) branch per row to the emit block - best for short sections
//                    (32/64 samples per centre: pointnet_s1 11.5 vs 16.8 us, s2 10.4 vs 13.5, s3 22.7 vs 23.9);
//   PER_ROW = false: per section one pass of R2P-predicated FMNMX over all 32 rows - no per-row branch
//                    (~20 clk of issue latency each with two warps per scheduler); best for long sections
//                    (128 samples per centre: pointnet_s4 41.6 -> 38.0 us).
template <bool PER_ROW, typename Emit>
__device__ __forceinline__ void section_max32(const uint32_t (&v)[32], unsigned em, float &run, Emit emit) {
    if (PER_ROW) {
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            run = fmaxf(run, __uint_as_float(v[r]));
            if ((em >> r) & 1u) {
                emit(run, r);
                run = -INFINITY;
            }
        }
        return;
    }
    int start = 0;
    while (em) {
        const int end = __ffs(em) - 1;
        em &= em - 1;
        const unsigned mask = (2u << end) - (1u << start);   // rows start..end (end == 31 wraps to ~0 << start)
        float m = run;
#pragma unroll
        for (int r = 0; r < 32; ++r)
            if ((mask >> r) & 1u) m = fmaxf(m, __uint_as_float(v[r]));
        emit(m, end);
        run = -INFINITY;
        start = end + 1;
    }
    if (start < 32) {                                        // open section continues in the next group
        const unsigned mask = ~0u << start;
#pragma unroll
        for (int r = 0; r < 32; ++r)
            if ((mask >> r) & 1u) run = fmaxf(run, __uint_as_float(v[r]));
    }
}

// Same contract as section_max32, organised for FEW instructions (the 16-warp epilogue of pointnet_tc2 is issue-bound):
// 8-row blocks without a section end take a 4+2+1+1 max tree; only blocks that contain an end walk their rows.
template <typename Emit>
__device__ __forceinline__ void section_max32_blocks(const uint32_t (&v)[32], unsigned em, float &run, Emit emit) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const unsigned e8 = (em >> (8 * b)) & 0xffu;
        if (e8 == 0) {                                   // warp-uniform
            const float m0 = fmaxf(__uint_as_float(v[8 * b + 0]), __uint_as_float(v[8 * b + 1]));
            const float m1 = fmaxf(__uint_as_float(v[8 * b + 2]), __uint_as_float(v[8 * b + 3]));
            const float m2 = fmaxf(__uint_as_float(v[8 * b + 4]), __uint_as_float(v[8 * b + 5]));
            const float m3 = fmaxf(__uint_as_float(v[8 * b + 6]), __uint_as_float(v[8 * b + 7]));
            run = fmaxf(run, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                run = fmaxf(run, __uint_as_float(v[8 * b + r]));
                if ((e8 >> r) & 1u) {
                    emit(run, 8 * b + r);
                    run = -INFINITY;
                }
            }
        }
    }
}

// Work units `n` over at most `slots` persistent CTAs (clusters): rounds = ceil(n / slots) is fixed by the
// grid; ceil(n / rounds) slots are enough to finish in that many rounds.  Returns the unit stride to use.
__host__ __device__ inline int balanced_stride(int n, int slots) {
    if (!FCN_BALANCE_ROUNDS || n <= 0 || slots <= 0) return slots;
    const int rounds = (n + slots - 1) / slots;
    return (n + rounds - 1) / rounds;
}

}  // namespace fcn
