// Self-test of the tcgen05 building blocks in isolation (descriptor encodings, swizzled operand
// layouts, TMEM read-back): D[128 x N] = A[128 x K] * W[N x K]^T with one CTA.
// A is written to shared memory by threads (as the PointNet epilogues do), W arrives as the
// pre-swizzled stage image produced by the host packer through a bulk copy.
#include "common.cuh"
#include "umma.cuh"

namespace fcn {
using namespace umma;

template <int N>
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(int K, const float *__restrict__ A, const uint8_t *__restrict__ w_img,
                     float *__restrict__ D) {
    extern __shared__ uint8_t smem_dyn[];
    uint8_t *smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int KB = K / 32;
    uint8_t *sA = smem;                      // KB x [128][128 B]
    uint8_t *sW = sA + KB * 16384;           // KB x [N][128 B]
    uint64_t *bars = (uint64_t *)(sW + KB * N * 128);
    uint32_t *tmem_slot = (uint32_t *)(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<128>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // A: thread = row, rounded to TF32, swizzled
    for (int k = 0; k < K; ++k) {
        const float v = to_tf32(A[(size_t)tid * K + k]);
        *(float *)(sA + (k >> 5) * 16384 + sw128_offset(tid, k & 31)) = v;
    }
    fence_proxy_async();
    if (tid == 0) {
        mbar_arrive_expect_tx(&bars[0], KB * N * 128);
        bulk_g2s(sW, w_img, KB * N * 128, &bars[0]);
    }
    __syncthreads();
    if (tid == 0) {
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc_tf32(128, N);
        for (int kb = 0; kb < KB; ++kb)
            for (int k = 0; k < 4; ++k)
                mma_tf32(tmem_base, make_desc_sw128(smem_u32(sA) + kb * 16384 + k * 32),
                         make_desc_sw128(smem_u32(sW) + kb * N * 128 + k * 32), idesc, (kb | k) != 0);
        mma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_wait_ld();
        for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tmem_base);
}

}  // namespace fcn

using namespace fcn;

extern "C" int fcn_selftest_umma(int N, int K, const float *A, const void *w_img, float *D,
                                 fcn_stream_t stream) {
    FCN_REQUIRE(N == 64 || N == 128, "N must be 64 or 128");
    FCN_REQUIRE(K >= 32 && K % 32 == 0 && K <= 256, "K must be a multiple of 32 in [32,256]");
    FCN_REQUIRE(A && w_img && D, "NULL pointer");
    const int smem = (K / 32) * (16384 + N * 128) + 64 + 1024;
    if (N == 64) {
        FCN_CUDA(cudaFuncSetAttribute(umma_selftest_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        umma_selftest_kernel<64><<<1, 128, smem, (cudaStream_t)stream>>>(K, A, (const uint8_t *)w_img, D);
    } else {
        FCN_CUDA(cudaFuncSetAttribute(umma_selftest_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        umma_selftest_kernel<128><<<1, 128, smem, (cudaStream_t)stream>>>(K, A, (const uint8_t *)w_img, D);
    }
    FCN_LAUNCH_CHECK();
    return FCN_OK;
}
