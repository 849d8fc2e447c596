// PointNet tile kernel, fp32 SIMT variant ("precision = 0").
//
// One CTA processes tiles of 64 row records {dx,dy,dz,section} produced by group_rows_kernel and
// applies the three folded (1x1 conv + BN + ReLU) layers of PointNetModule
// (/root/reference/models/det_base.py:95-97, factories models/common.py:45-49), then either
//   pooled   : max over the rows of each section (torch.max(-1), det_base.py:134-143) combined
//              into the position-major feature map with integer atomic max (values are >= 0
//              after ReLU, so int ordering == float ordering; empty sections keep the 0 that the
//              (num > 0) mask of det_base.py:100-101 produces), or
//   unpooled : the masked (B,C3,T,K) tensor that PointNetModule.forward returns (det_base.py:103).
// Activations never leave shared memory; weights stream through a cp.async double buffer.
#include "common.cuh"

namespace fcn {

constexpr int PT_ROWS = 64;      // rows per tile
constexpr int PT_THREADS = 256;  // 16 x 16 threads, 4 x 4 outputs each per 64-column chunk
constexpr int PT_KC = 32;        // K chunk
constexpr int PT_NC = 64;        // N chunk

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// OUT[64][N] = A[64][K] * Wt[K][N]; epilogue functor gets (nc, acc) per 64-column chunk.
template <int K, int N, class Epi>
__device__ __forceinline__ void gemm64(const float *A, const float *__restrict__ Wt, float *wst,
                                       Epi epi) {
    constexpr int LDA = K + 4;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int lrow = tid >> 4, lcol = (tid & 15) * 4;  // weight-chunk loader coordinates
    for (int nc = 0; nc < N / PT_NC; ++nc) {
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
        const float *wbase = Wt + nc * PT_NC;
        cp_async16(wst + lrow * PT_NC + lcol, wbase + (size_t)lrow * N + lcol);
        cp_async16(wst + (lrow + 16) * PT_NC + lcol, wbase + (size_t)(lrow + 16) * N + lcol);
        cp_async_commit();
        for (int kc = 0; kc < K / PT_KC; ++kc) {
            float *cur = wst + (kc & 1) * (PT_KC * PT_NC);
            if (kc + 1 < K / PT_KC) {
                float *nxt = wst + ((kc + 1) & 1) * (PT_KC * PT_NC);
                const float *src = wbase + (size_t)(kc + 1) * PT_KC * N;
                cp_async16(nxt + lrow * PT_NC + lcol, src + (size_t)lrow * N + lcol);
                cp_async16(nxt + (lrow + 16) * PT_NC + lcol, src + (size_t)(lrow + 16) * N + lcol);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const float *arow = A + (ty * 4) * LDA + kc * PT_KC;
#pragma unroll
            for (int kk = 0; kk < PT_KC; kk += 4) {
                float4 a[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) a[r] = *(const float4 *)(arow + r * LDA + kk);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 w = *(const float4 *)(cur + (kk + j) * PT_NC + tx * 4);
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
        epi(nc, acc);
    }
}

template <int C1, int C2, int C3>
struct PointnetSmem {
    static constexpr int LD1 = C1 + 4, LD2 = C2 + 4;
    static constexpr size_t bytes = sizeof(float4) * PT_ROWS + sizeof(int) * PT_ROWS * 3 +
                                    sizeof(float) * (PT_ROWS * LD1 + PT_ROWS * LD2) +
                                    sizeof(float) * 2 * PT_KC * PT_NC +
                                    sizeof(int) * PT_ROWS * (PT_NC + 1);
};

template <int C1, int C2, int C3>
__global__ void __launch_bounds__(PT_THREADS)
pointnet_simt_kernel(const __grid_constant__ fcn_pointnet_args p) {
    using S = PointnetSmem<C1, C2, C3>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *recs = (float4 *)smem_raw;
    int *rank = (int *)(recs + PT_ROWS);  // dense rank of the row's section inside the tile
    int *rank_sect = rank + PT_ROWS;      // section id of each rank
    int *row_flag = rank_sect + PT_ROWS;  // 1 = valid row, 0 = padding, 2 = masked section
    float *h1 = (float *)(row_flag + PT_ROWS);
    float *h2 = h1 + PT_ROWS * S::LD1;
    float *wst = h2 + PT_ROWS * S::LD2;
    int *smax = (int *)(wst + 2 * PT_KC * PT_NC);
    __shared__ int s_nrank, s_w0;

    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    pdl_wait();
    pdl_launch_dependents();
    const int ntiles = min(*p.ntiles, p.max_tiles);
    const int4 *tiles = (const int4 *)p.tiles;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int4 td = tiles[tile];
        const int b = td.x, row0 = td.y, nrows = td.z;
        const float4 *grows = (const float4 *)p.rows + (size_t)b * p.row_cap + row0;
        if (tid < PT_ROWS) {
            float4 r = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
            if (tid < nrows) r = grows[tid];
            recs[tid] = r;
        }
        __syncthreads();
        // dense section ranks of the (section-sorted) rows: two warps, ballot prefix
        int rk = 0, sect = 0;
        bool valid = false, head = false;
        if (tid < PT_ROWS) {
            const int w = __float_as_int(recs[tid].w);
            sect = w & 0x7fffffff;
            valid = tid < nrows;
            const int prev = tid > 0 ? (__float_as_int(recs[tid - 1].w) & 0x7fffffff) : -1;
            head = valid && (tid == 0 || sect != prev);
            const unsigned m = __ballot_sync(0xffffffffu, head);
            rk = __popc(m & (0xffffffffu >> (31 - (tid & 31)))) - 1;
            if (tid == 0) s_w0 = __popc(m);
            row_flag[tid] = !valid ? 0 : (w < 0 ? 2 : 1);
        }
        __syncthreads();
        if (tid < PT_ROWS) {
            if (tid >= 32) rk += s_w0;
            rank[tid] = valid ? rk : -1;
            if (head) rank_sect[rk] = sect;
            if (tid == PT_ROWS - 1) s_nrank = rk + 1;
        }
        __syncthreads();

        // ---- layer 1: 3 -> C1 on CUDA cores
        for (int i = tid; i < PT_ROWS * (C1 / 4); i += PT_THREADS) {
            const int r = i / (C1 / 4), c = (i - r * (C1 / 4)) * 4;
            const float4 q = recs[r];
            const float4 wx = __ldg((const float4 *)(p.w1t + c));
            const float4 wy = __ldg((const float4 *)(p.w1t + C1 + c));
            const float4 wz = __ldg((const float4 *)(p.w1t + 2 * C1 + c));
            const float4 bb = __ldg((const float4 *)(p.b1 + c));
            float4 o;
            o.x = fmaxf(fmaf(q.z, wz.x, fmaf(q.y, wy.x, fmaf(q.x, wx.x, bb.x))), 0.f);
            o.y = fmaxf(fmaf(q.z, wz.y, fmaf(q.y, wy.y, fmaf(q.x, wx.y, bb.y))), 0.f);
            o.z = fmaxf(fmaf(q.z, wz.z, fmaf(q.y, wy.z, fmaf(q.x, wx.z, bb.z))), 0.f);
            o.w = fmaxf(fmaf(q.z, wz.w, fmaf(q.y, wy.w, fmaf(q.x, wx.w, bb.w))), 0.f);
            *(float4 *)(h1 + r * S::LD1 + c) = o;
        }
        __syncthreads();

        // ---- layer 2: C1 -> C2, ReLU, to shared memory
        gemm64<C1, C2>(h1, p.w2t, wst, [&](int nc, float (&acc)[4][4]) {
            const float4 bb = __ldg((const float4 *)(p.b2 + nc * PT_NC + tx * 4));
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float4 o;
                o.x = fmaxf(acc[r][0] + bb.x, 0.f);
                o.y = fmaxf(acc[r][1] + bb.y, 0.f);
                o.z = fmaxf(acc[r][2] + bb.z, 0.f);
                o.w = fmaxf(acc[r][3] + bb.w, 0.f);
                *(float4 *)(h2 + (ty * 4 + r) * S::LD2 + nc * PT_NC + tx * 4) = o;
            }
        });
        __syncthreads();

        // ---- layer 3: C2 -> C3, ReLU, then pooled / unpooled epilogue
        const int nrank = s_nrank;
        gemm64<C2, C3>(h2, p.w3t, wst, [&](int nc, float (&acc)[4][4]) {
            const float4 bb = __ldg((const float4 *)(p.b3 + nc * PT_NC + tx * 4));
            const float bias[4] = {bb.x, bb.y, bb.z, bb.w};
            if (!p.unpooled) {
                for (int i = tid; i < nrank * (PT_NC + 1); i += PT_THREADS) smax[i] = 0;
                __syncthreads();
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int row = ty * 4 + r;
                    if (row_flag[row] == 1) {
                        int *dst = smax + rank[row] * (PT_NC + 1) + tx * 4;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            atomicMax(dst + j, __float_as_int(fmaxf(acc[r][j] + bias[j], 0.f)));
                    }
                }
                __syncthreads();
                int *feat = (int *)(p.out + (size_t)b * p.feat_pitch * p.ld_feat);
                for (int i = tid; i < nrank * PT_NC; i += PT_THREADS) {
                    const int rr = i >> 6, col = i & 63;
                    const int v = smax[rr * (PT_NC + 1) + col];
                    if (v > 0)
                        atomicMax(feat + (size_t)rank_sect[rr] * p.ld_feat + nc * PT_NC + col, v);
                }
                __syncthreads();
            } else {
                const size_t TK = (size_t)p.T * p.K;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int row = ty * 4 + r;
                    const int fl = row_flag[row];
                    if (fl == 0) continue;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = nc * PT_NC + tx * 4 + j;
                        const float v = fl == 2 ? 0.f : fmaxf(acc[r][j] + bias[j], 0.f);
                        p.out[((size_t)b * C3 + c) * TK + row0 + row] = v;
                    }
                }
            }
        });
        __syncthreads();
    }
}

template <int C1, int C2, int C3>
static int launch_simt(const fcn_pointnet_args &a, cudaStream_t stream) {
    using S = PointnetSmem<C1, C2, C3>;
    auto kern = pointnet_simt_kernel<C1, C2, C3>;
    FCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
    int occ = 1;
    FCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PT_THREADS, S::bytes));
    if (occ < 1) occ = 1;
    int grid = sm_count() * occ;
    if (grid > a.max_tiles) grid = a.max_tiles;
    if (grid < 1) return FCN_OK;
    FCN_CUDA(launch_pdl(kern, dim3(grid), dim3(PT_THREADS), (size_t)S::bytes, stream, a));
    return FCN_OK;
}

int pointnet_tiles_simt(const fcn_pointnet_args &a, cudaStream_t stream) {
    FCN_REQUIRE(a.tile_rows == PT_ROWS, "the fp32 SIMT variant needs tile_rows == 64");
    if (a.C1 == 64 && a.C2 == 64 && a.C3 == 128) return launch_simt<64, 64, 128>(a, stream);
    if (a.C1 == 128 && a.C2 == 128 && a.C3 == 256) return launch_simt<128, 128, 256>(a, stream);
    if (a.C1 == 256 && a.C2 == 256 && a.C3 == 512) return launch_simt<256, 256, 512>(a, stream);
    return invalid("fcn_pointnet_tiles",
                   "unsupported (C1,C2,C3); built: (64,64,128) (128,128,256) (256,256,512)");
}

}  // namespace fcn
