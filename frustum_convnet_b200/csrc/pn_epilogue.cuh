// Epilogue-3 helper shared by the tensor-core PointNet kernels: segmented max over the rows (= lanes) of
// a warp for 32 accumulator columns held in registers, then one coalesced atomic-max flush per section.
//
// v[j]   : accumulator of (row = lane, column c0 + j) straight from tcgen05.ld (thread = row)
// dist   : lane distance to the first row of this lane's section inside the warp (0 for padding rows)
// endmask: bit r set when row r is the last valid row of its section inside this warp
// A Hillis-Steele scan over lanes (5 shuffle steps, 32 independent columns => full ILP) leaves the section
// maximum in the section's last lane; that lane parks its 32 values in a 128-byte shared row so that the
// warp can add the bias, apply ReLU/TF32 rounding and issue ONE 128-byte coalesced RED.MAX per section.
// (max_r relu(x_r + b) == relu(max_r x_r + b); outputs >= 0, so integer order == float order.)
#pragma once
#include "umma.cuh"

namespace fcn {

__device__ __forceinline__ void segmax_flush32(uint32_t (&v)[32], int dist, unsigned endmask, int lane,
                                               float *srow, const int *sect_q, int *feat, int ld_feat,
                                               int c_lane, float bias) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const bool take = dist >= d;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float mine = __uint_as_float(v[j]);
            const float up = __shfl_up_sync(0xffffffffu, mine, d);
            v[j] = take ? __float_as_uint(fmaxf(mine, up)) : v[j];
        }
    }
    unsigned em = endmask;
    while (em) {
        const int e = __ffs(em) - 1;
        em &= em - 1;
        if (lane == e) {
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4)
                *(uint4 *)(srow + c4 * 4) = make_uint4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
        }
        __syncwarp();
        const float o = umma::to_tf32(srow[lane] + bias);   // monotone: max of rounded == rounded max
        if (o > 0.f) atomicMax(feat + (size_t)sect_q[e] * ld_feat + c_lane, __float_as_int(o));
        __syncwarp();
    }
}

// dist for segmax_flush32: rows are section-sorted; `start` marks the first row of a section in this warp
__device__ __forceinline__ int segment_dist(bool valid, bool start, int lane) {
    const unsigned sm = __ballot_sync(0xffffffffu, valid && start);
    const unsigned below = sm & (0xffffffffu >> (31 - lane));
    return (valid && below) ? lane - (31 - __clz(below)) : 0;
}

}  // namespace fcn
