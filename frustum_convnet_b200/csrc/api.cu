// C-ABI dispatch for the PointNet tile kernels and the conv GEMMs (argument validation +
// precision routing).  See include/frustum_b200.h.
#include "common.cuh"

namespace fcn {
int pointnet_tiles_simt(const fcn_pointnet_args &a, cudaStream_t stream);
int pointnet_tiles_tc(const fcn_pointnet_args &a, cudaStream_t stream);
int pointnet_tiles_tc2(const fcn_pointnet_args &a, cudaStream_t stream);
int conv_gemm_simt(const fcn_conv_args &a, cudaStream_t stream);
int conv_gemm_tc(const fcn_conv_args &a, cudaStream_t stream);
int conv_gemm_tma(const fcn_conv_args &a, cudaStream_t stream);
}  // namespace fcn

using namespace fcn;

extern "C" int fcn_pointnet_tiles(const fcn_pointnet_args *args, fcn_stream_t stream) {
    FCN_REQUIRE(args != nullptr, "args is NULL");
    fcn_pointnet_args a = *args;
    if (a.feat_pitch == 0) a.feat_pitch = a.T;
    FCN_REQUIRE(a.feat_pitch >= a.T, "feat_pitch must be >= T");
    FCN_REQUIRE(a.B >= 0 && a.T >= 1 && a.K >= 1, "bad B/T/K");
    if (a.B == 0 || a.max_tiles == 0) return FCN_OK;
    FCN_REQUIRE(a.rows && a.tiles && a.ntiles && a.out, "NULL pointer");
    FCN_REQUIRE(a.w1t && a.b1 && a.b2 && a.b3, "NULL weight pointer");
    FCN_REQUIRE(a.row_cap >= a.T * a.K, "row_cap too small");
    FCN_REQUIRE(a.unpooled || a.ld_feat >= a.C3, "ld_feat too small");
    if (a.precision == 0) {
        FCN_REQUIRE(a.w2t && a.w3t, "NULL weight pointer");
        return pointnet_tiles_simt(a, (cudaStream_t)stream);
    }
    if (a.precision == 1) return pointnet_tiles_tc(a, (cudaStream_t)stream);
    if (a.precision == 2) return pointnet_tiles_tc2(a, (cudaStream_t)stream);
    return invalid(__func__, "precision must be 0 (fp32 SIMT), 1 (TF32 tcgen05) or 2 (TF32 tcgen05, 2-CTA clusters)");
}

extern "C" int fcn_conv_gemm(const fcn_conv_args *args, fcn_stream_t stream) {
    FCN_REQUIRE(args != nullptr, "args is NULL");
    fcn_conv_args a = *args;
    if (a.P_m == 0) a.P_m = a.T_out;
    if (a.P_store == 0) a.P_store = a.T_store;
    for (int s = 0; s < a.n_seg && s < FCN_MAX_SEGS; ++s)
        if (a.seg[s].pitch == 0) a.seg[s].pitch = a.seg[s].T_src;
    FCN_REQUIRE(a.P_m >= a.T_out && a.P_store >= a.T_store, "pitches must cover the valid rows");
    FCN_REQUIRE(a.B >= 0 && a.T_out >= 0, "negative size");
    FCN_REQUIRE(a.n_seg >= 1 && a.n_seg <= FCN_MAX_SEGS, "n_seg out of range");
    FCN_REQUIRE(a.K_pad > 0 && a.K_pad % 32 == 0, "K_pad must be a positive multiple of 32");
    FCN_REQUIRE(a.n_cols > 0 && a.n_cols % 64 == 0, "n_cols must be a positive multiple of 64");
    FCN_REQUIRE(a.Cout > 0 && a.Cout % 4 == 0 && a.up >= 1 && a.up * a.Cout <= a.n_cols, "bad Cout/up");
    FCN_REQUIRE(a.ld_out % 4 == 0 && a.c_off % 4 == 0 && a.c_off + a.Cout <= a.ld_out, "bad output slice");
    int k = 0;
    for (int s = 0; s < a.n_seg; ++s) {
        FCN_REQUIRE(a.seg[s].src != nullptr, "NULL segment source");
        FCN_REQUIRE(a.seg[s].ld % 4 == 0 && a.seg[s].C >= 1 && a.seg[s].C <= a.seg[s].ld, "bad segment ld/C");
        FCN_REQUIRE(a.seg[s].stride >= 1 && a.seg[s].T_src >= 1, "bad segment stride/T");
        FCN_REQUIRE(a.seg[s].pitch >= a.seg[s].T_src, "segment pitch must be >= T_src");
        k += ((a.seg[s].C + 31) / 32) * 32;
    }
    FCN_REQUIRE(k <= a.K_pad && a.K_pad - k < 64, "K_pad does not match the padded segments");
    if (a.B * a.T_out == 0) return FCN_OK;
    FCN_REQUIRE(a.wt && a.bias && a.out, "NULL pointer");
    if (a.precision == 0) return conv_gemm_simt(a, (cudaStream_t)stream);
    if (a.precision == 1 || a.precision == 2) return conv_gemm_tc(a, (cudaStream_t)stream);
    if (a.precision == 3 || a.precision == 4) return conv_gemm_tma(a, (cudaStream_t)stream);
    return invalid(__func__, "precision must be 0 (fp32 SIMT), 1|2 (TF32 cp.async gather) or 3|4 (TF32 TMA)");
}
