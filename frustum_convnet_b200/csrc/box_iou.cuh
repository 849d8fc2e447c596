// Pairwise rotated-box IoU of the train-branch metrics: replaces the CPU Boost.Geometry
// call `rbbox_iou_3d_pair` (ops/pybind11/box_ops.h:173-260, called at models/det_base.py:495) that forces a
// device->host copy of every predicted box per training step.
//
// The arithmetic lives in this header as plain C++ so that tests/test_box_iou_cpu.py can compile the SAME
// functions for the host (g++) and check them against oracle/box_iou.py without a GPU; the kernels in
// box_iou.cu only add the thread mapping.  No part of the product calls the host build.
#pragma once

#ifdef __CUDACC__
#define FCN_HD __host__ __device__ __forceinline__
#else
#define FCN_HD inline
#endif

namespace fcn {

struct Pt2 {
    float x, y;
};

// signed area, counter-clockwise positive
FCN_HD float shoelace_ccw(const Pt2 *p, int n) {
    float a = 0.f;
    for (int i = 0; i < n; ++i) {
        const Pt2 u = p[i], v = p[i + 1 == n ? 0 : i + 1];
        a += u.x * v.y - v.x * u.y;
    }
    return 0.5f * a;
}

// Sutherland-Hodgman: clip the convex quadrilateral `subj` against the convex quadrilateral `clip`
// (either orientation).  Returns the vertex count (<= 8) of the intersection written to `out`.
FCN_HD int clip_quads(const Pt2 *subj, const Pt2 *clip, Pt2 *out) {
    Pt2 buf[2][8];
    int n = 4, cur = 0;
    for (int i = 0; i < 4; ++i) buf[0][i] = subj[i];
    const float sgn = shoelace_ccw(clip, 4) >= 0.f ? 1.f : -1.f;
    for (int e = 0; e < 4 && n > 0; ++e) {
        const Pt2 a = clip[e], b = clip[(e + 1) & 3];
        const float ex = b.x - a.x, ey = b.y - a.y;
        const Pt2 *in = buf[cur];
        Pt2 *o = buf[cur ^ 1];
        int m = 0;
        for (int j = 0; j < n; ++j) {
            const Pt2 p = in[j], q = in[j + 1 == n ? 0 : j + 1];
            const float sp = sgn * (ex * (p.y - a.y) - ey * (p.x - a.x));
            const float sq = sgn * (ex * (q.y - a.y) - ey * (q.x - a.x));
            if (sp >= 0.f && m < 8) o[m++] = p;
            if ((sp >= 0.f) != (sq >= 0.f) && m < 8) {
                const float t = sp / (sp - sq);
                o[m].x = p.x + t * (q.x - p.x);
                o[m].y = p.y + t * (q.y - p.y);
                ++m;
            }
        }
        n = m;
        cur ^= 1;
    }
    for (int i = 0; i < n; ++i) out[i] = buf[cur][i];
    return n;
}

// c, q: the 8 corners (x, y, z) of one box pair in the reference's corner order (model_util.py:48-72).
// iou[0] = bird's-eye-view IoU, iou[1] = 3-D IoU; (0, 0) when the BEV polygons do not overlap.
FCN_HD void rbbox_iou_pair(const float *c, const float *q, float *iou) {
    const int order[4] = {6, 7, 4, 5};                       // box_ops.h:206-224
    Pt2 P[4], Q[4], I[8];
    // coordinates relative to the first polygon vertex: areas are translation invariant and the fp32 cross
    // products then work on box-sized (metres) instead of scene-sized (tens of metres) magnitudes
    const float ox = c[order[0] * 3 + 0], oz = c[order[0] * 3 + 2];
    for (int i = 0; i < 4; ++i) {
        P[i].x = c[order[i] * 3 + 0] - ox;
        P[i].y = c[order[i] * 3 + 2] - oz;
        Q[i].x = q[order[i] * 3 + 0] - ox;
        Q[i].y = q[order[i] * 3 + 2] - oz;
    }
    iou[0] = 0.f;
    iou[1] = 0.f;
    // Degenerate boxes.  The reference hands Boost un-`correct`ed rings (box_ops.h:206-224); a ring whose signed
    // area is not positive in Boost's clockwise convention (a decoded box with exactly one negative footprint
    // size, e.g. from untrained weights) violates the polygon concept and bg::intersection / bg::union_ are
    // unspecified on it.  Decision (pinned in oracle/box_iou.py and tests/test_box_iou_cpu.py): such a pair
    // scores (0, 0) - an invalid box gets no overlap credit - which keeps 0 <= IoU_3D <= IoU_2D <= 1.  Negative
    // HEIGHTS need no special case: `vol = max(0, area*h)` and `max(0, ymax-ymin)` (box_ops.h:242-245) already
    // yield a zero 3-D IoU, exactly as the reference arithmetic does.
    const float area = -shoelace_ccw(P, 4), qarea = -shoelace_ccw(Q, 4);   // clockwise positive (Boost default)
    if (!(area > 0.f) || !(qarea > 0.f)) return;
    const int n = clip_quads(P, Q, I);
    if (n < 3) return;
    float inter = shoelace_ccw(I, n);
    inter = inter < 0.f ? -inter : inter;
    if (!(inter > 0.f)) return;
    const float uni = area + qarea - inter;
    const float ymax = c[1] < q[1] ? c[1] : q[1];                          // min of the corner-0 heights
    const float ymin = c[13] > q[13] ? c[13] : q[13];                      // max of the corner-4 heights
    const float h = c[1] - c[13], qh = q[1] - q[13];
    const float dy = ymax - ymin;
    const float inter_vol = inter * (dy > 0.f ? dy : 0.f);
    const float vol = area * h > 0.f ? area * h : 0.f, qvol = qarea * qh > 0.f ? qarea * qh : 0.f;
    iou[0] = inter / uni;
    iou[1] = inter_vol / (vol + qvol - inter_vol);
}

}  // namespace fcn
