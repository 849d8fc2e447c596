"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's rotated 3-D NMS.

Follows ``rotate_nms_3d_cc`` (ops/pybind11/rbbox_iou.py:294-311): order = scores.argsort()[::-1]; corners by
``boxes3d2corners`` (:121-148); ``standup_iou`` = 3-D IoU of the axis-aligned bounding cubes (:62-96); then
``rotate_non_max_suppression_3d_cpu`` (ops/pybind11/nms_cpu.h:148-240): greedy in score order, box j is
suppressed by a kept box i when standup_iou(i, j) > 0 and the rotated 3-D IoU is >= thresh; ``keep[:top_k]``.

The rotated IoU itself is ``oracle.box_iou`` (PARITY UNPINNED for the Boost polygon clip, see its header; the
degenerate-ring decision documented there applies here too).  Ties in the score order: numpy's default
``argsort`` is not stable, so the reference leaves them unspecified; this restatement (and the kernel) use the
reverse of a STABLE ascending sort, i.e. among equal scores the larger index comes first."""
import numpy as np

from . import box_iou as ob


def boxes3d2corners(boxes):
    b = np.asarray(boxes, dtype=np.float64)
    return ob.box3d_corners(b[:, 0:3], b[:, 6], b[:, 3:6])      # sizes (l, w, h), same corner order


def rotate_nms_3d_cc(dets, thresh, top_k=300):
    dets = np.asarray(dets, dtype=np.float64)
    n = dets.shape[0]
    if n == 0:
        return []
    order = np.argsort(dets[:, 7], kind="stable")[::-1]
    corners = boxes3d2corners(dets)
    mn, mx = corners.min(1), corners.max(1)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    for a in range(n):
        i = order[a]
        if suppressed[i]:
            continue
        keep.append(int(i))
        for b in range(a + 1, n):
            j = order[b]
            if suppressed[j]:
                continue
            ext = np.minimum(mx[i], mx[j]) - np.maximum(mn[i], mn[j])
            if not (ext > 0).all():                     # standup IoU <= 0 (nms_cpu.h:192)
                continue
            iou3d = ob.rbbox_iou_3d_pair(corners[i:i + 1], corners[j:j + 1])[0, 1]
            if iou3d >= thresh:
                suppressed[j] = True
    return keep[:top_k]


def pair_margins(dets, thresh):
    """min |IoU_3D - thresh| over all pairs: tests use it to skip numerically ambiguous inputs."""
    dets = np.asarray(dets, dtype=np.float64)
    c = boxes3d2corners(dets)
    n = len(dets)
    m = np.inf
    for i in range(n):
        for j in range(i + 1, n):
            m = min(m, abs(ob.rbbox_iou_3d_pair(c[i:i + 1], c[j:j + 1])[0, 1] - thresh))
    return m
