"""Rotated 3-D NMS row on the CPU: known answers for the oracle restatement of
rotate_nms_3d_cc (ops/pybind11/rbbox_iou.py:294-311, nms_cpu.h:148-240)."""
import numpy as np

from oracle import nms as onms


def _det(cx, cz, l, w, h, ry, score, cy=1.0):
    return [cx, cy, cz, l, w, h, ry, score]


def test_known_answers():
    # three identical boxes: only the best-scored survives; a far box is untouched
    d = np.array([_det(0, 10, 4, 2, 1.5, 0.1, 0.5), _det(0, 10, 4, 2, 1.5, 0.1, 0.9), _det(0, 10, 4, 2, 1.5, 0.1, 0.7),
                  _det(30, 40, 4, 2, 1.5, 0.3, 0.2)])
    assert onms.rotate_nms_3d_cc(d, 0.5) == [1, 3]
    # shifted by half a length: 3-D IoU = 1/3 -> kept at thresh 0.5, suppressed at thresh 0.3
    d = np.array([_det(0, 10, 4, 2, 1.5, 0.0, 0.9), _det(2, 10, 4, 2, 1.5, 0.0, 0.8)])
    assert onms.rotate_nms_3d_cc(d, 0.5) == [0, 1]
    assert onms.rotate_nms_3d_cc(d, 0.3) == [0]
    # no height overlap: BEV identical but 3-D IoU 0
    d = np.array([_det(0, 10, 4, 2, 1.5, 0.0, 0.9, cy=0.0), _det(0, 10, 4, 2, 1.5, 0.0, 0.8, cy=5.0)])
    assert onms.rotate_nms_3d_cc(d, 0.1) == [0, 1]
    # chain: A suppresses B, B would suppress C but is gone -> C survives (greedy semantics)
    d = np.array([_det(0.0, 10, 4, 2, 1.5, 0.0, 0.9), _det(1.2, 10, 4, 2, 1.5, 0.0, 0.8), _det(2.4, 10, 4, 2, 1.5, 0.0, 0.7)])
    assert onms.rotate_nms_3d_cc(d, 0.5) == [0, 2]
    # top_k, empty, ties (larger index first among equal scores)
    assert onms.rotate_nms_3d_cc(d, 0.5, top_k=1) == [0]
    assert onms.rotate_nms_3d_cc(np.zeros((0, 8)), 0.5) == []
    d = np.array([_det(0, 10, 4, 2, 1.5, 0.0, 0.5), _det(50, 10, 4, 2, 1.5, 0.0, 0.5)])
    assert onms.rotate_nms_3d_cc(d, 0.5) == [1, 0]


def test_order_and_subset_properties():
    rng = np.random.default_rng(3)
    for _ in range(5):
        n = 40
        d = np.concatenate([rng.normal([0, 1, 20], [3, 0.2, 3], (n, 3)), rng.uniform([3, 1.4, 1.3], [4.5, 1.9, 1.8], (n, 3)),
                            rng.uniform(-np.pi, np.pi, (n, 1)), rng.random((n, 1))], 1)
        keep = onms.rotate_nms_3d_cc(d, 0.25)
        assert len(set(keep)) == len(keep) and keep[0] == int(np.argmax(d[:, 7]))
        assert all(d[keep[i], 7] >= d[keep[i + 1], 7] for i in range(len(keep) - 1))
        # no two kept boxes overlap above the threshold
        c = onms.boxes3d2corners(d)
        from oracle import box_iou as ob
        for a in range(len(keep)):
            for b in range(a + 1, len(keep)):
                assert ob.rbbox_iou_3d_pair(c[keep[a]:keep[a] + 1], c[keep[b]:keep[b] + 1])[0, 1] < 0.25
