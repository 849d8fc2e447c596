"""GPU: batched rotated 3-D NMS (csrc/nms3d.cu, ``fcn_rotate_nms_3d``) against the oracle (oracle/nms.py).
The keep lists are index sets, so parity is EXACT wherever no pair sits within 1e-4 of the IoU threshold (fp32
kernel vs float64 oracle); inputs with such a pair are regenerated."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dets(rng, n, spread):
    return np.concatenate([rng.normal([0, 1, 20], [spread, 0.2, spread], (n, 3)),
                           rng.uniform([3, 1.4, 1.3], [4.5, 1.9, 1.8], (n, 3)),
                           rng.uniform(-np.pi, np.pi, (n, 1)), rng.random((n, 1))], 1).astype(np.float32)


def test_single_list_matches_oracle_and_reference_contract():
    from frustum_convnet_b200.nms import cube_nms, rotate_nms_3d_cc
    from oracle import nms as onms
    rng = np.random.default_rng(1)
    done = 0
    while done < 12:
        n = int(rng.integers(2, 90))
        d = _dets(rng, n, spread=float(rng.uniform(1.0, 6.0)))
        thr = float(rng.choice([0.1, 0.25, 0.5, 0.7]))
        if onms.pair_margins(d, thr) < 1e-4:
            continue
        want = onms.rotate_nms_3d_cc(d, thr)
        got = rotate_nms_3d_cc(torch.from_numpy(d).cuda(), thr)
        assert got == want, (n, thr)
        done += 1
    assert cube_nms(torch.zeros((0, 8), device="cuda"), 0.5) == []
    one = torch.from_numpy(_dets(rng, 1, 1.0)).cuda()
    assert cube_nms(one, 0.5) == [0]
    with pytest.raises(RuntimeError):
        cube_nms(one.cpu(), 0.5)


def test_batched_segments_top_k_and_ties():
    from frustum_convnet_b200.nms import rotate_nms_3d_batched
    from oracle import nms as onms
    rng = np.random.default_rng(2)
    segs, lens = [], [0, 1, 37, 120, 300, 5]
    for n in lens:
        while True:
            d = _dets(rng, n, spread=3.0)
            if n < 2 or onms.pair_margins(d, 0.3) >= 1e-4:
                break
        segs.append(d)
    segs[5][:, 7] = 0.5                                   # all-equal scores: larger index first
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    dets = torch.from_numpy(np.concatenate(segs)).cuda()
    keep, cnt = rotate_nms_3d_batched(dets, torch.from_numpy(off), 0.3, top_k=40)
    keep, cnt = keep.cpu().numpy(), cnt.cpu().numpy()
    for s, d in enumerate(segs):
        want = [off[s] + k for k in onms.rotate_nms_3d_cc(d, 0.3, top_k=40)]
        assert cnt[s] == len(want) and keep[s, :cnt[s]].tolist() == want, s
        assert (keep[s, cnt[s]:] == -1).all()
