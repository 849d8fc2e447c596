"""TEST INFRASTRUCTURE - generates tests/golden/input_builder.npz by running the REFERENCE'S OWN
``datasets.provider_sample.ProviderDataset.__getitem__`` (cfgs/det_sample.yaml: RTC, 1024 points, from_rgb_detection
inputs) on synthetic frustums.  The instance is created without its pickle-loading constructor and fed the lists it
would have read; np.random is seeded per item so that the committed ``choice`` is exactly what the reference drew.
Run in the authoring container only (needs /root/reference):  python -m oracle.make_golden_inputs"""
import os
import sys

import numpy as np

from . import ref_import

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ref_import._install_shims()
    for k in [k for k in sys.modules if k.startswith(("configs", "models", "datasets"))]:
        del sys.modules[k]
    from configs.config import cfg, merge_cfg_from_file  # type: ignore
    merge_cfg_from_file(os.path.join(ref_import.REF, "cfgs", "det_sample.yaml"))
    from datasets import provider_sample as ps  # type: ignore
    from datasets.dataset_info import DATASET_INFO  # type: ignore
    rng = np.random.default_rng(2024)
    B, N = 6, cfg.DATA.NUM_SAMPLES
    counts = [37, 1024, 2500, 800, 1500, 5]
    ds = object.__new__(ps.ProviderDataset)
    ds.npoints, ds.one_hot, ds.from_rgb_detection = N, True, True
    ds.category_info = DATASET_INFO[cfg.DATA.DATASET_NAME]
    classes = ["Car", "Pedestrian", "Cyclist", "Car", "Car", "Cyclist"]
    ds.type_list = classes
    pts = [np.stack([rng.uniform(-10, 10, n), rng.uniform(-1.5, 1.5, n), rng.uniform(2, 60, n)], 1).astype(np.float32)
           for n in counts]
    ds.input_list = pts
    ds.frustum_angle_list = [float(a) for a in rng.uniform(-2.2, -0.9, B)]
    ds.box2d_list = [np.array([x, y, x + w, y + h]) for x, y, w, h in
                     zip(rng.uniform(0, 1000, B), rng.uniform(100, 250, B), rng.uniform(20, 200, B), rng.uniform(20, 150, B))]
    Pm = np.array([[721.5377, 0.0, 609.5593, 44.85728], [0.0, 721.5377, 172.854, 0.2163791], [0.0, 0.0, 1.0, 0.002745884]])
    ds.calib_list = [{"P2": (Pm + rng.normal(0, 0.5, (3, 4)) * (Pm != 0)).reshape(-1)} for _ in range(B)]
    ds.prob_list = [0.9] * B
    outs, choices = [], []
    for b in range(B):
        np.random.seed(1000 + b)
        outs.append(ds[b])
        np.random.seed(1000 + b)
        choices.append(np.random.choice(counts[b], N, counts[b] < N))
    g = {
        "points": np.concatenate(pts), "offsets": np.concatenate([[0], np.cumsum(counts)]).astype(np.int32),
        "choice": np.stack(choices).astype(np.int32), "frustum_angle": np.asarray(ds.frustum_angle_list, dtype=np.float64),
        "box2d": np.stack(ds.box2d_list).astype(np.float64),
        "P": np.stack([c["P2"].reshape(3, 4) for c in ds.calib_list]).astype(np.float64),
        "cls_index": np.asarray([ds.category_info.CLASSES.index(c) for c in classes], dtype=np.int32),
        "strides": np.asarray(cfg.DATA.STRIDE, dtype=np.float64), "max_depth": np.float64(cfg.DATA.MAX_DEPTH),
    }
    for k in ("point_cloud", "center_ref1", "center_ref2", "center_ref3", "center_ref4", "rot_angle", "one_hot"):
        g["ref_" + k] = np.stack([o[k].numpy() for o in outs])
    path = os.path.join(ROOT, "tests", "golden", "input_builder.npz")
    np.savez_compressed(path, **g)
    print(path, {k: v.shape for k, v in g.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
