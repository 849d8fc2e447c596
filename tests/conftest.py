import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# must mirror oracle/make_golden.py::CASES (name -> workload, B, data seed, weight seed, max_depth, N)
GOLDEN_CASES = {
    "car_full_b1": ("car", 1, 101, 7, None, None),
    "car_small_b3": ("car", 3, 102, 8, 17.5, None),
    "car_oddn_b2": ("car", 2, 103, 8, 17.5, 777),
    "people_small_b2": ("people", 2, 104, 9, 7.0, None),
    "sunrgbd_full_b2": ("sunrgbd", 2, 105, 10, None, None),
    "refine_car_b4": ("refine_car", 4, 106, 11, None, None),
}


def load_golden(name):
    """Returns (golden npz dict, regenerated inputs, numpy state dict, workload dict, cfg)."""
    from frustum_convnet_b200 import config, synth
    workload, B, dseed, wseed, md, N = GOLDEN_CASES[name]
    cfg, w = config.load_workload(workload)
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    data = synth.make_frustums(workload, B, seed=dseed, max_depth=md, N=N)
    chk = float(sum(np.asarray(v, dtype=np.float64).sum() for v in data.values()))
    assert abs(chk - float(g["input_checksum"])) <= 1e-6 * max(1.0, abs(chk)), \
        "synthetic generator drifted from the golden fixtures"
    sd = synth.make_state_dict(w["arch"], w["num_vec"], cfg.DATA.DATASET_NAME, seed=wseed)
    return g, data, sd, w, cfg


@pytest.fixture(scope="session")
def golden_loader():
    return load_golden
