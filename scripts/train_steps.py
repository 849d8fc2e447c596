"""A few training steps of refine_car (for ncu launch lists / profiling)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frustum_convnet_b200 import config, synth  # noqa: E402
from frustum_convnet_b200.det_base import PointNetDet  # noqa: E402
from frustum_convnet_b200.train_engine import TrainStep  # noqa: E402

cfg, w = config.load_workload("refine_car")
sd = synth.make_state_dict(w["arch"], 3, "KITTI", seed=7)
m = PointNetDet(3, num_vec=3)
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
m = m.cuda().train()
ts = TrainStep(m, lr=1e-3, weight_decay=1e-4)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
data = {k: torch.from_numpy(v).cuda() for k, v in synth.make_frustums("refine_car", B, seed=5, with_labels=True).items()}
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    ts.step(data)
torch.cuda.synchronize()
print("done", len(ts.engines))
