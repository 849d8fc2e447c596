"""Timeline of the persistent FCN kernel: CTA 0 dumps clock64 stamps per job (producer / MMA / epilogue)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FCN_MEGA"] = "1"
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 16
os.environ["FCN_MEGA_GRID"] = str(grid)
from frustum_convnet_b200 import config, synth  # noqa: E402
from frustum_convnet_b200.det_base import PointNetDet  # noqa: E402

cfg, w = config.load_workload("car")
sd = synth.make_state_dict(w["arch"], 3, "KITTI", seed=7)
m = PointNetDet(3, num_vec=3)
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
m.precision, m.use_cuda_graph = 1, False
m = m.cuda().eval()
data = {k: torch.from_numpy(v).cuda() for k, v in synth.make_frustums("car", 32, seed=1234).items()}
for _ in range(3):
    m(data)
torch.cuda.synchronize()
plan = list(m.engine()._plans.values())[0]
dbg = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
plan.mega_args.dbg_clocks = dbg.data_ptr()
m(data)
torch.cuda.synchronize()
plan.mega_args.dbg_clocks = None
d = dbg.cpu().numpy().reshape(64, 16)
descs = plan.mega_descs()
from frustum_convnet_b200 import mega  # noqa: E402
_, _, jobs, _ = mega.build_tables(descs)
t0 = d[0, 0]
print("grid", grid, "cols: job layer NS | fetch depsdone fence loads_issued | mma_wait_acc mma_first_full mma_end | epi_wait epi_accfull epi_tmemfree epi_flag   (clk rel. to first fetch)")
for q in range(64):
    if d[q, 0] == 0:
        break
    j = int(d[q, 1])
    r = lambda i: int(d[q, i] - t0) if d[q, i] else -1
    print("%3d j%4d %-13s NS%3d | %8d %8d %8d %8d | %8d %8d %8d %8d | %8d %8d %8d %8d" % (
        q, j, descs[jobs[j]["layer"]].name, int(d[q, 8]), r(0), r(2), r(3), r(4), r(5), r(6), r(9), r(7), r(10), r(11), r(12), r(13)))
