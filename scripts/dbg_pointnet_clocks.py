"""Dump pipeline timestamps of the tensor-core PointNet kernel (CTA 0) for one scale."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from frustum_convnet_b200 import config, synth, _lib
from frustum_convnet_b200.det_base import PointNetDet

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg, w = config.load_workload("car")
sd = synth.make_state_dict(w["arch"], 3, "KITTI", seed=7)
m = PointNetDet(3, num_vec=3)
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
m.precision = 1
m = m.cuda().eval()
data = synth.make_frustums("car", 32, seed=1)
d = {k: torch.from_numpy(v).cuda() for k, v in data.items()}
for _ in range(3):
    m(d)
torch.cuda.synchronize()
plan = m.engine().plan(32, 1024, [280, 140, 70, 35])
a = plan.pn_args[scale]
buf = torch.zeros(8192, dtype=torch.int64, device="cuda")
a.dbg_clocks = buf.data_ptr()
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _lib.call("fcn_pointnet_tiles", C.byref(a), st)
torch.cuda.synchronize()
t = buf.cpu().numpy()
a.dbg_clocks = None
jobs = t[:4000].reshape(-1, 4)
jobs = jobs[jobs[:, 0] > 0]
l1 = t[6144:6144+32*8].reshape(-1, 32)
tiles = t[4096:6144].reshape(-1, 16)
tiles = tiles[tiles[:, 0] > 0]
base = min(jobs[0, 0], tiles[0, 0])
print("scale", scale + 1, "jobs", len(jobs), "tiles(CTA0)", len(tiles))
print("precision", a.precision)
print("MMA warp per job: start | wait_a  wait_w  issue")
for j, r in enumerate(jobs[:56]):
    print("%3d %8d | %6d %6d %6d" % (j, r[0] - base, r[1] - r[0], r[2] - r[1], r[3] - r[2]))
nj = 24 if a.precision == 2 and scale == 3 else 0
for it in range(len(jobs) // nj if nj else 0):
    blk = jobs[it * nj:(it + 1) * nj]
    seg = lambda lo, hi: (blk[hi - 1, 3] - blk[lo, 0], int((blk[lo:hi, 1] - blk[lo:hi, 0]).sum()), int((blk[lo:hi, 2] - blk[lo:hi, 1]).sum()), int((blk[lo:hi, 3] - blk[lo:hi, 2]).sum()))
    print("pair %d: start %d span %d | L2 span/wait_a/wait_w/issue %s | c0 %s | c1 %s | gap to next %d" % (
        it, blk[0, 0] - base, blk[-1, 3] - blk[0, 0], seg(0, 8), seg(8, 16), seg(16, 24),
        (jobs[(it + 1) * nj, 0] - blk[-1, 3]) if (it + 1) * nj < len(jobs) else -1))
print("compute warp 0 per tile: start recs_ready | L1_done->acc2wait acc2_ready epi2_done | chunk: acc3_ready epi3_done ...")
for r in tiles:
    print(" ".join("%8d" % (x - base if x > 0 else -1) for x in r[:16]))
print("layer-1 stamps (warp 0): start | per K block: arrive, next a_free passed")
for r in l1:
    if r[0] > 0:
        print(" ".join("%7d" % (x - base if x > 0 else -1) for x in r[:17]))
