"""Dump per-K-block pipeline timestamps of the tensor-core conv kernel (CTA (0,0)) for one layer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from frustum_convnet_b200 import config, synth, _lib
from frustum_convnet_b200.det_base import PointNetDet

layer = sys.argv[1] if len(sys.argv) > 1 else "block4_conv2"
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
eng = m.engine()
plan = eng.plan(32, 1024, [280, 140, 70, 35])
idx = [L.name for L in eng.layers].index(layer)
a = plan.conv_args[idx]
KB = a.K_pad // 64
buf = torch.zeros(KB * 8, dtype=torch.int64, device="cuda")
a.dbg_clocks = buf.data_ptr()
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.call("fcn_conv_gemm", C.byref(a), st)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(KB, 8)
base = t[0, 0]
print("layer", layer, "KB", KB, "grid M tiles", (a.B * a.T_out + 127) // 128, "precision", a.precision)
print("kb  mma:t_start  +wait_a  +wait_w  +issue | prod: t_start +wait_empty +issue")
for kb in range(KB):
    r = t[kb]
    print("%3d %9d %7d %7d %7d | %9d %7d %7d" % (kb, r[0] - base, r[1] - r[0], r[2] - r[1], r[3] - r[2],
                                                   r[4] - base, r[5] - r[4], r[6] - r[5]))
a.dbg_clocks = None
