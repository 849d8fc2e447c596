import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from frustum_convnet_b200 import config, synth, _lib
from frustum_convnet_b200.det_base import PointNetDet
cfg, w = config.load_workload("car")
sd = synth.make_state_dict(w["arch"], 3, "KITTI", seed=7)
m = PointNetDet(3, num_vec=3)
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
m.precision = 1
m = m.cuda().eval()
data = synth.make_frustums("car", 32, seed=1)
d = {k: torch.from_numpy(v).cuda() for k, v in data.items()}
m(d); torch.cuda.synchronize()
plan = m.engine().plan(32, 1024, [280, 140, 70, 35])
st = torch.cuda.current_stream().cuda_stream
for s in [int(x) for x in sys.argv[1:]] or [0, 1, 2, 3]:
    a = plan.pn_args[s]
    try:
        nt = int(plan.ntiles[s].item())
        for i in range(3000):
            if i % 5 == 0:   # shuffle the tile table like the atomicAdd order of group_emit does
                perm = torch.randperm(nt, device="cuda")
                plan.tiles[s][:nt] = plan.tiles[s][:nt][perm]
            _lib.call("fcn_pointnet_tiles", C.byref(a), st)
            if i % 500 == 499:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        print("scale", s + 1, "ok", flush=True)
    except Exception as e:
        print("scale", s + 1, "FAILED", str(e)[:80], flush=True)
        break

# timing of each scale kernel (hot, back to back)
for s in range(4):
    a = plan.pn_args[s]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200):
        _lib.call("fcn_pointnet_tiles", C.byref(a), st)
    e1.record(); torch.cuda.synchronize()
    print("scale", s + 1, "us per launch", e0.elapsed_time(e1) * 1000 / 200)
