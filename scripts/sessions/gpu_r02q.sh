#!/bin/bash
# 2 GPUs: train bench (all-reduce overlap) + eval bench peer, on the same box
TAG=${1:-r02q}
mkdir -p gpurun_out
PORT=$((29600 + RANDOM % 200))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus 2 --train --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_train_2gpu.json 2> gpurun_out/bench_${TAG}_train_2gpu.err
timeout 600 python bench.py --train --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_train_1gpu.json 2> gpurun_out/bench_${TAG}_train_1gpu.err
python - <<PY
import json
for n in ("train_2gpu", "train_1gpu"):
    try:
        txt = open("gpurun_out/bench_${TAG}_%s.json" % n).read()
        d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print(n, "value %.0f e2e %.0f ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("phases_ms"), d["autograd_gpu_baseline"]["value"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/bench_${TAG}_train_2gpu.err
