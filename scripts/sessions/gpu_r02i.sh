#!/bin/bash
# 2-GPU session: peer-store exchange test, bench peer vs nccl
TAG=${1:-r02i}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py -q -x -s 2>&1 | tail -8
for ex in peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29511 + RANDOM % 200)) \
      bench.py --gpus 2 --steps 20 --warmup 5 --exchange $ex > gpurun_out/bench_${TAG}_2gpu_$ex.json 2> gpurun_out/bench_${TAG}_2gpu_$ex.err; echo "rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_2gpu_$ex.json"))
    print("$ex 2gpu value %.0f e2e %.0f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["config"]["collective"][:60])
except Exception as e:
    print("$ex failed", e)
PY
  tail -3 gpurun_out/bench_${TAG}_2gpu_$ex.err
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err
python -c "
import json
d=json.load(open('gpurun_out/bench_${TAG}_1gpu.json'))
print('1gpu value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['kernel_ms'], d['timing'])
"
