#!/bin/bash
# 8-GPU session: eval bench (peer stores), train bench (all-reduce), same flags as the driver
TAG=${1:-r02_8gpu}
N=${2:-8}
mkdir -p gpurun_out
P1=$((29700 + RANDOM % 100)); P2=$((29850 + RANDOM % 100)); P3=$((29400 + RANDOM % 100))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P1 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_eval.json 2> gpurun_out/bench_${TAG}_eval.err
[ "$3" = "skip-nccl" ] || timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P2 \
    bench.py --gpus $N --steps 20 --warmup 5 --exchange nccl > gpurun_out/bench_${TAG}_eval_nccl.json 2> gpurun_out/bench_${TAG}_eval_nccl.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P3 \
    bench.py --gpus $N --train --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_train.json 2> gpurun_out/bench_${TAG}_train.err
python - <<PY
import json
for n in ("eval", "eval_nccl", "train"):
    try:
        txt = open("gpurun_out/bench_${TAG}_%s.json" % n).read()
        d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print(n, "N=%d value %.0f e2e %.0f ms/step %.4f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]), d["config"].get("collective", "")[:50])
    except Exception as e:
        print(n, "failed", e)
PY
tail -qn 2 gpurun_out/bench_${TAG}_eval.err gpurun_out/bench_${TAG}_train.err
