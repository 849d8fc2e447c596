#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_peer.py tests/test_gpu_mega.py -q 2>&1 | tail -2
run() { tag=$1; shift; P=$((29300 + RANDOM % 500)); env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --warmup 5 $EXTRA > gpurun_out/bench_r02t2_$tag.json 2> gpurun_out/bench_r02t2_$tag.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r02t2_$tag.json') if l.startswith('{')][-1])
print('$tag', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4), d['kernel_ms']['fcn_mega'])"; }
EXTRA="--steps 20" run peer_k20 A=1
EXTRA="--steps 200" run peer_k200 A=1
