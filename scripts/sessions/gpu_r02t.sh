#!/bin/bash
# 2 ranks: replicas without any exchange vs peer stores (where does the N>1 per-step cost come from?)
run() { tag=$1; shift; P=$((29300 + RANDOM % 500)); env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 $EXTRA > gpurun_out/bench_r02t_$tag.json 2> gpurun_out/bench_r02t_$tag.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r02t_$tag.json') if l.startswith('{')][-1])
print('$tag', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4), d['host_issue_us_per_step'], d['kernel_ms']['fcn_mega'])"; }
EXTRA="--exchange nccl" run replicas_nocomm FCN_BENCH_NO_COMM=1
EXTRA="" run peer A=1
EXTRA="--steps 200" run peer_k200 A=1
EXTRA="--exchange nccl --steps 200" run replicas_nocomm_k200 FCN_BENCH_NO_COMM=1
