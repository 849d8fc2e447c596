#!/bin/bash
TAG=${1:-r02m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train.py tests/test_gpu_train_metrics.py -q -x 2>&1 | tail -6
timeout 600 python bench.py --train --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_train.json 2> gpurun_out/bench_${TAG}_train.err
tail -c 2500 gpurun_out/bench_${TAG}_train.json; tail -5 gpurun_out/bench_${TAG}_train.err
