#!/bin/bash
# tcgen05 bring-up: self-test first, then the TF32 PointNet tests, then a TF32 bench.
TAG=${1:-tc}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -s 2>&1 | tail -60 > gpurun_out/pytest_tc_$TAG.log
cat gpurun_out/pytest_tc_$TAG.log | tail -45
timeout 300 python bench.py --steps 100 --warmup 5 --precision 1 --no-cpu-baseline > gpurun_out/bench_tc_$TAG.json 2> gpurun_out/bench_tc_$TAG.err
tail -c 2500 gpurun_out/bench_tc_$TAG.json; tail -5 gpurun_out/bench_tc_$TAG.err
