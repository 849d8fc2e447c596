#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench, the ncu launch list.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag]'
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py --steps 100 --warmup 5 $BENCH_ARGS > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:'group_rows_kernel|pointnet_.*_kernel|conv_gemm_.*_kernel|decode_eval_kernel' -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --pool-mb 2 \
    > gpurun_out/bench_under_ncu_$TAG.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches_$TAG.csv
