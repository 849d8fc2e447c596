#!/bin/bash
# ncu launch list (per-launch device time, cold-cache/serialised: compare SHARES) of our kernels.
TAG=${1:-x}; shift
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:'group_.*_kernel|pointnet_.*_kernel|conv_gemm_.*_kernel|decode_eval_kernel' -c 300 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --pool-mb 2 "$@" \
    > gpurun_out/bench_under_ncu_$TAG.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches_$TAG.csv
