#!/bin/bash
# ncu --set full of ONE launch: usage gpu_ncu_one.sh <tag> <kernel-regex> <skip> [env assignments...]
TAG=$1; KRE=$2; SKIP=$3; shift 3
mkdir -p gpurun_out
env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c 1 \
    -f -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pool-mb 2 --min-seconds 0 --max-regions 3 \
    > gpurun_out/prof_$TAG.log 2>&1
echo "ncu $TAG exit $?"; ls -la gpurun_out/prof_$TAG.ncu-rep
