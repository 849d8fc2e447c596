#!/bin/bash
# ncu --set full capture of one launch of a kernel: usage gpu_ncu_full.sh <tag> <kernel-regex> <skip> [bench args]
TAG=$1; KRE=$2; SKIP=$3; shift 3
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c 1 \
    -f -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pool-mb 2 "$@" \
    > gpurun_out/prof_$TAG.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_$TAG.ncu-rep
