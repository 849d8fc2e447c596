#!/bin/bash
# A/B of kernel-variant libraries on the car bench:  scripts/gpu_ab.sh [bench args] -- name1 name2 ...
args=(); while [ "$1" != "--" ] && [ $# -gt 0 ]; do args+=("$1"); shift; done; shift
run() { timeout 200 python bench.py --warmup 10 --no-cpu-baseline "${args[@]}" 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],4), d['clocks'])"; }
run default
for n in "$@"; do FCN_LIB_PATH=$PWD/frustum_convnet_b200/variants/libfrustum_b200_$n.so run $n; done
run default
