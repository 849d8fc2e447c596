#!/bin/bash
# NT=256: streams x grid, and 20-step regions
line() { python -c "
import sys,json
d=json.loads([l for l in open('$1') if l.startswith('{')][-1])
print('$2', round(d['value']), round(d['e2e']['value']), d['kernel_ms']['fcn_mega'], d['fcn_mega']['ctas'])"; }
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline $EXTRA > gpurun_out/bench_r02o_$tag.json 2>/dev/null; line gpurun_out/bench_r02o_$tag.json $tag; }
EXTRA="--streams 14" run car_nt256_g12_s14 FCN_MEGA_NT256=1
EXTRA="--streams 14" run car_nt256_g16_s14 FCN_MEGA_NT256=1 FCN_MEGA_GRID=16
EXTRA="--streams 18" run car_nt256_g12_s18 FCN_MEGA_NT256=1
EXTRA="--streams 14" run car_nt128_s14 FCN_MEGA_NT256=0
EXTRA="--steps 20 --warmup 5" run car_k20_nt128 FCN_MEGA_NT256=0
EXTRA="--steps 20 --warmup 5" run car_k20_nt256_g16 FCN_MEGA_NT256=1 FCN_MEGA_GRID=16
EXTRA="--steps 20 --warmup 5" run car_k20_nt256_g24 FCN_MEGA_NT256=1 FCN_MEGA_GRID=24
EXTRA="--steps 20 --warmup 5 --workload people" run people_k20_nt128 FCN_MEGA_NT256=0
EXTRA="--steps 20 --warmup 5 --workload people" run people_k20_nt256 FCN_MEGA_NT256=1
EXTRA="--steps 20 --warmup 5 --workload people" run people_k20_nt256_g48 FCN_MEGA_NT256=1 FCN_MEGA_GRID=48
EXTRA="--steps 20 --warmup 5 --workload sunrgbd" run sun_k20_nt128 FCN_MEGA_NT256=0
EXTRA="--steps 20 --warmup 5 --workload sunrgbd" run sun_k20_nt256 FCN_MEGA_NT256=1
