#!/bin/bash
# 256-wide N tiles in the persistent FCN kernel: parity, bench A/B, grid sweep
line() { python -c "
import sys,json
d=json.loads([l for l in open('$1') if l.startswith('{')][-1])
print('$2', round(d['value']), round(d['e2e']['value']), d['kernel_ms']['fcn_mega'], d['fcn_mega'])"; }
timeout 600 python -m pytest tests/test_gpu_mega.py tests/test_gpu_parity.py tests/test_gpu_bench_config.py -q -x 2>&1 | tail -2
FCN_MEGA_NT256=1 timeout 600 python -m pytest tests/test_gpu_mega.py tests/test_gpu_parity.py tests/test_gpu_bench_config.py -q -x 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02n_base.json 2>/dev/null; line gpurun_out/bench_r02n_base.json nt128
for g in 0 16 24; do FCN_MEGA_NT256=1 FCN_MEGA_GRID=$g timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02n_nt256_g$g.json 2>gpurun_out/bench_r02n_nt256_g$g.err; line gpurun_out/bench_r02n_nt256_g$g.json nt256_grid$g; done
timeout 300 python bench.py --workload people --no-cpu-baseline > gpurun_out/bench_r02n_people_base.json 2>/dev/null; line gpurun_out/bench_r02n_people_base.json people_nt128
FCN_MEGA_NT256=1 timeout 300 python bench.py --workload people --no-cpu-baseline > gpurun_out/bench_r02n_people_nt256.json 2>/dev/null; line gpurun_out/bench_r02n_people_nt256.json people_nt256
