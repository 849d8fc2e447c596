#!/bin/bash
# grid / stream sweep of the persistent FCN kernel against the faster PointNet kernels
line() { python -c "
import sys,json
d=json.loads([l for l in open('$1') if l.startswith('{')][-1])
print('$2', round(d['value']), round(d['e2e']['value']), d['kernel_ms'], round(d['roofline']['frac'],3))"; }
timeout 300 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -1
for g in 16 20 24 32 40; do FCN_MEGA_GRID=$g timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02y_g$g.json 2>/dev/null; line gpurun_out/bench_r02y_g$g.json grid$g; done
for st in 6 12 16; do timeout 300 python bench.py --streams $st --no-cpu-baseline > gpurun_out/bench_r02y_s$st.json 2>/dev/null; line gpurun_out/bench_r02y_s$st.json streams$st; done
