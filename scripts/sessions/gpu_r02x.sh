#!/bin/bash
# 2-CTA PointNet kernel, fine-grained slices: parity, bench, timeline
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_bench_config.py tests/test_gpu_mega.py -q -x 2>&1 | tail -4
line() { python -c "
import sys,json
d=json.loads([l for l in open('$1') if l.startswith('{')][-1])
print('$2', round(d['value']), round(d['e2e']['value']), d['kernel_ms'], round(d['roofline']['frac'],3))"; }
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02x_k200.json 2> gpurun_out/bench_r02x_k200.err; line gpurun_out/bench_r02x_k200.json new_k200
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02x_k20.json 2>/dev/null; line gpurun_out/bench_r02x_k20.json new_k20
timeout 200 python scripts/dbg_pointnet_clocks.py 3 > gpurun_out/dbg_s4_x.txt 2>&1; tail -8 gpurun_out/dbg_s4_x.txt
