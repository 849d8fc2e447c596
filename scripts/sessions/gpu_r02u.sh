#!/bin/bash
# tail-split PointNet kernels: parity, then bench + per-kernel times; then the queued round-1 variants (old libs)
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_bench_config.py tests/test_gpu_mega.py -q -x 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02u_k200.json 2> gpurun_out/bench_r02u_k200.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02u_k20.json 2> gpurun_out/bench_r02u_k20.err
python -c "
import json
for n in ('k200','k20'):
    d=json.loads([l for l in open('gpurun_out/bench_r02u_%s.json'%n) if l.startswith('{')][-1])
    print(n, round(d['value']), round(d['e2e']['value']), d['kernel_ms'], round(d['roofline']['frac'],3))
"
