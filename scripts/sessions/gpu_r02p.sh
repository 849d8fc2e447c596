#!/bin/bash
# final checks after the wide-tile change: mega tests, default-path benches where "auto" picks wide tiles
line() { python -c "
import sys,json
d=json.loads([l for l in open('$1') if l.startswith('{')][-1])
print('$2', round(d['value']), round(d['e2e']['value']), d['kernel_ms']['fcn_mega'], d['fcn_mega']['ctas'], d['roofline']['kernel'], round(d['roofline']['frac'],3))"; }
timeout 600 python -m pytest tests/test_gpu_mega.py tests/test_gpu_bench_config.py -q -x 2>&1 | tail -2
timeout 300 python bench.py --workload people --steps 20 --warmup 5 > gpurun_out/bench_r02d_people.json 2>/dev/null; line gpurun_out/bench_r02d_people.json people_k20
timeout 300 python bench.py --workload people --no-cpu-baseline > gpurun_out/bench_r02d_people_k200.json 2>/dev/null; line gpurun_out/bench_r02d_people_k200.json people_k200
timeout 300 python bench.py --workload people --points 512 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02d_people512.json 2>/dev/null; line gpurun_out/bench_r02d_people512.json people512_k20
for b in 128 512; do timeout 300 python bench.py --batch $b --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02d_car_b$b.json 2>/dev/null; line gpurun_out/bench_r02d_car_b$b.json car_b$b; done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02d_car_k20.json 2>/dev/null; line gpurun_out/bench_r02d_car_k20.json car_k20
