#!/bin/bash
# evidence for DESIGN 5.1 / 5.2 / 5.3 on the final build: clock64 timeline of the 2-CTA PointNet kernel, ncu of the
# persistent FCN kernel with 256-wide tiles (people), single-GPU train line
timeout 200 python scripts/dbg_pointnet_clocks.py 3 > gpurun_out/r02e_clocks_pointnet_s4.txt 2>&1; tail -12 gpurun_out/r02e_clocks_pointnet_s4.txt
timeout 400 python bench.py --train --steps 20 --warmup 5 > gpurun_out/bench_r02e_train.json 2> gpurun_out/bench_r02e_train.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r02e_train.json') if l.startswith('{')][-1])
print('train', round(d['value']), round(d['e2e']['value']), d['phases_ms'], d['autograd_gpu_baseline']['value'])"
env timeout 600 ncu --set full --clock-control none --import-source on -k regex:fcn_mega_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_r02e_mega_people python bench.py --workload people --steps 2 --warmup 3 --no-cpu-baseline --pool-mb 2 --min-seconds 0 --max-regions 3 \
    > gpurun_out/prof_r02e_mega_people.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_r02e_mega_people.ncu-rep
