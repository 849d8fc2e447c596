#!/bin/bash
# K=20 regions: stream count vs step imbalance
line() { python -c "
import sys,json
d=json.loads([l for l in open('$1') if l.startswith('{')][-1])
print('$2', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4))"; }
for st in 4 5 8 10 20; do timeout 300 python bench.py --steps 20 --warmup 5 --streams $st --no-cpu-baseline > gpurun_out/bench_r02s_$st.json 2>/dev/null; line gpurun_out/bench_r02s_$st.json k20_streams$st; done
for st in 10; do timeout 300 python bench.py --streams $st --no-cpu-baseline > gpurun_out/bench_r02s_k200_$st.json 2>/dev/null; line gpurun_out/bench_r02s_k200_$st.json k200_streams$st; done
for k in 16 24 40; do timeout 300 python bench.py --steps $k --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02s_k$k.json 2>/dev/null; line gpurun_out/bench_r02s_k$k.json k${k}_streams8; done
