#!/bin/bash
# 16-compute-warp 2-CTA PointNet kernel with layer 1 chasing the previous pair's last MMAs: parity, bench, timeline
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_bench_config.py tests/test_gpu_mega.py -q -x 2>&1 | tail -4
line() { python -c "
import sys,json
d=json.loads([l for l in open('$1') if l.startswith('{')][-1])
print('$2', round(d['value']), round(d['e2e']['value']), d['kernel_ms'], round(d['roofline']['frac'],3))"; }
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02w_k200.json 2> gpurun_out/bench_r02w_k200.err; line gpurun_out/bench_r02w_k200.json new_k200
FCN_LIB_PATH=$PWD/frustum_convnet_b200/variants/libfrustum_b200_base.so timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02w_base.json 2>/dev/null; line gpurun_out/bench_r02w_base.json base_k200
echo "== scale 3 on the 2-CTA kernel"
FCN_PN_CLUSTER_MIN=128 timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_bench_config.py -q -x 2>&1 | tail -2
FCN_PN_CLUSTER_MIN=128 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02w_c128.json 2>/dev/null; line gpurun_out/bench_r02w_c128.json c128_k200
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02w_k20.json 2>/dev/null; line gpurun_out/bench_r02w_k20.json new_k20
timeout 200 python scripts/dbg_pointnet_clocks.py 3 > gpurun_out/dbg_s4_w16.txt 2>&1; tail -8 gpurun_out/dbg_s4_w16.txt
