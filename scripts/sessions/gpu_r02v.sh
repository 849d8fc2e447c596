#!/bin/bash
# region-swap 2-CTA PointNet kernel: parity, bench + per-kernel times; A/B against the previous kernel (variants/base)
# and against running scale 3 on the 2-CTA kernel too (FCN_PN_CLUSTER_MIN=128)
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_bench_config.py tests/test_gpu_mega.py -q -x 2>&1 | tail -4
line() { python -c "
import sys,json
d=json.loads([l for l in open('$1') if l.startswith('{')][-1])
print('$2', round(d['value']), round(d['e2e']['value']), d['kernel_ms'], round(d['roofline']['frac'],3))"; }
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02v_k200.json 2> gpurun_out/bench_r02v_k200.err; line gpurun_out/bench_r02v_k200.json new_k200
FCN_LIB_PATH=$PWD/frustum_convnet_b200/variants/libfrustum_b200_base.so timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02v_base.json 2>/dev/null; line gpurun_out/bench_r02v_base.json base_k200
echo "== scale 3 on the 2-CTA kernel"
FCN_PN_CLUSTER_MIN=128 timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_bench_config.py -q -x 2>&1 | tail -2
FCN_PN_CLUSTER_MIN=128 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02v_c128.json 2>/dev/null; line gpurun_out/bench_r02v_c128.json c128_k200
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02v_k200b.json 2>/dev/null; line gpurun_out/bench_r02v_k200b.json new_k200_again
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02v_k20.json 2>/dev/null; line gpurun_out/bench_r02v_k20.json new_k20
