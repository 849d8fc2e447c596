#!/bin/bash
# A/B: block section max in the 1-CTA PointNet kernel; then ncu evidence of the shipped kernels (launch list, s4, s3)
line() { python -c "
import sys,json
d=json.loads([l for l in open('$1') if l.startswith('{')][-1])
print('$2', round(d['value']), round(d['e2e']['value']), d['kernel_ms'], d['roofline']['kernel'], round(d['roofline']['frac'],3))"; }
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02z_k200.json 2>gpurun_out/bench_r02z_k200.err; line gpurun_out/bench_r02z_k200.json default
export V=$PWD/frustum_convnet_b200/variants/libfrustum_b200_smaxblk.so
FCN_LIB_PATH=$V timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -q -x 2>&1 | tail -1
FCN_LIB_PATH=$V timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02z_smaxblk.json 2>/dev/null; line gpurun_out/bench_r02z_smaxblk.json smaxblk
FCN_LIB_PATH=$V timeout 300 python bench.py --workload people --no-cpu-baseline > gpurun_out/bench_r02z_smaxblk_people.json 2>/dev/null; line gpurun_out/bench_r02z_smaxblk_people.json smaxblk_people
timeout 300 python bench.py --workload people --no-cpu-baseline > gpurun_out/bench_r02z_people.json 2>/dev/null; line gpurun_out/bench_r02z_people.json default_people
TAG=r02b
NCU_BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pool-mb 2 --min-seconds 0 --max-regions 3 --streams 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:'group_.*_kernel|pointnet_.*_kernel|conv_gemm_.*_kernel|fcn_mega.*|decode_eval_kernel' -c 300 --csv \
    --log-file gpurun_out/launches_${TAG}_eval.csv $NCU_BENCH > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
echo "ncu list exit $?"
bash scripts/gpu_ncu_one.sh ${TAG}_s4 pointnet_tc2_kernel 7
bash scripts/gpu_ncu_one.sh ${TAG}_s3 pointnet_tc2_kernel 6
