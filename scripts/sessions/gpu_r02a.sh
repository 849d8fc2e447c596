#!/bin/bash
# Round-2 session A: full GPU suite (no -x), smoke, two bench runs (default and the driver's --steps 20 --warmup 5)
# for the reproducibility check, the ncu launch list and --set full captures of the SHIPPED kernels.
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_$TAG.log 2>&1
tail -15 gpurun_out/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_k20.json 2> gpurun_out/bench_${TAG}_k20.err
tail -c 1500 gpurun_out/bench_${TAG}_k20.json; tail -3 gpurun_out/bench_${TAG}_k20.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}_k200.json 2> gpurun_out/bench_${TAG}_k200.err
tail -c 1500 gpurun_out/bench_${TAG}_k200.json; tail -3 gpurun_out/bench_${TAG}_k200.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_k20b.json 2> gpurun_out/bench_${TAG}_k20b.err
tail -c 600 gpurun_out/bench_${TAG}_k20b.json
NCU_BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pool-mb 2 --min-seconds 0 --max-regions 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:'group_.*_kernel|pointnet_.*_kernel|conv_gemm_.*_kernel|fcn_mega.*|decode_eval_kernel' -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv $NCU_BENCH > gpurun_out/bench_under_ncu_$TAG.log 2>&1
echo "ncu list exit $?"; wc -l gpurun_out/launches_$TAG.csv
for spec in "s4:pointnet_tc2_kernel<256:2" "s3:pointnet_tc_kernel<128:2" "convtma:conv_gemm_tma_kernel<128:20"; do
    IFS=: read name kre skip <<< "$spec"
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$kre" -s $skip -c 1 \
        -f -o gpurun_out/prof_${TAG}_$name $NCU_BENCH > gpurun_out/prof_${TAG}_$name.log 2>&1
    echo "ncu $name exit $?"; ls -la gpurun_out/prof_${TAG}_$name.ncu-rep
done
