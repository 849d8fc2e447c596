#!/bin/bash
TAG=r02
mkdir -p gpurun_out
NCU_BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pool-mb 2 --min-seconds 0 --max-regions 3 --streams 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:'group_.*_kernel|pointnet_.*_kernel|conv_gemm_.*_kernel|fcn_mega.*|decode_eval_kernel' -c 300 --csv \
    --log-file gpurun_out/launches_${TAG}_eval_mega.csv $NCU_BENCH > gpurun_out/bench_under_ncu_${TAG}_mega.log 2>&1
echo "ncu list exit $?"
bash scripts/gpu_ncu_one.sh ${TAG}_mega_final fcn_mega_kernel 3
for wl in people sunrgbd; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_final_${wl}.json 2> gpurun_out/bench_${TAG}_final_${wl}.err
done
timeout 600 python bench.py --workload people --points 512 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_final_people512.json 2> gpurun_out/bench_${TAG}_final_people512.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_final_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]), d["fcn_mega"], d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
