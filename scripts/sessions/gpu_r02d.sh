#!/bin/bash
# mega grid / streams sweep
TAG=${1:-r02d}
mkdir -p gpurun_out
for cfgm in "1:8:8" "1:12:8" "1:16:8" "1:24:8" "1:16:12" "1:16:16" "1:24:12" "0:0:12"; do
    IFS=: read mega grid streams <<< "$cfgm"
    FCN_MEGA=$mega FCN_MEGA_GRID=$grid timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --streams $streams \
        > gpurun_out/bench_${TAG}_m${mega}_g${grid}_s${streams}.json 2> gpurun_out/bench_${TAG}_m${mega}_g${grid}_s${streams}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_m${mega}_g${grid}_s${streams}.json"))
    print("mega=${mega} grid=${grid} streams=${streams} value %.0f e2e %.0f lat %.3f ms fcn %s" % (d["value"], d["e2e"]["value"], d["latency"]["median_ms"], d["kernel_ms"].get("fcn_mega")))
except Exception as e:
    print("mega=${mega} grid=${grid} failed", e)
PY
    tail -2 gpurun_out/bench_${TAG}_m${mega}_g${grid}_s${streams}.err
done
