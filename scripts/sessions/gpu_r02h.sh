#!/bin/bash
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mega.py -q -x 2>&1 | tail -3
python scripts/dbg_mega_clocks.py 16 > gpurun_out/dbg_mega16_$TAG.txt 2>&1
for cfgm in "1:16:8" "1:20:8" "1:24:8" "1:32:8" "1:48:8" "1:20:12" "1:24:12" "1:32:12" "1:24:16"; do
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
done
