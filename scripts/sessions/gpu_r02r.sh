#!/bin/bash
TAG=${1:-r02r}
mkdir -p gpurun_out
for cfgm in "sunrgbd:0:0" "sunrgbd:1:4" "sunrgbd:1:6" "sunrgbd:1:8" "sunrgbd:1:12" "people:0:0" "people:1:32" "people:1:48" "people:1:72"; do
    IFS=: read wl mega grid <<< "$cfgm"
    FCN_MEGA=$mega FCN_MEGA_GRID=$grid timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline \
        > gpurun_out/bench_${TAG}_${wl}_m${mega}_g${grid}.json 2> gpurun_out/bench_${TAG}_${wl}_m${mega}_g${grid}.err
    python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_${TAG}_${wl}_m${mega}_g${grid}.json") if l.startswith("{")][-1])
    print("${wl} mega=${mega} grid=${grid} value %.0f e2e %.0f lat %.3f ms fcn %s" % (d["value"], d["e2e"]["value"], d["latency"]["median_ms"], d["kernel_ms"].get("fcn_mega")))
except Exception as e:
    print("${wl} mega=${mega} grid=${grid} failed", e)
PY
done
