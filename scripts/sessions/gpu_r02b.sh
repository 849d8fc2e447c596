#!/bin/bash
# Round-2 session B: persistent FCN kernel - parity vs the per-layer path, then bench A/B (FCN_MEGA 0/1, grid sizes).
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mega.py -q -s -x > gpurun_out/pytest_${TAG}_mega.log 2>&1
tail -12 gpurun_out/pytest_${TAG}_mega.log
if grep -q "passed" gpurun_out/pytest_${TAG}_mega.log && ! grep -q "failed" gpurun_out/pytest_${TAG}_mega.log; then
  for cfgm in "0:0" "1:0" "1:36" "1:72" "1:148"; do
    IFS=: read mega grid <<< "$cfgm"
    FCN_MEGA=$mega FCN_MEGA_GRID=$grid timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline \
        > gpurun_out/bench_${TAG}_m${mega}_g${grid}.json 2> gpurun_out/bench_${TAG}_m${mega}_g${grid}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_m${mega}_g${grid}.json"))
    print("mega=${mega} grid=${grid} value %.0f e2e %.0f lat %.3f ms" % (d["value"], d["e2e"]["value"], d["latency"]["median_ms"]), d["kernel_ms"])
except Exception as e:
    print("mega=${mega} grid=${grid} failed", e)
PY
    tail -2 gpurun_out/bench_${TAG}_m${mega}_g${grid}.err
  done
fi
timeout 900 python -m pytest tests/test_gpu_bench_config.py tests/test_gpu_tc.py tests/test_gpu_train_metrics.py -q -s > gpurun_out/pytest_${TAG}_tol.log 2>&1
tail -8 gpurun_out/pytest_${TAG}_tol.log
