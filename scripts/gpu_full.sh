#!/bin/bash
# Full GPU validation: whole suite, smoke, default bench (driver flags), per-workload lines
TAG=${1:-r02full}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1
tail -6 gpurun_out/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_car_k20.json 2> gpurun_out/bench_${TAG}_car_k20.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}_car_k200.json 2> gpurun_out/bench_${TAG}_car_k200.err
for wl in people sunrgbd; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_${wl}.json 2> gpurun_out/bench_${TAG}_${wl}.err
done
timeout 600 python bench.py --workload people --points 512 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_people512.json 2> gpurun_out/bench_${TAG}_people512.err
timeout 600 python bench.py --precision 0 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_car_fp32.json 2> gpurun_out/bench_${TAG}_car_fp32.err
for b in 128 512; do
  timeout 600 python bench.py --batch $b --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_car_b$b.json 2> gpurun_out/bench_${TAG}_car_b$b.err
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*.json")):
    try:
        txt = open(f).read()
        d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.0f e2e %.0f ms %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]),
              "roof", d.get("roofline", {}).get("kernel"), round(d.get("roofline", {}).get("frac", 0), 3),
              "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
