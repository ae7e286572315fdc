#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for CFG in water_23k; do
  timeout 400 python -u bench.py --config $CFG --steps 1500 --warmup 50 > gpurun_out/bench_${CFG}_s5.json 2> gpurun_out/bench_${CFG}_s5.err
  echo "$CFG rc=$?"; tail -c 400 gpurun_out/bench_${CFG}_s5.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${CFG}_s5.json")); print("$CFG", d["value"], d["ns_per_day"], "ns/day", d["ms_per_step"], "ms | e2e", d["e2e"]["ns_per_day"], d["e2e"]["ms_per_step"], "| roofline", round(d["roofline"]["frac"],4), d["phases_ms_per_step"], "rebuilds", d["config"]["nlist_rebuilds_in_timed"], 'launches/e2e step', d['e2e']['gpu_launches_per_step'])
except Exception as e: print('no json', e)
PY
done
MDK_OPTS=pme_cufft=1 timeout 300 python -u bench.py --config water_23k --steps 1500 --warmup 50 --skip-extras 2>/dev/null | tail -1
