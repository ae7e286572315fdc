#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "pme or mesh or reproducible or langevin" > gpurun_out/pytest_gpu_pme.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_pme.log
timeout 400 python -u bench.py --config water_23k --steps 1500 --warmup 50 > gpurun_out/bench_water_23k_s6.json 2> gpurun_out/bench_water_23k_s6.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_water_23k_s6.json")); print("23k", d["value"], d["ns_per_day"], "ns/day", d["ms_per_step"], "ms | e2e", d["e2e"]["ns_per_day"], d["e2e"]["ms_per_step"], d["phases_ms_per_step"])
except Exception as e: print('no json', e)
PY
MDK_OPTS=pme_cufft=1 timeout 300 python -u bench.py --config water_23k --steps 1500 --warmup 50 --skip-extras 2>/dev/null | tail -1
