#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 400 python -u bench.py --config water_23k --steps 1500 --warmup 50 > gpurun_out/bench_water_23k_s8.json 2> gpurun_out/bench_water_23k_s8.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_water_23k_s8.json")); print("23k", d["value"], d["ns_per_day"], "ns/day", d["ms_per_step"], "ms | e2e", d["e2e"]["ns_per_day"], d["e2e"]["ms_per_step"], d["phases_ms_per_step"])
except Exception as e: print('no json', e)
PY
MDK_OPTS=pme_cufft=1 timeout 300 python -u bench.py --config water_23k --steps 1500 --warmup 50 --skip-extras 2>/dev/null | tail -1
timeout 300 python -u bench.py --config protein_92k --steps 1000 --warmup 50 --skip-extras 2>/dev/null | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches_r01c_water23k.csv python bench.py --config water_23k --steps 30 --warmup 3 --relax 0.05 --no-graph --skip-extras > gpurun_out/ncu_launch.log 2>&1; echo "ncu launch rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 40 -c 2 -o gpurun_out/prof_pair_r01c_23k -f python bench.py --config water_23k --steps 10 --warmup 3 --relax 0.05 --no-graph --skip-extras > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
