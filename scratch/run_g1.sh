#!/bin/bash
CFG=${1:-water_23k}; STEPS=${2:-1000}; TAG=${3:-x}
timeout 420 python -u bench.py --config $CFG --steps $STEPS --warmup 50 > gpurun_out/bench_${CFG}_${TAG}.json 2> gpurun_out/bench_${CFG}_${TAG}.err
echo rc=$? ; tail -c 600 gpurun_out/bench_${CFG}_${TAG}.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${CFG}_${TAG}.json")); print("$CFG", d["value"], "ns/day", d["ms_per_step"], "ms | e2e", d["e2e"]["value"], "| roofline", round(d["roofline"]["frac"],4), "| slot eff", round(d["nlist"]["slot_efficiency"],3), d["phases_ms_per_step"], "rebuilds", d["config"]["nlist_rebuilds_in_timed"])
except Exception as e: print('no json', e)
PY
