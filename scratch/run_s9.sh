#!/bin/bash
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/smoke.log
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time; echo "bench rc=$?"; cat gpurun_out/bench_default.time | tail -4; tail -c 300 gpurun_out/bench_default.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_default.json"))
print({k: d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','ns_per_day','gpu_launches','clocks')})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['ns_per_day'])
print('roofline', d['roofline']['frac'], d['roofline']['traffic'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['ns_per_day'])
PY
