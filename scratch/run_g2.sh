#!/bin/bash
N=${1:-2}; CFG=${2:-protein_92k}; STEPS=${3:-300}
timeout 420 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N --config $CFG --steps $STEPS --warmup 30 > gpurun_out/bench_${CFG}_g$N.json 2> gpurun_out/bench_${CFG}_g$N.err
echo rc=$? ; tail -c 1500 gpurun_out/bench_${CFG}_g$N.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${CFG}_g$N.json")); print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["phases_ms_per_step"], d["config"]["parallelism"])
except Exception as e: print('no json', e); print(open("gpurun_out/bench_${CFG}_g$N.json").read()[:500])
PY
