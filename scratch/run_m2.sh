#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 gpurun_out/pytest_multi.log
for CFG in protein_92k protein_1m; do
  STEPS=300; [ $CFG = protein_1m ] && STEPS=100
  timeout 600 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N --config $CFG --steps $STEPS --warmup 20 > gpurun_out/bench_${CFG}_g${N}_s3.json 2> gpurun_out/bench_${CFG}_g${N}_s3.err
  echo "$CFG g$N rc=$?"; tail -c 600 gpurun_out/bench_${CFG}_g${N}_s3.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/bench_${CFG}_g${N}_s3.json") if l.startswith('{')][-1]
    print("$CFG", d["n_gpus"], d["value"], d["ns_per_day"], "ns/day", d["ms_per_step"], "ms | e2e", d["e2e"]["ns_per_day"], d["e2e"]["ms_per_step"], d["phases_ms_per_step"], "rebuilds", d["config"]["nlist_rebuilds_in_timed"], d["config"]["parallelism"])
except Exception as e: print('no json', e)
PY
done
