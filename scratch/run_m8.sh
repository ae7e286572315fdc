#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N --steps 200 --warmup 20 --relax 0.3 > gpurun_out/bench_protein_1m_g${N}_s9.json 2> gpurun_out/bench_protein_1m_g${N}_s9.err
echo "1m g$N rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_protein_1m_g${N}_s9.err | tail -c 800
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/bench_protein_1m_g${N}_s9.json") if l.startswith('{')][-1]
    print(d["config"]["workload"], d["n_gpus"], d["value"], d["ns_per_day"], "ns/day", d["ms_per_step"], "ms | e2e", d["e2e"]["ns_per_day"], d["e2e"]["ms_per_step"], d["phases_ms_per_step"], "rebuilds", d["config"]["nlist_rebuilds_in_timed"], d["config"]["parallelism"])
    print('single-gpu phases', d['phases_ms_per_step_single_gpu'])
except Exception as e: print('no json', e)
PY
