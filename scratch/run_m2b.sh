#!/bin/bash
N=${1:-2}
timeout 500 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N --steps 100 --warmup 20 --relax 0.3 > gpurun_out/bench_protein_1m_g${N}_dbg.json 2> gpurun_out/bench_protein_1m_g${N}_dbg.err
echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_protein_1m_g${N}_dbg.err | tail -c 500
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_protein_1m_g${N}_dbg.json") if l.startswith('{')][-1]
print(d["n_gpus"], d["value"], d["ms_per_step"], "rebuilds", d["config"]["nlist_rebuilds_in_timed"], 'Epot', d['e2e']['potential_energy_last_step'], d["phases_ms_per_step"])
PY
