#!/bin/bash
# final round-2 code at N GPUs: NCCL parity test, then the driver's command
N=${1:-2}; O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | grep -v "NCCL INFO" | tail -8 > $O/r02f_multi${N}_test.log; tail -5 $O/r02f_multi${N}_test.log
timeout 900 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N --steps 200 --warmup 20 > $O/r02f_bench_n$N.json 2> $O/r02f_bench_n$N.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r02f_bench_n$N.json") if l.startswith('{')][-1]
    print(d["config"]["workload"], d["n_gpus"], d["ns_per_day"], "ns/day", d["ms_per_step"], "ms | e2e", d["e2e"]["ms_per_step"], d["phases_ms_per_step"], d["config"].get("domain_decomposition"))
except Exception as e: print('no json', e)
PY
