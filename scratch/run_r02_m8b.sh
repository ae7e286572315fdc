#!/bin/bash
# N = 8: list skin 3.0 A instead of 2.0 (fewer rebuilds; the skin shell is cheap since the far-class list order)
N=8; O=gpurun_out
for sk in 3.0; do
MDK_SKIN=$sk timeout 600 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N --steps 200 --warmup 20 > $O/r02f_bench_n${N}_skin$sk.json 2> $O/r02f_bench_n${N}_skin$sk.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r02f_bench_n${N}_skin$sk.json") if l.startswith('{')][-1]
    print('skin $sk', d["ns_per_day"], "ns/day", d["ms_per_step"], "ms | e2e", d["e2e"]["ms_per_step"], d["phases_ms_per_step"], d["config"].get("domain_decomposition"))
except Exception as e: print('no json', e)
PY
done
