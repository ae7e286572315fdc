#!/bin/bash
# decomposed step: merged halo + sub-mesh exchange (default), potential boxes received beside the pair kernel (option), A/B
N=${1:-2}; O=gpurun_out; shift
if [ "$1" = "test" ]; then shift
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | grep -v "NCCL INFO" | tail -6 > $O/r02h_multi${N}_test.log; tail -4 $O/r02h_multi${N}_test.log
fi
for opts in "$@"; do
  tag=$(echo $opts | tr '=,' '__')
  MDK_OPTS=$opts timeout 150 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N --steps 200 --warmup 20 > $O/r02h_bench_n${N}_$tag.json 2> $O/r02h_bench_n${N}_$tag.err
  echo "$opts rc=$?"
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r02h_bench_n${N}_$tag.json") if l.startswith('{')][-1]
    print('$opts', d["ns_per_day"], "ns/day", d["ms_per_step"], "ms | e2e", d["e2e"]["ms_per_step"], 'E', d["e2e"].get("potential_energy_last_step"), d["phases_ms_per_step"])
except Exception as e: print('no json', e)
PY
done
