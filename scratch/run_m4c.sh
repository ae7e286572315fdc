#!/bin/bash
N=4
run() { tag=$1; shift
  env "$@" timeout 200 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N --steps 100 --warmup 20 --relax 0.3 --skip-extras 2>gpurun_out/m4c_$tag.err | grep '^{' | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$tag', d['ms_per_step'], 'rebuilds', d['rebuilds'])"
}
run ring NCCL_NVLS_ENABLE=0 NCCL_ALGO=Ring
run nograph MDK_OPTS=graph=0
