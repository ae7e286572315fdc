#!/bin/bash
mkdir -p gpurun_out
run() { # cfg tag opts
  MDK_OPTS=$3 timeout 300 python -u bench.py --config $1 --steps 1500 --warmup 50 --skip-extras 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$1 $2', round(d['ms_per_step']*1000, 1), 'us', round(d['ns_per_day'], 1), 'ns/day rebuilds', d['rebuilds'])"
}
for CFG in water_23k protein_92k; do
  run $CFG base ""
  run $CFG bps3 "pair_blocks_per_sm=3"
  run $CFG bps2 "pair_blocks_per_sm=2"
  run $CFG bps5 "pair_blocks_per_sm=5"
done
