#!/bin/bash
run() { # tag env...
  tag=$1; shift
  env "$@" timeout 300 python -u bench.py --config water_23k --steps 1500 --warmup 50 --skip-extras 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$tag', round(d['ms_per_step']*1000, 1), 'us rebuilds', d['rebuilds'], 'launches/step', round(d['gpu_launches']/d['steps'],1))"
}
run all X=1
run no_pme MDK_TERMS_MASK=0xfb
run no_bonded MDK_TERMS_MASK=0x0f
run pair_only MDK_TERMS_MASK=0x03
run lj_only MDK_TERMS_MASK=0x01
run serial MDK_OPTS=concurrent=0
run nograph MDK_OPTS=graph=0
