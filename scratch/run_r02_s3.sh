#!/bin/bash
# builder row narrowing (parity), short-lived pair blocks (pair_units_per_warp) A/B
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py 2>&1 | tail -15 > $O/r02d_tests.log; tail -4 $O/r02d_tests.log
run() { local name=$1 cfg=$2 steps=$3; shift 3
  env "$@" python bench.py --config $cfg --no-sub --steps $steps > $O/r02d_$name.json 2> $O/r02d_$name.err; }
for u in 0 1 2 4; do
  run 92k_u$u protein_92k 500 MDK_OPTS=pair_units_per_warp=$u
  run 23k_u$u water_23k 1000 MDK_OPTS=pair_units_per_warp=$u
done
for u in 1 2; do run 1m_u$u protein_1m 100 MDK_OPTS=pair_units_per_warp=$u; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02d_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); p=d['phases_ms_per_step']
        print(f.split('/')[-1], 'ms %.4f'%d['ms_per_step'], 'pair %.4f nlist %.4f'%(p['pair_ms'],p['nlist_ms']), 'reb/rep', d['config'].get('nlist_rebuilds_per_rep'), 'e2e %.4f'%d['e2e']['ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
