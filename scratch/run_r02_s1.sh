#!/bin/bash
# far-class list order + lane=candidate list filter: parity, then A/B on the three boxes
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py 2>&1 | tail -15 > $O/r02b_tests_1.log; tail -4 $O/r02b_tests_1.log
for fs in 0 1; do
  MDK_OPTS=far_split=$fs python bench.py --config protein_92k --no-sub --steps 500 > $O/r02b_92k_fs$fs.json 2> $O/r02b_92k_fs$fs.err
  MDK_OPTS=far_split=$fs python bench.py --config water_23k --no-sub --steps 1000 > $O/r02b_23k_fs$fs.json 2> $O/r02b_23k_fs$fs.err
done
python bench.py --no-sub --steps 100 > $O/r02b_1m_fs1.json 2> $O/r02b_1m_fs1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); p=d['phases_ms_per_step']
        print(f.split('/')[-1], 'ms %.4f'%d['ms_per_step'], 'pair %.4f nlist %.4f'%(p['pair_ms'],p['nlist_ms']), 'frac %.3f'%d['roofline']['frac'], d['nlist'])
    except Exception as e: print(f, 'ERR', e)
PY
