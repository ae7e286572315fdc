#!/bin/bash
# pair kernel: folded switch / erfc constants (parity), unroll 4 / 8 / 16, 4-warp blocks (8 per SM)
O=gpurun_out
L=mdpy_b200/libmdpyb200
cp $L.so /tmp/lib_default.so
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchmark_parity.py -m gpu -q -x 2>&1 | tail -8 > $O/r02e_tests.log; tail -3 $O/r02e_tests.log
run() { local name=$1 cfg=$2 steps=$3; shift 3
  env "$@" python bench.py --config $cfg --no-sub --steps $steps > $O/r02e_$name.json 2> $O/r02e_$name.err; }
run 92k_default protein_92k 500 A=1
run 23k_default water_23k 1000 A=1
run 1m_default protein_1m 100 A=1
for v in u4 u16; do cp ${L}_$v.so $L.so; run 92k_$v protein_92k 500 A=1; run 23k_$v water_23k 1000 A=1; done
cp ${L}_w4.so $L.so; run 92k_w4 protein_92k 500 MDK_OPTS=pair_blocks_per_sm=8; run 23k_w4 water_23k 1000 MDK_OPTS=pair_blocks_per_sm=8
cp /tmp/lib_default.so $L.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02e_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); p=d['phases_ms_per_step']
        print(f.split('/')[-1], 'ms %.4f'%d['ms_per_step'], 'pair %.4f nlist %.4f'%(p['pair_ms'],p['nlist_ms']), 'frac %.3f'%d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
PY
