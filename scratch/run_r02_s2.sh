#!/bin/bash
# experiments after the far-class list order: pair blocks per SM (room for the side streams), skin, the 10 M box, rebuild launch list
O=gpurun_out
run() { # name, config, steps, env...
  local name=$1 cfg=$2 steps=$3; shift 3
  env "$@" python bench.py --config $cfg --no-sub --steps $steps > $O/r02c_$name.json 2> $O/r02c_$name.err
}
run 92k_pbs3 protein_92k 500 MDK_OPTS=pair_blocks_per_sm=3
run 1m_pbs3 protein_1m 100 MDK_OPTS=pair_blocks_per_sm=3
run 23k_pbs3 water_23k 1000 MDK_OPTS=pair_blocks_per_sm=3
run 23k_pbs2 water_23k 1000 MDK_OPTS=pair_blocks_per_sm=2
run 1m_skin25 protein_1m 100 MDK_SKIN=2.5
run 1m_skin30 protein_1m 100 MDK_SKIN=3.0
run 92k_skin25 protein_92k 500 MDK_SKIN=2.5
run 92k_skin30 protein_92k 500 MDK_SKIN=3.0
run 23k_skin25 water_23k 1000 MDK_SKIN=2.5
timeout 600 python bench.py --config water_10m --no-sub --steps 20 --warmup 3 > $O/r02c_10m.json 2> $O/r02c_10m.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_cell|DeviceRadix|k_gather_sorted|k_tables|k_block_bbox|k_build|k_dd_bounds|k_refresh' -c 60 --csv --log-file $O/launches_r02c_rebuild_92k.csv \
    python bench.py --config protein_92k --steps 30 --warmup 3 --relax 0.5 --no-graph --skip-extras > /dev/null 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02c_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); p=d['phases_ms_per_step']
        print(f.split('/')[-1], 'ms %.4f'%d['ms_per_step'], 'pair %.4f nlist %.4f'%(p['pair_ms'],p['nlist_ms']), 'reb/rep', d['config'].get('nlist_rebuilds_per_rep'), 'e2e %.4f'%d['e2e']['ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
python scratch/launch_summary.py $O/launches_r02c_rebuild_92k.csv
