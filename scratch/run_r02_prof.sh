#!/bin/bash
# Round-2 profiling pass (one B200, under gpurun): launch lists (ncu --metrics gpu__time_duration.sum, cold-cache and
# serialised: compare SHARES) and one full capture of each kernel the review asked evidence for.
set -x
OUT=gpurun_out
for cfg in water_23k protein_92k; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 500 --csv --log-file $OUT/launches_r02_$cfg.csv \
      python bench.py --config $cfg --steps 30 --warmup 3 --relax 0.3 --no-graph --skip-extras > $OUT/ncu_launch_$cfg.log 2>&1
done
# full captures at 92k (the target config): pair kernel, list builder, spread, gather, excluded-pair correction, Langevin update
ncu --set full --clock-control none --import-source on -k regex:'k_pair|k_build_lists|k_spread|k_gather<|k_excl|k_langevin|k_bonds|k_angles' \
    -s 400 -c 24 -o $OUT/prof_r02_92k python bench.py --config protein_92k --steps 12 --warmup 3 --relax 0.3 --no-graph --skip-extras > $OUT/ncu_full_92k.log 2>&1
ncu -i $OUT/prof_r02_92k.ncu-rep --page raw --csv > $OUT/prof_r02_92k_raw.csv 2>/dev/null
ls -la $OUT | tail -20
