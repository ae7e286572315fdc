#!/bin/bash
# Round-2 profiling pass (one B200, under gpurun).
#  1. launch lists (ncu --metrics gpu__time_duration.sum; cold-cache and serialised: compare SHARES) of the three boxes,
#     kernels launched one by one from the host (--no-graph) so that ncu sees plain launches;
#  2. one full capture per kernel the review asked evidence for, at the 92k target config (separate ncu passes so that
#     every kernel is inside its window), + the filter-then-compute pair variant for the record.
set -x
OUT=gpurun_out
for cfg in water_23k protein_92k protein_1m; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 500 --csv --log-file $OUT/launches_r02_$cfg.csv \
      python bench.py --config $cfg --steps 30 --warmup 3 --relax 0.5 --no-graph --skip-extras > $OUT/ncu_launch_$cfg.log 2>&1
done
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum"
for k in k_pair k_build_lists k_spread_smem k_gather k_aux_terms k_langevin k_mesh k_gather_sorted k_tables_sorted; do
  ncu --metrics $M --clock-control none -k regex:"$k" -s 6 -c 3 --csv --log-file $OUT/ncu_r02_92k_$k.csv \
      python bench.py --config protein_92k --steps 40 --warmup 3 --relax 0.5 --no-graph --skip-extras > /dev/null 2>&1
done
# mesh kernels live in the 23k box (64^3 fused FFT kernels)
ncu --metrics $M --clock-control none -k regex:'k_mesh|k_spread_smem|k_gather<' -s 9 -c 9 --csv --log-file $OUT/ncu_r02_23k_mesh.csv \
    python bench.py --config water_23k --steps 40 --warmup 3 --relax 0.5 --no-graph --skip-extras > /dev/null 2>&1
MDK_OPTS=pair_v5=1 ncu --metrics $M --clock-control none -k regex:k_pair5 -s 6 -c 2 --csv --log-file $OUT/ncu_r02_92k_k_pair5.csv \
    python bench.py --config protein_92k --steps 40 --warmup 3 --relax 0.5 --no-graph --skip-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_pair<' -s 6 -c 1 -o $OUT/prof_r02_pair_92k \
    python bench.py --config protein_92k --steps 40 --warmup 3 --relax 0.5 --no-graph --skip-extras > /dev/null 2>&1
ncu -i $OUT/prof_r02_pair_92k.ncu-rep --page raw --csv > $OUT/prof_r02_pair_92k_raw.csv 2>/dev/null
ls -la $OUT | tail -30
