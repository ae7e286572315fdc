#!/bin/bash
# Final single-GPU pass of round 2: tests, default bench (the driver's command), the 20-step check, the reference arm, smoke,
# then the profiling pass (launch lists with graphs off, per-kernel counters at 92k, pair-kernel traffic on all three boxes,
# one full capture of the pair kernel).
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py 2>&1 | tail -15 > $O/r02f_tests.log; tail -4 $O/r02f_tests.log
python bench.py > $O/r02f_bench_n1.json 2> $O/r02f_bench_n1.err
python bench.py --steps 20 --warmup 5 --no-sub > $O/r02f_bench_n1_k20.json 2>/dev/null
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02f_bench_ref.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02f_smoke.log 2>&1; tail -2 $O/r02f_smoke.log
for cfg in water_23k protein_92k protein_1m; do
  # skip the relaxation phase ((210 + 20 + 3) steps x ~14.5 launches with graphs off): the window lies in the timed production steps
  ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 700 --csv --log-file $O/launches_r02f_$cfg.csv \
      python bench.py --config $cfg --steps 40 --warmup 3 --relax 0.3 --no-graph --skip-extras > $O/ncu_launch_$cfg.log 2>&1
done
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum"
for k in k_pair k_build_lists k_spread_smem k_gather k_aux_terms k_langevin; do
  ncu --metrics $M --clock-control none -k regex:"^$k\$" -s 6 -c 2 --csv --log-file $O/ncu_r02f_92k_$k.csv \
      python bench.py --config protein_92k --steps 40 --warmup 3 --relax 1.0 --no-graph --skip-extras > /dev/null 2>&1
done
for cfg in water_23k protein_1m; do
  ncu --metrics $M --clock-control none -k regex:'^k_pair$' -s 6 -c 2 --csv --log-file $O/ncu_r02f_${cfg}_k_pair.csv \
      python bench.py --config $cfg --steps 30 --warmup 3 --relax 1.0 --no-graph --skip-extras > /dev/null 2>&1
done
ncu --metrics $M --clock-control none -k regex:'k_mesh|k_spread_smem|^k_gather$' -s 9 -c 9 --csv --log-file $O/ncu_r02f_23k_mesh.csv \
    python bench.py --config water_23k --steps 40 --warmup 3 --relax 1.0 --no-graph --skip-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^k_pair$' -s 6 -c 1 -f -o $O/prof_r02f_pair_92k \
    python bench.py --config protein_92k --steps 40 --warmup 3 --relax 1.0 --no-graph --skip-extras > $O/ncu_full_r02f.log 2>&1
ncu -i $O/prof_r02f_pair_92k.ncu-rep --page raw --csv > $O/prof_r02f_pair_92k_raw.csv 2>/dev/null
ls -la $O | tail -5
