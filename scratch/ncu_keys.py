import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys=['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__waves_per_multiprocessor',
'sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum',
'smsp__thread_inst_executed_per_inst_executed.ratio','sm__warps_active.avg.pct_of_peak_sustained_active',
'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
'dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
'lts__t_bytes.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__cycles_elapsed.max',
'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum','smsp__sass_thread_inst_executed_op_fmul_pred_on.sum','smsp__sass_thread_inst_executed_op_fadd_pred_on.sum',
'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']
for r in rows[2:]:
    print('-----')
    for k in keys:
        if k in hdr:
            i=hdr.index(k); print('%-86s %-16s %s'%(k,units[i],r[i][:90]))
