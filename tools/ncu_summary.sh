#!/bin/bash
# usage: ncu_summary.sh report.ncu-rep  -> key metrics (with units) of the first kernel in the report
ncu -i "$1" --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr,units=rows[0],rows[1]
keys=['gpu__time_duration.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct','smsp__warps_active.avg.per_cycle_active','smsp__warps_eligible.avg.per_cycle_active','pcsamp_warps_issue_stalled','thread_inst_executed_per_inst','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct','dram__bytes_read.sum ','dram__bytes_write.sum ','sass__inst_executed_local','sm__throughput.avg.pct','gpu__dram_throughput.avg.pct','Kernel Name','smsp__sass_thread_inst_executed_op_dfma_pred_on.sum ','smsp__sass_thread_inst_executed_op_dmul_pred_on.sum ','smsp__sass_thread_inst_executed_op_dadd_pred_on.sum ']
for r in rows[2:3]:
    for h,u,v in sorted(zip(hdr,units,r)):
        if any(k in h+' ' for k in keys) and 'not_issued' not in h and not (h.startswith('smsp__pcsamp') and v in ('0','')): print('%s [%s] = %s'%(h,u,v))
"
