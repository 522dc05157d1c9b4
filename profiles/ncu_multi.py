#!/usr/bin/env python
"""Key metrics of EVERY kernel of an ncu report (raw page exported as CSV):

    ncu -i X.ncu-rep --page raw --csv > raw.csv ; python profiles/ncu_multi.py raw.csv
"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]
KEYS=["gpu__time_duration.sum","launch__registers_per_thread","launch__grid_size","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__thread_inst_executed_per_inst_executed.ratio","sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
units=rows[1]
for r in rows[2:]:
    d=dict(zip(hdr,r)); u=dict(zip(hdr,units))
    print("kernel:", d['Kernel Name'][:110])
    for k in KEYS:
        if k in d: print(f"  {k} [{u[k]}] = {d[k]}")
    for k,v in d.items():
        if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio') and float(v or 0)>0.3: print(f"  {k} = {v}")
