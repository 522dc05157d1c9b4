#!/usr/bin/env python
"""Summarise an ncu report (raw + source pages exported as CSV) into a short text.

    ncu -i X.ncu-rep --page raw --csv > raw.csv ; ncu -i X.ncu-rep --page source --csv > src.csv
    python profiles/ncu_summary.py raw.csv src.csv
"""
import csv
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    print("kernel:", name[:140])
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS or h.startswith("smsp__average_warps_issue_stalled"):
            if h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) < 0.2:
                continue
            print(f"  {h} [{u}] = {v}")


def source(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    tot_i = sum(float(r[idx["Instructions Executed"]] or 0) for r in data)
    tot_s = sum(float(r[idx["# Samples"]] or 0) for r in data)
    groups, cur = [], None
    for n, r in enumerate(data):
        ie = float(r[idx["Instructions Executed"]] or 0)
        sm = float(r[idx["# Samples"]] or 0)
        if cur and abs(cur["ie"] - ie) <= 1e-9 * max(1.0, ie):
            cur["n"] += 1
            cur["samp"] += sm
            cur["end"] = n
        else:
            cur = {"ie": ie, "n": 1, "samp": sm, "start": n, "end": n}
            groups.append(cur)
    print(f"source page: {tot_i:.4g} warp instructions, {tot_s:.0f} samples; basic blocks > 1 %:")
    for g in groups:
        share = g["ie"] * g["n"] / tot_i
        if share > 0.01 or g["samp"] / tot_s > 0.015:
            ops = []
            for i in range(g["start"], g["end"] + 1):
                tok = data[i][idx["Source"]].split()
                ops.append((tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0])
            c = Counter(ops)
            print(f"  sass {g['start']:4d}-{g['end']:4d} n={g['n']:3d} exec={g['ie']:.3e} inst%={100 * share:5.1f} "
                  f"samp%={100 * g['samp'] / tot_s:5.1f} thr={data[g['start']][idx['Avg. Threads Executed']]} "
                  f"{dict(c.most_common(7))}")


if __name__ == "__main__":
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        source(sys.argv[2])
