#!/bin/bash
# Round-2 ncu evidence (one B200): launch lists of the default bench command and --set full captures of the
# dominant kernels at full size.  Each ncu run follows a plain run of the same command that exited 0.
set -u
O=gpurun_out
G="python bench.py --workload grid --sub grid_hmin --no-cpu --no-e2e --steps 1 --warmup 1"
P="python bench.py --workload proj --sub none --no-cpu --no-e2e --steps 1 --warmup 1"
$G > $O/plain_grid.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_grid_r2.csv $G > $O/ncu_l1.log 2>&1
$P > $O/plain_proj.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_proj_r2.csv $P > $O/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"cell_kernel|emit_kernel|count_kernel" -s 3 -c 3 -o $O/r2_grid_f64 $G > $O/ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"cell_kernel|emit_kernel|count_kernel" -s 12 -c 3 -o $O/r2_gridhmin_f64 $G > $O/ncu_h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"render_kernel" -s 1 -c 1 -o $O/r2_proj_f64 $P > $O/ncu_p.log 2>&1
tail -2 $O/ncu_g.log $O/ncu_h.log $O/ncu_p.log
