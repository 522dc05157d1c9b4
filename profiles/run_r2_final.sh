#!/bin/bash
# Round-2 closing measurement set (one B200): GPU tests, the default bench line, the reference arm, then ncu:
# launch lists of the bench commands and --set full captures of the round-end cell / emit / count kernels.
# Each ncu run follows a plain run of the same command that exited 0.
set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > $O/bench_default_r2.json 2> $O/bench_default_r2.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_r2.json 2> $O/bench_reference_r2.err; echo "reference rc=$?"
G="python bench.py --workload grid --sub grid_hmin --no-cpu --no-e2e --steps 1 --warmup 1"
P="python bench.py --workload proj --sub none --no-cpu --no-e2e --steps 1 --warmup 1"
$G > $O/plain_grid.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_grid_r2.csv $G > $O/ncu_l1.log 2>&1
$P > $O/plain_proj.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_proj_r2.csv $P > $O/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"cell_kernel|emit_kernel|count_kernel" -s 3 -c 3 -o $O/r2_grid_f64 $G > $O/ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"cell_kernel" -s 3 -c 1 -o $O/r2_gridhmin_f64 $G > $O/ncu_h.log 2>&1
tail -n 2 $O/ncu_g.log $O/ncu_h.log
