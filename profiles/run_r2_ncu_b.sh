#!/bin/bash
# Round-2 ncu evidence, second set (one B200): the lattice-form exact kernels and the round-end cell kernel.
# Each ncu run follows a plain run of the same command that exited 0.
set -u
O=gpurun_out
X="python profiles/exact_probe.py"
G="python bench.py --workload grid --sub none --no-cpu --no-e2e --steps 1 --warmup 1"
$X > $O/plain_exact.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"exact_lattice_kernel|exact2_lines_kernel" -s 2 -c 3 -o $O/r2_exact $X > $O/ncu_x.log 2>&1
$G > $O/plain_grid.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"cell_kernel" -s 1 -c 1 -o $O/r2_grid_f64_v2 $G > $O/ncu_g2.log 2>&1
tail -2 $O/plain_exact.log $O/ncu_x.log $O/ncu_g2.log
