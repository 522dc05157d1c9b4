# experiment (round 2): warps per CTA of the cell kernel x cell size, config 4 (grid) device pipeline
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 2 --sub none --workload grid"
P='import json,sys; d=json.loads(sys.stdin.read()); print({k:round(v,2) for k,v in d["phases_ms"].items()}, round(d["ms_per_step"],2))'
for lib in "" _w24 _w32; do for c in 2 3; do echo "lib=$lib CELL_LOG=$c"; SPHB_LIB_OVERRIDE=$PWD/sarracen_b200/lib/libsphb200$lib.so SPHB_CELL_LOG=$c $B 2>/dev/null | python -c "$P"; done; done
