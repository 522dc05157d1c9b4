run() { python bench.py --workload grid --emulate-shard $1 --sub none --no-cpu --no-e2e --steps 5 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('%.2f cell %.2f bin %.2f' % (r['ms_per_step'], r['phases_ms']['ms_direct'], r['phases_ms']['ms_bin']))"; }
for sh in 8 4 1; do
for cfg in "2 2 3" "2 2 4" "2 3 3" "2 3 4" "3 3 2" "2 2 2"; do set -- $cfg; echo -n "shard $sh  x$1 y$2 z$3: "; SPHB_CELL_LOG=$1 SPHB_CELL_YLOG=$2 SPHB_CELL_ZLOG=$3 run $sh; done; done
