run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --workload grid --sub grid_hmin --no-cpu --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print('grid %.2f compute %.2f coll %.2f hidden %.2f | hmin %.2f compute %.2f coll %.2f' % (r['ms_per_step'], r['ms_compute'], r['ms_collective'], r['collective_hidden'], r['workloads']['grid_hmin']['ms_per_step'], r['workloads']['grid_hmin']['ms_compute'], r['workloads']['grid_hmin']['ms_collective']))
"; }
echo "peer c8"; run
echo "peer c4"; SPHB_Z_CHUNKS=4 run
echo "nccl c2"; SPHB_Z_TRANSPORT=nccl SPHB_Z_CHUNKS=2 run
echo "nccl c4"; SPHB_Z_TRANSPORT=nccl SPHB_Z_CHUNKS=4 run
