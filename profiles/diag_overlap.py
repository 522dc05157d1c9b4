"""Diagnostic (round 2): where the time of the staged, overlapped voxel-grid step goes.
torchrun --nproc-per-node N profiles/diag_overlap.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import sarracen_b200 as s
from sarracen_b200 import ops, parallel
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
o = dist.ProcessGroupNCCL.Options(); o.is_high_priority_stream = True
dist.init_process_group("nccl", device_id=dev, pg_options=o)
wl = bench.describe(os.environ.get("WL", "grid"), 1 << 27, "f64")
cols = bench.generate(wl, rank, world, "f64")
d = {k: torch.from_numpy(v).to(dev) for k, v in cols.items()}
r = ops.Raster(512, 512, 0, 1, 0, 1, 512, 0, 1)
wf = s.CubicSplineKernel().w
out = [torch.empty((512, 512, 512), dtype=torch.float64, device=dev)]
def timed(fn, n=5):
    for _ in range(2): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
res = {}
res["unstaged"] = timed(lambda: ops.scatter3d(d["x"], d["y"], d["z"], d["h"], [d["w"]], wf, 2.0, r, out=out))
for c in (2, 4, 8):
    res[f"staged{c}"] = timed(lambda: ops.scatter3d(d["x"], d["y"], d["z"], d["h"], [d["w"]], wf, 2.0, r, out=out,
                                                     stage_planes=[(k + 1) * 512 // c for k in range(c)], on_stage=lambda k: None))
    res[f"overlap{c}"] = timed(lambda: parallel.scatter3d_sharded(d["x"], d["y"], d["z"], d["h"], [d["w"]], wf, 2.0, r, out=out, chunks=c))
    res[f"sequential{c}"] = timed(lambda: parallel.scatter3d_sharded(d["x"], d["y"], d["z"], d["h"], [d["w"]], wf, 2.0, r, out=out, chunks=c, overlap=False))
slab = torch.empty((512 // world, 512, 512), dtype=torch.float64, device=dev)
res["reduce_scatter_alone"] = timed(lambda: parallel.combine_grid_along_z(out[0], slab))
if rank == 0:
    print({k: round(v, 2) for k, v in res.items()})
dist.destroy_process_group()
