"""Diagnostic (round 2): which part of the peer-memory exchange overlaps the staged render.
torchrun --nproc-per-node N profiles/diag_peer_overlap.py"""
import os, sys
import torch, torch.distributed as dist
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
if wl["key"] != "proj" and "h" not in d:
    pass
r = ops.Raster(512, 512, 0, 1, 0, 1, 512, 0, 1)
wf = s.CubicSplineKernel().w
out = [torch.empty((512, 512, 512), dtype=torch.float64, device=dev)]
C = int(os.environ.get("CHUNKS", "4"))
per, b = 512 // C, 512 // (C * world)
slot = b * 512 * 512
peer = parallel._peer_slots(None, dev, C * world * slot)
hdl = peer.hdl
side = parallel._side_stream(dev)
main = torch.cuda.current_stream(dev)
mine = torch.empty((C * b, 512, 512), dtype=torch.float64, device=dev)


def timed(fn, n=5):
    for _ in range(2): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def staged(copies, barriers, sums, copy_to_self=False):
    events = [torch.cuda.Event() for _ in range(C)]
    ops.scatter3d(d["x"], d["y"], d["z"], d["h"], [d["w"]], wf, 2.0, r, out=out,
                  stage_planes=[(k + 1) * per for k in range(C)], on_stage=lambda k: events[k].record(main))
    with torch.cuda.stream(side):
        if barriers: hdl.barrier(channel=0)
        for k in range(C):
            side.wait_event(events[k])
            if copies:
                for step in range(1, world):
                    q = rank if copy_to_self else (rank + step) % world
                    dst = hdl.get_buffer(q, (slot,), torch.float64, (k * world + rank) * slot)
                    dst.copy_(out[0][k * per + q * b:k * per + (q + 1) * b].reshape(-1), non_blocking=True)
            if barriers: hdl.barrier(channel=1)
            if sums:
                recv = peer.buf[k * world * slot:(k + 1) * world * slot].view(world, slot)
                ops.sum_slots_(mine[k * b:(k + 1) * b], out[0][k * per + rank * b:k * per + (rank + 1) * b], recv, rank)
    main.wait_stream(side)


res = {}
res["unstaged"] = timed(lambda: ops.scatter3d(d["x"], d["y"], d["z"], d["h"], [d["w"]], wf, 2.0, r, out=out))
res["staged"] = timed(lambda: staged(False, False, False))
res["staged+copies_to_self"] = timed(lambda: staged(True, False, False, True))
res["staged+copies"] = timed(lambda: staged(True, False, False))
res["staged+barriers"] = timed(lambda: staged(False, True, False))
res["staged+copies+barriers"] = timed(lambda: staged(True, True, False))
res["staged+all"] = timed(lambda: staged(True, True, True))
res["sharded_peer"] = timed(lambda: parallel.scatter3d_sharded(d["x"], d["y"], d["z"], d["h"], [d["w"]], wf, 2.0, r, out=out, chunks=C))
res["sharded_nccl"] = timed(lambda: parallel.scatter3d_sharded(d["x"], d["y"], d["z"], d["h"], [d["w"]], wf, 2.0, r, out=out, chunks=2, transport="nccl"))
res["exchange_alone_peer"] = timed(lambda: parallel.reduce_along_z(out[0], chunks=C))
res["exchange_alone_nccl"] = timed(lambda: parallel.reduce_along_z(out[0], chunks=1, transport="nccl"))
if rank == 0:
    print({k: round(v, 2) for k, v in res.items()})
dist.destroy_process_group()
