"""NCCL correctness check of the multi-GPU path (needs >= 2 GPUs; run by tests/test_gpu_multi.py under
torchrun, or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py

1. Every rank renders its particle shard on its own GPU; images are combined with an NCCL reduce, voxel
   grids slab by slab along z with the reduction overlapping the render (`scatter3d_sharded`) and with the
   monolithic reduce-scatter; both are compared with the whole set rendered on one GPU and with the CPU oracle.
2. `parallel.enable()`: the nine-method backend API called with the SAME full NumPy columns on every rank
   returns the combined result (images on all ranks; grids whole or as z-slabs).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sarracen_b200 as s  # noqa: E402
from sarracen_b200 import ops, parallel  # noqa: E402
from sarracen_b200.parallel import (combine_grid_along_z, combine_image, owned_planes, scatter3d_sharded,  # noqa: E402
                                    shard_range, slab_range)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
_opts = dist.ProcessGroupNCCL.Options()
_opts.is_high_priority_stream = True
dist.init_process_group("nccl", device_id=dev, pg_options=_opts)

rng = np.random.default_rng(42)
n = 400_003
x, y, z = rng.uniform(0, 1, (3, n))
h = np.exp(rng.uniform(np.log(0.002), np.log(0.08), n))
w = rng.uniform(0.1, 1, n) * h ** 3
w2 = rng.normal(size=n) * h ** 3
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
xd, yd, zd, hd, wd = (t(a) for a in (x, y, z, h, w))
lo, hi = shard_range(n, rank, world)
k = s.QuinticSplineKernel()
col = k.get_column_kernel_func(1000)
c = s.CubicSplineKernel()


def relmax(a, b):
    return float((a - b).abs().max() / b.abs().max())


# ---- 1. explicit sharding -----------------------------------------------------------------------
r2 = ops.Raster(300, 200, 0, 1, 0, 1)
part = ops.scatter2d(xd[lo:hi], yd[lo:hi], hd[lo:hi], [wd[lo:hi]], col, 3.0, 2, r2)[0]
combine_image(part, dst=0)
nz = 16 * world + 3          # unequal slabs
r3 = ops.Raster(40, 36, 0, 1, 0, 1, nz, 0, 1)
full2 = ops.scatter2d(xd, yd, hd, [wd], col, 3.0, 2, r2)[0]
full3 = ops.scatter3d(xd, yd, zd, hd, [wd], c.w, 2.0, r3)[0]
zl, zh = slab_range(nz, rank, world)
errs = []
for overlap in (True, False):
    slab = scatter3d_sharded(xd[lo:hi], yd[lo:hi], zd[lo:hi], hd[lo:hi], [wd[lo:hi]], c.w, 2.0, r3, overlap=overlap)[0]
    errs.append(relmax(slab, full3[owned_planes(nz, rank, world)]))
gpart = ops.scatter3d(xd[lo:hi], yd[lo:hi], zd[lo:hi], hd[lo:hi], [wd[lo:hi]], c.w, 2.0, r3)[0]
slab2 = torch.empty((zh - zl, 36, 40), dtype=torch.float64, device=dev)
combine_grid_along_z(gpart, slab2)
errs.append(relmax(slab2, full3[zl:zh]))
# a grid whose small particles take the cell path and whose large ones the tile path, 64 planes per rank
r3b = ops.Raster(64, 64, 0, 1, 0, 1, 64 * world, 0, 1)
full3b = ops.scatter3d(xd, yd, zd, hd, [wd], c.w, 2.0, r3b)[0]
used = []
for tr in ("peer", "nccl"):       # copy-engine pushes through symmetric memory / NCCL reduce-scatter
    for ch in (4, 1):             # block-cyclic planes (4 z-chunks, overlapped) and contiguous slabs
        slab3 = scatter3d_sharded(xd[lo:hi], yd[lo:hi], zd[lo:hi], hd[lo:hi], [wd[lo:hi]], c.w, 2.0, r3b, chunks=ch,
                                  transport=tr)[0]
        used.append(parallel.last_transport)
        errs.append(relmax(slab3, full3b[owned_planes(64 * world, rank, world, ch)]))
    # two weights in one exchange, and the exchange of a finished grid
    a3, b3 = scatter3d_sharded(xd[lo:hi], yd[lo:hi], zd[lo:hi], hd[lo:hi], [wd[lo:hi], hd[lo:hi]], c.w, 2.0, r3b,
                               transport=tr)
    fa, fb = ops.scatter3d(xd, yd, zd, hd, [wd, hd], c.w, 2.0, r3b)
    own = owned_planes(64 * world, rank, world)
    errs += [relmax(a3, fa[own]), relmax(b3, fb[own])]
    gp = ops.scatter3d(xd[lo:hi], yd[lo:hi], zd[lo:hi], hd[lo:hi], [wd[lo:hi]], c.w, 2.0, r3b)[0]
    errs.append(relmax(parallel.reduce_along_z(gp, transport=tr), full3b[own]))
# one plane of an odd number of pixels per rank and chunk: the receive slots are not 16-byte multiples
r3c = ops.Raster(33, 35, 0, 1, 0, 1, 8 * world, 0, 1)
full3c = ops.scatter3d(xd, yd, zd, hd, [wd], c.w, 2.0, r3c)[0]
slab3c = scatter3d_sharded(xd[lo:hi], yd[lo:hi], zd[lo:hi], hd[lo:hi], [wd[lo:hi]], c.w, 2.0, r3c, chunks=8,
                           transport="peer")[0]
errs.append(relmax(slab3c, full3c[owned_planes(8 * world, rank, world, 8)]))
assert owned_planes(64 * world, rank, world, 1) == list(range(*slab_range(64 * world, rank, world)))
e = torch.tensor(errs, device=dev, dtype=torch.float64)
dist.all_reduce(e, op=dist.ReduceOp.MAX)

# ---- 2. the backend API in multi-GPU mode ----------------------------------------------------------
B = s.B200Backend
single = {
    "proj": B.interpolate_3d_projection(x, y, w, h, col, 3.0, 300, 200, 0, 1, 0, 1, False),
    "grid": B.interpolate_3d_grid(x, y, z, w, h, c.w, 2.0, 40, 36, nz, 0, 1, 0, 1, 0, 1),
    "cvec": B.interpolate_3d_cross_vec(x, y, z, 0.45, w, w2, h, c.w, 2.0, 120, 90, 0, 1, 0, 1),
    "line": B.interpolate_3d_line(x, y, z, w, h, c.w, 2.0, 77, 0.1, 0.9, 0.2, 0.8, 0.3, 0.7),
    "exact": B.interpolate_2d_render(x[:3000], y[:3000], w[:3000], h[:3000], c.w, 2.0, 64, 48, 0, 1, 0, 1, True),
}
parallel.enable(images="all", grids="all")
multi = {
    "proj": B.interpolate_3d_projection(x, y, w, h, col, 3.0, 300, 200, 0, 1, 0, 1, False),
    "grid": B.interpolate_3d_grid(x, y, z, w, h, c.w, 2.0, 40, 36, nz, 0, 1, 0, 1, 0, 1),
    "cvec": B.interpolate_3d_cross_vec(x, y, z, 0.45, w, w2, h, c.w, 2.0, 120, 90, 0, 1, 0, 1),
    "line": B.interpolate_3d_line(x, y, z, w, h, c.w, 2.0, 77, 0.1, 0.9, 0.2, 0.8, 0.3, 0.7),
    "exact": B.interpolate_2d_render(x[:3000], y[:3000], w[:3000], h[:3000], c.w, 2.0, 64, 48, 0, 1, 0, 1, True),
}
parallel.enable(images="root", grids="slab")
slab_api = B.interpolate_3d_grid(x, y, z, w, h, c.w, 2.0, 40, 36, nz, 0, 1, 0, 1, 0, 1)
parallel.disable()
api = []
for name in single:
    a, b = single[name], multi[name]
    if isinstance(a, tuple):
        api += [np.abs(a[i] - b[i]).max() / np.abs(a[i]).max() for i in range(2)]
    else:
        api.append(np.abs(a - b).max() / np.abs(a).max())
api.append(np.abs(slab_api - single["grid"][owned_planes(nz, rank, world)]).max() / np.abs(single["grid"]).max())
parallel.enable(images="all", grids="all")   # a grid that divides evenly: chunked, block-cyclic, gathered
g_multi = B.interpolate_3d_grid(x, y, z, w, h, c.w, 2.0, 64, 64, 64 * world, 0, 1, 0, 1, 0, 1)
parallel.disable()
api.append(np.abs(g_multi - full3b.cpu().numpy()).max() / float(full3b.abs().max()))
assert slab_api.shape == (zh - zl, 36, 40)
ea = torch.tensor(api, device=dev, dtype=torch.float64)
dist.all_reduce(ea, op=dist.ReduceOp.MAX)

if rank == 0:
    e2 = relmax(part, full2)
    from oracle import OracleBackend, OracleKernel
    ref = OracleBackend.interpolate_3d_projection(x, y, w, h, OracleKernel.column("quintic"), 3.0, 300, 200, 0, 1, 0, 1,
                                                  False)
    eo = float(np.abs(part.cpu().numpy() - ref).max() / np.abs(ref).max())
    print(f"world={world} image: sharded vs single {e2:.2e}, vs oracle {eo:.2e}; grid slabs vs single "
          f"(unequal slabs overlapped / sequential, reduce-scatter, 4 z-chunks, contiguous) {[f'{v:.1e}' for v in e.tolist()]}; "
          f"backend API in multi-GPU mode vs single {ea.max().item():.2e}; transports used {sorted(set(used))}")
    assert e2 < 1e-12 and eo < 1e-10 and e.max().item() < 1e-12 and ea.max().item() < 1e-12
    print("multi-GPU check ok")
dist.destroy_process_group()
