"""NCCL correctness check of the sharded pipeline (needs >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py

Every rank renders its particle shard on its own GPU; images are combined with NCCL reduce, voxel grids with
reduce_scatter along z; the result is compared with the whole set rendered on one GPU and with the CPU oracle.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sarracen_b200 as s  # noqa: E402
from sarracen_b200 import ops  # noqa: E402
from sarracen_b200.parallel import combine_grid_along_z, combine_image, shard_range, slab_range  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

rng = np.random.default_rng(42)
n = 200_003
x, y, z = rng.uniform(0, 1, (3, n))
h = np.exp(rng.uniform(np.log(0.002), np.log(0.08), n))
w = rng.uniform(0.1, 1, n) * h ** 3
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
xd, yd, zd, hd, wd = (t(a) for a in (x, y, z, h, w))
lo, hi = shard_range(n, rank, world)
k = s.QuinticSplineKernel()
col = k.get_column_kernel_func(1000)

r2 = ops.Raster(300, 200, 0, 1, 0, 1)
part = ops.scatter2d(xd[lo:hi], yd[lo:hi], hd[lo:hi], [wd[lo:hi]], col, 3.0, 2, r2)[0]
combine_image(part, dst=0)
nz = 16 * world
r3 = ops.Raster(40, 36, 0, 1, 0, 1, nz, 0, 1)
c = s.CubicSplineKernel()
gpart = ops.scatter3d(xd[lo:hi], yd[lo:hi], zd[lo:hi], hd[lo:hi], [wd[lo:hi]], c.w, 2.0, r3)[0]
zl, zh = slab_range(nz, rank, world)
slab = torch.empty((zh - zl, 36, 40), dtype=torch.float64, device=dev)
combine_grid_along_z(gpart, slab)

full2 = ops.scatter2d(xd, yd, hd, [wd], col, 3.0, 2, r2)[0]
full3 = ops.scatter3d(xd, yd, zd, hd, [wd], c.w, 2.0, r3)[0]
e3 = float((slab - full3[zl:zh]).abs().max() / full3.abs().max())
errs = torch.tensor([e3], device=dev, dtype=torch.float64)
dist.all_reduce(errs, op=dist.ReduceOp.MAX)
if rank == 0:
    e2 = float((part - full2).abs().max() / full2.abs().max())
    from oracle import OracleBackend, OracleKernel
    ref = OracleBackend.interpolate_3d_projection(x, y, w, h, OracleKernel.column("quintic"), 3.0, 300, 200, 0, 1, 0, 1,
                                                  False)
    eo = float(np.abs(part.cpu().numpy() - ref).max() / np.abs(ref).max())
    print(f"world={world} image: sharded vs single {e2:.2e}, vs oracle {eo:.2e}; grid slabs vs single {errs.item():.2e}")
    assert e2 < 1e-12 and eo < 1e-10 and errs.item() < 1e-12
    print("multi-GPU check ok")
dist.destroy_process_group()
