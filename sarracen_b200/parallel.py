"""Multi-GPU plumbing: particle sharding and the one exchange step.

The scatter is a sum over particles, so it shards by particle with no
data-path collective until the end: rank r renders the
contiguous index range [r N / G, (r+1) N / G) -- the same partition the
reference's CPU backend gives its threads (cpu_backend.py:258-260) -- into a
full-size partial raster, then the partial rasters are summed over
NVLink/NVSwitch with NCCL: `reduce` to one rank for images, `reduce_scatter`
along z for voxel grids (rank r keeps planes [r nz / G, (r+1) nz / G)).
One process per GPU, `torch.distributed` for the plumbing.
"""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, exhaustive, non-overlapping particle range of `rank`."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return (n * rank) // world, (n * (rank + 1)) // world


def slab_range(nz: int, rank: int, world: int) -> Tuple[int, int]:
    """z-planes rank `rank` owns after the reduce-scatter (nz must divide evenly)."""
    if nz % world:
        raise ValueError(f"z_pixels={nz} is not divisible by the number of ranks {world}")
    per = nz // world
    return rank * per, (rank + 1) * per


def combine_image(img: torch.Tensor, dst: int = 0) -> torch.Tensor:
    """Sum the partial images onto rank `dst` (valid there only)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(img, dst=dst, op=dist.ReduceOp.SUM)
    return img


def combine_images_packed(imgs, dst: int = 0):
    """Several partial images (vec components, normalisation grid) in ONE collective."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return imgs
    packed = torch.stack(list(imgs))
    dist.reduce(packed, dst=dst, op=dist.ReduceOp.SUM)
    return [packed[i] for i in range(len(imgs))]


def combine_grid_along_z(grid: torch.Tensor, slab: torch.Tensor) -> torch.Tensor:
    """Reduce-scatter a (nz, ny, nx) partial grid along z; rank r receives the sum of
    planes slab_range(nz, r, G) in `slab`."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        slab.copy_(grid)
        return slab
    world = dist.get_world_size()
    lo, hi = slab_range(grid.shape[0], dist.get_rank(), world)
    if tuple(slab.shape) != (hi - lo,) + tuple(grid.shape[1:]):
        raise ValueError("slab has the wrong shape")
    if grid.is_cuda:
        dist.reduce_scatter_tensor(slab, grid, op=dist.ReduceOp.SUM)
    else:
        # gloo (CPU tests) has no reduce_scatter: same result through all_reduce
        tmp = grid.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM)
        slab.copy_(tmp[lo:hi])
    return slab
