"""Multi-GPU: particle sharding and the one exchange step, usable from the backend API.

The scatter is a sum over particles, so it shards by particle with no data-path collective
until the end: rank r renders the contiguous index range [r N / G, (r+1) N / G)
-- the partition the reference's CPU backend gives its threads (cpu_backend.py:258-260) -- into
a full-size partial raster, then the partial rasters are summed over NVLink / NVSwitch with NCCL:

  images        one `reduce` to a root rank (or `all_reduce` when every rank wants the picture);
                the two components of a `_vec` call, or value + normalisation grid, travel packed
                in ONE collective;
  voxel grids   reduce along z: rank k owns z-planes slab_range(nz, k, G).  The render of a voxel
                grid is issued slab by slab (`ops.scatter3d(..., stage_planes=...)`: the sorted
                work list is z-major, so slab k is final as soon as stage k has run), and the
                reduction of slab k onto rank k runs on NCCL's stream while slabs k+1.. are still
                being rendered (`scatter3d_sharded`).  The monolithic form, one
                `reduce_scatter_tensor` after the whole render, is `combine_grid_along_z`.

One process per GPU, `torch.distributed` for the plumbing.  `enable()` switches the nine
`B200Backend` methods (and with them `backend='gpu'` of Sarracen's own front-end) to this mode:
every rank passes the SAME full columns, renders its shard and receives the combined result.
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, exhaustive, non-overlapping particle range of `rank`."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return (n * rank) // world, (n * (rank + 1)) // world


def slab_range(nz: int, rank: int, world: int) -> Tuple[int, int]:
    """z-planes rank `rank` owns after the reduction along z (any nz: the first nz % world
    ranks hold one plane more)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(nz, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _rank(group=None) -> int:
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def _global(group, k: int) -> int:
    """Global rank of rank k of `group` (torch.distributed addresses dst / src by global rank)."""
    return k if group is None else dist.get_global_rank(group, k)


def combine_image(img: torch.Tensor, dst: Optional[int] = 0, group=None) -> torch.Tensor:
    """Sum the partial images onto rank `dst` (valid there only); dst=None: onto every rank."""
    if _world(group) > 1:
        if dst is None:
            dist.all_reduce(img, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.reduce(img, dst=_global(group, dst), op=dist.ReduceOp.SUM, group=group)
    return img


def combine_images_packed(imgs: Sequence[torch.Tensor], dst: Optional[int] = 0, group=None) -> List[torch.Tensor]:
    """Several partial images (vec components, normalisation grid) in ONE collective."""
    imgs = list(imgs)
    if _world(group) == 1:
        return imgs
    if len(imgs) == 1:
        return [combine_image(imgs[0], dst, group)]
    packed = torch.stack(imgs)
    combine_image(packed, dst, group)
    return [packed[i] for i in range(len(imgs))]


def combine_grid_along_z(grid: torch.Tensor, slab: torch.Tensor, group=None) -> torch.Tensor:
    """Monolithic form: reduce-scatter a (nz, ny, nx) partial grid along z; rank r receives the
    sum of planes slab_range(nz, r, G) in `slab`."""
    world = _world(group)
    if world == 1:
        slab.copy_(grid)
        return slab
    lo, hi = slab_range(grid.shape[0], _rank(group), world)
    if tuple(slab.shape) != (hi - lo,) + tuple(grid.shape[1:]):
        raise ValueError("slab has the wrong shape")
    if grid.is_cuda and grid.shape[0] % world == 0:
        dist.reduce_scatter_tensor(slab, grid, op=dist.ReduceOp.SUM, group=group)
    else:
        # gloo (CPU tests) has no reduce_scatter, and NCCL's needs equal slabs: one reduce per owner
        for k in range(world):
            klo, khi = slab_range(grid.shape[0], k, world)
            if khi > klo:
                part = grid[klo:khi].clone() if not grid.is_cuda else grid[klo:khi]
                dist.reduce(part, dst=_global(group, k), op=dist.ReduceOp.SUM, group=group)
                if k == _rank(group):
                    slab.copy_(part)
    return slab


def scatter3d_sharded(x, y, z, h, weights, weight_function, radius, raster, out=None, slabs=None,
                      group=None, overlap: bool = True) -> List[torch.Tensor]:
    """Voxel grid of THIS rank's particle shard, combined along z: returns, per weight, this rank's
    slab (planes slab_range(nz, rank, G)) of the sum over all ranks.

    The render is issued in G stages, stage k finishing z-slab k; the reduction of slab k onto
    its owner starts as soon as stage k is done and overlaps stages k+1.. (NCCL runs on its own
    stream).  `out` are full-size float64 work grids (allocated if omitted), `slabs` optional
    destination tensors; without them the returned slabs are views into `out`."""
    from . import ops
    world, rank = _world(group), _rank(group)
    nz = raster.nz
    bounds = [slab_range(nz, k, world) for k in range(world)]
    if world == 1 or not overlap:
        out = ops.scatter3d(x, y, z, h, weights, weight_function, radius, raster, out=out)
        works = []
        if world > 1:
            for k, (lo, hi) in enumerate(bounds):
                for o in out:
                    works.append(dist.reduce(o[lo:hi], dst=_global(group, k), op=dist.ReduceOp.SUM, group=group,
                                                 async_op=True))
    else:
        dev = x.device
        main = torch.cuda.current_stream(dev)
        events = [torch.cuda.Event() for _ in range(world)]
        works = []

        def after_stage(k):
            # called by ops.scatter3d right after the kernels of stage k have been enqueued
            events[k].record(main)

        # The reductions are enqueued from a side stream that waits for stage k only, so the NCCL
        # stream does not wait for the later stages (it synchronises with the stream current at the
        # time of the call).
        side = _side_stream(dev)
        stage_planes = [hi for (_lo, hi) in bounds]
        pending = []

        def on_stage(k):
            after_stage(k)
            pending.append(k)

        out = ops.scatter3d(x, y, z, h, weights, weight_function, radius, raster, out=out,
                            stage_planes=stage_planes, on_stage=on_stage)
        # NCCL calls are issued after the launches (host side is asynchronous: all stages are in the
        # stream queue by now and the events fire as the GPU gets there)
        with torch.cuda.stream(side):
            for k in pending:
                lo, hi = bounds[k]
                side.wait_event(events[k])
                for o in out:
                    if hi > lo:
                        works.append(dist.reduce(o[lo:hi], dst=_global(group, k), op=dist.ReduceOp.SUM, group=group,
                                                 async_op=True))
    for w in works:
        w.wait()   # the current stream waits for the NCCL stream
    lo, hi = bounds[rank]
    mine = [o[lo:hi] for o in out]
    if slabs is not None:
        for s, m in zip(slabs, mine):
            s.copy_(m)
        return list(slabs)
    return mine


def reduce_along_z(grid: torch.Tensor, group=None) -> torch.Tensor:
    """Sum a finished (nz, ny, nx) partial grid along z onto the slab owners (one reduce per slab, in
    place); returns this rank's slab as a view of `grid`."""
    world, rank = _world(group), _rank(group)
    bounds = [slab_range(grid.shape[0], k, world) for k in range(world)]
    if world > 1:
        works = [dist.reduce(grid[lo:hi], dst=_global(group, k), op=dist.ReduceOp.SUM, group=group, async_op=True)
                 for k, (lo, hi) in enumerate(bounds) if hi > lo]
        for w in works:
            w.wait()
    lo, hi = bounds[rank]
    return grid[lo:hi]


_side_streams = {}


def _side_stream(dev) -> torch.cuda.Stream:
    s = _side_streams.get(dev.index)
    if s is None:
        s = _side_streams[dev.index] = torch.cuda.Stream(dev)
    return s


# ---- backend mode -----------------------------------------------------------------------------
class _Mode:
    def __init__(self, group, images, grids):
        self.group, self.images, self.grids = group, images, grids


_mode: Optional[_Mode] = None


def enable(group=None, images: str = "all", grids: str = "all") -> None:
    """Make `B200Backend` multi-GPU: every rank of `group` (default: the world) calls the backend
    (or Sarracen's front-end with backend='gpu') with the same full columns; each renders its
    shard_range() of the particles and the partial rasters are combined as described above.

    images: 'all'  -- every rank receives the picture (all_reduce; the drop-in behaviour), or
            'root' -- only rank 0's return value is the sum (reduce), the others get zeros.
    grids:  'all'  -- every rank receives the whole (nz, ny, nx) grid (overlapped reduce along z,
                      then all_gather of the slabs), or
            'slab' -- every rank receives only its slab_range() of z-planes (no gather: the form
                      `north_star` names; the return shape is (nz_r, ny, nx))."""
    global _mode
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("torch.distributed is not initialised")
    if images not in ("all", "root") or grids not in ("all", "slab"):
        raise ValueError("images must be 'all' or 'root', grids 'all' or 'slab'")
    _mode = _Mode(group, images, grids)


def disable() -> None:
    global _mode
    _mode = None


def active() -> Optional[_Mode]:
    """The multi-GPU mode if enabled and the group has more than one rank."""
    if _mode is not None and _world(_mode.group) > 1:
        return _mode
    return None


def shard_columns(arrays, mode: _Mode):
    """This rank's slice of every column (views, no copy)."""
    n = len(arrays[0])
    lo, hi = shard_range(n, _rank(mode.group), _world(mode.group))
    return [a[lo:hi] for a in arrays]


def finish_images(imgs: Sequence[torch.Tensor], mode: _Mode) -> List[torch.Tensor]:
    out = combine_images_packed(imgs, None if mode.images == "all" else 0, mode.group)
    if mode.images == "root" and _rank(mode.group) != 0:
        out = [torch.zeros_like(o) for o in out]
    return out


def gather_slabs(slab: torch.Tensor, nz: int, mode: _Mode) -> torch.Tensor:
    """All ranks' slabs -> the whole grid on every rank."""
    world = _world(mode.group)
    full = torch.empty((nz,) + tuple(slab.shape[1:]), dtype=slab.dtype, device=slab.device)
    if nz % world == 0:
        dist.all_gather_into_tensor(full, slab.contiguous(), group=mode.group)
    else:
        parts = [full[slice(*slab_range(nz, k, world))] for k in range(world)]
        for k, p in enumerate(parts):   # unequal slabs: one broadcast per owner
            if k == _rank(mode.group):
                p.copy_(slab)
            dist.broadcast(p, src=_global(mode.group, k), group=mode.group)
    return full
