"""Multi-GPU: particle sharding and the one exchange step, usable from the backend API.

The scatter is a sum over particles, so it shards by particle with no data-path collective
until the end: rank r renders the contiguous index range [r N / G, (r+1) N / G)
-- the partition the reference's CPU backend gives its threads (cpu_backend.py:258-260) -- into
a full-size partial raster, then the partial rasters are summed over NVLink / NVSwitch:

  images        one NCCL `reduce` to a root rank (or `all_reduce` when every rank wants the picture);
                the two components of a `_vec` call, or value + normalisation grid, travel packed
                in ONE collective;
  voxel grids   a reduce-scatter along z, cut into C z-chunks.  The render of a voxel grid is issued
                chunk by chunk (`ops.scatter3d(..., stage_planes=...)`: the sorted work lists are
                z-major, so chunk c is final as soon as stage c has run), and the exchange of chunk c
                runs while chunks c+1.. are still being rendered (`scatter3d_sharded`).  Rank r then
                owns planes [r b, (r+1) b) of every chunk (`owned_planes`: block-cyclic along z;
                C = 1 gives contiguous slabs).  Two transports: 'peer' (default) -- the copy engines
                push the partial planes into the owners' symmetric-memory slots and `sphb_sum_slots`
                adds them up, no SM taken from the render -- and 'nccl', one `reduce_scatter` per
                chunk on NCCL's stream.  `reduce_along_z` is the same exchange for a finished grid,
                `combine_grid_along_z` the monolithic NCCL form into contiguous slabs.

One process per GPU, `torch.distributed` for the plumbing.  Initialise the NCCL group with a
high-priority stream (`opts = dist.ProcessGroupNCCL.Options(); opts.is_high_priority_stream = True;
dist.init_process_group("nccl", pg_options=opts)`): with the 'nccl' transport the reduction of a
finished chunk only overlaps the render if its few CTAs get onto the SMs ahead of the queued render
CTAs.  `enable()` switches the nine `B200Backend` methods (and with them `backend='gpu'` of Sarracen's
own front-end) to this mode: every rank passes the SAME full columns, renders its shard and receives
the combined result.
"""
import os
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


# z-chunks of the overlapped exchange.  The stages of a staged render overlap each other's tails (two internal
# streams, sphb_scatter3d_staged) and the copy-engine transport does not compete with the render for SMs, so
# finer chunks only shorten the exposed tail: config 4 on 2 GPUs, render alone 20.14 ms, with the exchange
# 20.82 (4 chunks) / 20.65 ms (8 chunks); the exchange on its own takes 1.1 ms (profiles/diag_peer_overlap.py).
DEFAULT_CHUNKS = 8


def z_chunks(nz: int, world: int, chunks: Optional[int] = None) -> int:
    """Number of z-chunks the reduction along z is cut into: `chunks` (default DEFAULT_CHUNKS)
    reduced until every chunk splits evenly over the ranks; 0 if not even one chunk does (then
    the ranks own unequal contiguous slabs, see owned_planes)."""
    c = int(os.environ.get("SPHB_Z_CHUNKS") or DEFAULT_CHUNKS) if chunks is None else int(chunks)
    c = max(1, c)
    while c > 1 and nz % (c * world):
        c -= 1
    return c if nz % (c * world) == 0 else 0


def owned_planes(nz: int, rank: int, world: int, chunks: Optional[int] = None) -> List[int]:
    """z-planes of the summed grid that end up on `rank`, in the order they are returned.

    The reduction along z is a reduce-scatter PER Z-CHUNK (so that it can start as soon as that chunk
    is rendered): with C = z_chunks(...) chunks of nz / C planes each, rank r receives planes
    [r b, (r + 1) b) of every chunk, b = nz / (C G) -- a block-cyclic layout.  C = 1 is the plain
    contiguous slab; when nz does not divide evenly the ranks own the contiguous slab_range()."""
    c = z_chunks(nz, world, chunks)
    if c == 0:
        lo, hi = slab_range(nz, rank, world)
        return list(range(lo, hi))
    b = nz // (c * world)
    return [k * (nz // c) + rank * b + i for k in range(c) for i in range(b)]


# ---- exchange over peer memory -------------------------------------------------------------------
# NCCL's reduce-scatter is a kernel: next to a render that fills every SM it waits for SM slots and then
# holds them, so little of it hides behind the render (measured: 17-37 % at 2 and 8 GPUs).  The transport
# below moves the partial planes with the COPY ENGINES instead: every rank owns receive slots in symmetric
# (peer-mapped) memory (`torch.distributed._symmetric_memory`), the owner of a finished z-chunk's planes is
# sent each rank's partial planes by plain device-to-device copies over NVLink -- no SM involved -- and one
# small kernel of this library (`sphb_sum_slots`) adds the G - 1 received partials to the rank's own.  Two
# device-side barriers per exchange (slots free / copies landed) come from the symmetric-memory handle.
_peer_cache = {}
_peer_failed = False
last_transport = None   # 'peer' or 'nccl': what the most recent exchange along z used (reported by bench.py)


class _PeerSlots:
    def __init__(self, group, dev, numel):
        import torch.distributed._symmetric_memory as symm
        self.numel = numel
        self.buf = symm.empty(numel, dtype=torch.float64, device=dev)
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)


def _peer_slots(group, dev, numel) -> Optional[_PeerSlots]:
    """Symmetric receive slots of at least `numel` doubles on every rank (collective: all ranks of the
    group call it with the same size), or None where symmetric memory is not available."""
    global _peer_failed
    if _peer_failed:
        return None
    key = (id(group) if group is not None else 0, dev.index)
    slots = _peer_cache.get(key)
    if slots is None or slots.numel < numel:
        try:
            slots = _peer_cache[key] = _PeerSlots(group, dev, numel)
        except Exception as e:   # no P2P mapping on this machine / build: NCCL carries the exchange
            import warnings
            warnings.warn(f"symmetric memory unavailable ({e!r}); the z-exchange uses NCCL reduce_scatter")
            _peer_failed = True
            return None
    return slots


TRANSPORTS = ("peer", "nccl")


def _exchange_peer(peer, out, mine, rank, world, c, per, b, slot, events=None):
    """Peer-memory exchange of the partial grids `out` (one per weight) into `mine`; `events[k]` (if
    given) marks chunk k of `out` final on the current stream, otherwise the whole of `out` is final."""
    global last_transport
    last_transport = "peer"
    dev = out[0].device
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    hdl = peer.hdl
    if events is None:
        side.wait_stream(main)
    lanes = _copy_lanes(dev, min(PEER_COPY_LANES, world - 1))
    with torch.cuda.stream(side):
        # every rank is past the sums of the previous exchange: the slots may be overwritten
        hdl.barrier(channel=0)
        free = torch.cuda.Event()
        free.record(side)
        for k in range(c):
            if events is not None:
                side.wait_event(events[k])
            # the pushes to the G - 1 owners go out on several streams: one copy engine each
            for j, lane in enumerate(lanes):
                if k == 0:
                    lane.wait_event(free)
                if events is not None:
                    lane.wait_event(events[k])
                elif k == 0:
                    lane.wait_stream(main)
                with torch.cuda.stream(lane):
                    for step in range(1 + j, world, len(lanes)):
                        q = (rank + step) % world
                        for w, o in enumerate(out):
                            dst = hdl.get_buffer(q, (slot,), torch.float64, ((w * c + k) * world + rank) * slot)
                            dst.copy_(o[k * per + q * b:k * per + (q + 1) * b].reshape(-1), non_blocking=True)
                    sent = torch.cuda.Event()
                    sent.record(lane)
                side.wait_event(sent)
            hdl.barrier(channel=1)   # the copies of chunk k have landed on every rank
            for w, (o, m) in enumerate(zip(out, mine)):
                recv = peer.buf[(w * c + k) * world * slot:(w * c + k + 1) * world * slot].view(world, slot)
                _ops().sum_slots_(m[k * b:(k + 1) * b], o[k * per + rank * b:k * per + (rank + 1) * b], recv, rank)
    main.wait_stream(side)
    return mine


def _ops():
    from . import ops
    return ops


def scatter3d_sharded(x, y, z, h, weights, weight_function, radius, raster, out=None, group=None,
                      chunks: Optional[int] = None, overlap: bool = True,
                      transport: Optional[str] = None) -> List[torch.Tensor]:
    """Voxel grid of THIS rank's particle shard, summed over the ranks along z: returns, per weight,
    a (len(owned_planes), ny, nx) tensor holding the planes owned_planes(nz, rank, G, chunks).

    The render is issued in C z-chunk stages (`ops.scatter3d(stage_planes=...)`); as soon as chunk c
    is final its exchange starts on a side stream and overlaps the render of chunks c+1..; only the last
    chunk's exchange (1 / C of the volume) is exposed.  transport='peer' (default where symmetric memory
    works): copy-engine pushes into the owners' receive slots + `sphb_sum_slots`; 'nccl': one
    `reduce_scatter` per chunk on NCCL's stream.  `out` are full-size float64 work grids (allocated if
    omitted)."""
    from . import ops
    world, rank = _world(group), _rank(group)
    nz = raster.nz
    if transport is None:
        transport = os.environ.get("SPHB_Z_TRANSPORT") or ("peer" if (overlap and world > 1 and x.is_cuda) else "nccl")
    if transport not in TRANSPORTS:
        raise ValueError(f"transport must be one of {TRANSPORTS}")
    c = z_chunks(nz, world, chunks)
    if world == 1:
        return ops.scatter3d(x, y, z, h, weights, weight_function, radius, raster, out=out)
    if c == 0:
        out = ops.scatter3d(x, y, z, h, weights, weight_function, radius, raster, out=out)
        return [reduce_along_z(o, group, chunks) for o in out]
    per, b = nz // c, nz // (c * world)
    dev = x.device
    nw = len(weights)
    mine = [torch.empty((c * b,) + (raster.ny, raster.nx), dtype=torch.float64, device=dev) for _ in range(nw)]
    works = []
    slot = b * raster.ny * raster.nx                       # doubles one rank sends one owner per chunk and weight
    peer = _peer_slots(group, dev, nw * c * world * slot) if (transport == "peer" and overlap) else None
    if peer is not None:
        main = torch.cuda.current_stream(dev)
        events = [torch.cuda.Event() for _ in range(c)]
        done = set()

        def on_stage(k):
            # called by ops.scatter3d right after the kernels of stage k have been enqueued
            events[k].record(main)
            done.add(k)

        out = ops.scatter3d(x, y, z, h, weights, weight_function, radius, raster, out=out,
                            stage_planes=[(k + 1) * per for k in range(c)], on_stage=on_stage)
        for k in range(c):
            if k not in done:   # a render that could not be staged (split particle batches): final only now
                events[k].record(main)
        return _exchange_peer(peer, out, mine, rank, world, c, per, b, slot, events)
    global last_transport
    last_transport = "nccl"
    if not overlap or c == 1:
        out = ops.scatter3d(x, y, z, h, weights, weight_function, radius, raster, out=out)
        for k in range(c):
            for o, m in zip(out, mine):
                works.append(dist.reduce_scatter_tensor(m[k * b:(k + 1) * b], o[k * per:(k + 1) * per],
                                                        op=dist.ReduceOp.SUM, group=group, async_op=True))
    else:
        main = torch.cuda.current_stream(dev)
        events = [torch.cuda.Event() for _ in range(c)]
        done = []

        def on_stage(k):
            # called by ops.scatter3d right after the kernels of stage k have been enqueued
            events[k].record(main)
            done.append(k)

        out = ops.scatter3d(x, y, z, h, weights, weight_function, radius, raster, out=out,
                            stage_planes=[(k + 1) * per for k in range(c)], on_stage=on_stage)
        # The collectives are enqueued from a side stream that waits for stage k only (NCCL's stream
        # synchronises with the stream that is current at the time of the call, so from the main stream
        # it would wait for every stage).  The host is far ahead of the GPU here: all stages are queued.
        side = _side_stream(dev)
        with torch.cuda.stream(side):
            for k in sorted(done):
                side.wait_event(events[k])
                for o, m in zip(out, mine):
                    works.append(dist.reduce_scatter_tensor(m[k * b:(k + 1) * b], o[k * per:(k + 1) * per],
                                                            op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in works:
        w.wait()   # the current stream waits for the NCCL stream
    return mine


def reduce_along_z(grid: torch.Tensor, group=None, chunks: Optional[int] = None,
                   transport: Optional[str] = None) -> torch.Tensor:
    """Sum a FINISHED (nz, ny, nx) partial grid along z; returns the planes
    owned_planes(nz, rank, G, chunks) of the sum (same layout as scatter3d_sharded, no overlap)."""
    world, rank = _world(group), _rank(group)
    nz = grid.shape[0]
    if world == 1:
        return grid
    c = z_chunks(nz, world, chunks)
    transport = transport or os.environ.get("SPHB_Z_TRANSPORT")
    if c > 0 and grid.is_cuda and transport in (None, "peer") and grid.dtype == torch.float64 and grid.is_contiguous():
        per, b = nz // c, nz // (c * world)
        slot = b * grid.shape[1] * grid.shape[2]
        peer = _peer_slots(group, grid.device, c * world * slot)
        if peer is not None:
            mine = torch.empty((c * b,) + tuple(grid.shape[1:]), dtype=grid.dtype, device=grid.device)
            return _exchange_peer(peer, [grid], [mine], rank, world, c, per, b, slot)[0]
    if c == 0 or not grid.is_cuda:
        # unequal slabs (or gloo, which has no reduce_scatter): one in-place reduce per owner
        if c == 0:
            bounds = [[slab_range(nz, k, world)] for k in range(world)]
        else:
            per, b = nz // c, nz // (c * world)
            bounds = [[(j * per + k * b, j * per + (k + 1) * b) for j in range(c)] for k in range(world)]
        work = grid if grid.is_cuda else grid.clone()
        for k in range(world):
            for lo, hi in bounds[k]:
                if hi > lo:
                    dist.reduce(work[lo:hi], dst=_global(group, k), op=dist.ReduceOp.SUM, group=group)
        return torch.cat([work[lo:hi] for lo, hi in bounds[rank]]) if len(bounds[rank]) > 1 \
            else work[bounds[rank][0][0]:bounds[rank][0][1]]
    per, b = nz // c, nz // (c * world)
    mine = torch.empty((c * b,) + tuple(grid.shape[1:]), dtype=grid.dtype, device=grid.device)
    works = [dist.reduce_scatter_tensor(mine[k * b:(k + 1) * b], grid[k * per:(k + 1) * per], op=dist.ReduceOp.SUM,
                                        group=group, async_op=True) for k in range(c)]
    for w in works:
        w.wait()
    return mine


_side_streams = {}
_copy_lane_cache = {}
# streams the pushes of one chunk are spread over.  One: at 8 GPUs four lanes made the exchange on its own
# slower (2.07 -> 2.70 ms) and the overlapped step no faster (6.99 / 6.98 ms), profiles/exp_r2_transport8.sh
PEER_COPY_LANES = 1


def _copy_lanes(dev, n: int):
    lanes = _copy_lane_cache.setdefault(dev.index, [])
    while len(lanes) < n:
        lanes.append(torch.cuda.Stream(dev, priority=-1))
    return lanes[:n]


def _side_stream(dev) -> torch.cuda.Stream:
    s = _side_streams.get(dev.index)
    if s is None:
        s = _side_streams[dev.index] = torch.cuda.Stream(dev, priority=-1)
    return s


# ---- backend mode -----------------------------------------------------------------------------
class _Mode:
    def __init__(self, group, images, grids, chunks):
        self.group, self.images, self.grids, self.chunks = group, images, grids, chunks


_mode: Optional[_Mode] = None


def enable(group=None, images: str = "all", grids: str = "all", chunks: Optional[int] = None) -> None:
    """Make `B200Backend` multi-GPU: every rank of `group` (default: the world) calls the backend
    (or Sarracen's front-end with backend='gpu') with the same full columns; each renders its
    shard_range() of the particles and the partial rasters are combined as described above.

    images: 'all'  -- every rank receives the picture (all_reduce; the drop-in behaviour), or
            'root' -- only rank 0's return value is the sum (reduce), the others get zeros.
    grids:  'all'  -- every rank receives the whole (nz, ny, nx) grid (overlapped reduce-scatter along
                      z, then all_gather), or
            'slab' -- every rank receives only the planes owned_planes(nz, rank, G, chunks) (no
                      gather: the form `north_star` names; the return shape is (nz / G, ny, nx)).
    chunks: z-chunks of the overlapped exchange (default 8; 1 = contiguous slabs)."""
    global _mode
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("torch.distributed is not initialised")
    if images not in ("all", "root") or grids not in ("all", "slab"):
        raise ValueError("images must be 'all' or 'root', grids 'all' or 'slab'")
    _mode = _Mode(group, images, grids, chunks)


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


def gather_slabs(mine: torch.Tensor, nz: int, mode: _Mode) -> torch.Tensor:
    """Every rank's owned planes -> the whole (nz, ny, nx) grid on every rank."""
    world, rank = _world(mode.group), _rank(mode.group)
    c = z_chunks(nz, world, mode.chunks)
    tail = tuple(mine.shape[1:])
    if c > 0 and mine.is_cuda:
        b = nz // (c * world)
        gathered = torch.empty((world, c, b) + tail, dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(gathered.view((world * c * b,) + tail), mine.contiguous(), group=mode.group)
        # [rank][chunk][plane] -> z order [chunk][rank][plane]
        return gathered.permute(1, 0, 2, *range(3, 3 + len(tail))).reshape((nz,) + tail)
    full = torch.empty((nz,) + tail, dtype=mine.dtype, device=mine.device)
    for k in range(world):   # unequal slabs / gloo: one broadcast per owner
        planes = owned_planes(nz, k, world, mode.chunks)
        piece = mine.contiguous() if k == rank else torch.empty((len(planes),) + tail, dtype=mine.dtype,
                                                                 device=mine.device)
        dist.broadcast(piece, src=_global(mode.group, k), group=mode.group)
        full[planes] = piece
    return full
