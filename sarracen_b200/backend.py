"""`B200Backend`: the drop-in for Sarracen's `GPUBackend`.

Same nine static methods, same argument order and meaning, same return types
as the reference's backend interface (sarracen/interpolate/base_backend.py:
7-177): NumPy arrays (or pandas Series) in, new float64 NumPy arrays out.
Internally every method moves its columns to the current CUDA device once,
calls ONE hand-written CUDA pipeline through the C ABI (sarracen_b200.ops) and
copies the image back.  The `_vec` methods accumulate both weights in a single
footprint walk instead of the reference's two passes (cpu_backend.py:58-63).

`install()` rebinds the name `GPUBackend` inside
`sarracen.interpolate.interpolate`, which `get_backend('gpu')` resolves at call
time (interpolate.py:1640-1658): every `backend='gpu'` call of
`interpolate_*`, `SarracenDataFrame.render` and `.sph_interpolate` then lands
here with no signature change.
"""
import warnings
from typing import Tuple

import numpy as np
import torch

from . import ops, parallel
from .kernels import WeightFunction

_RADIUS_KIND = {2.0: "cubic", 2.5: "quartic", 3.0: "quintic"}
_resolved = {}


def resolve_weight_function(weight_function, kernel_radius) -> WeightFunction:
    """Map what the front-end passes as `weight_function` to a device code path.

    Accepts a `WeightFunction` (this package's kernels), or one of Sarracen's
    numba dispatchers: the class-level `CubicSplineKernel.w` etc. (matched by
    qualified name), or the column-table closure built by
    `BaseKernel.get_column_kernel_func` (base_kernel.py:84-99), whose free
    variables are (column_kernel, radius, samples).  Anything else is
    rejected: arbitrary Python callables cannot run inside the CUDA kernels and
    there is no CPU fallback.
    """
    if isinstance(weight_function, WeightFunction):
        return weight_function
    if getattr(weight_function, "kind", None) in ("cubic", "quartic", "quintic", "lut"):
        return WeightFunction(weight_function.kind, getattr(weight_function, "radius", kernel_radius),
                              getattr(weight_function, "lut", None))
    key = id(weight_function)
    if key in _resolved and _resolved[key][0] is weight_function:
        return _resolved[key][1]
    py = getattr(weight_function, "py_func", None)
    if py is None:
        raise TypeError(f"cannot map weight_function {weight_function!r} to a device kernel")
    wf = None
    free = getattr(py.__code__, "co_freevars", ())
    if "column_kernel" in free:
        cells = dict(zip(free, (c.cell_contents for c in py.__closure__)))
        wf = WeightFunction("lut", float(cells["radius"]), np.asarray(cells["column_kernel"], dtype=np.float64))
    else:
        qual = getattr(py, "__qualname__", "")
        for name in ("Cubic", "Quartic", "Quintic"):
            if qual.startswith(name + "SplineKernel."):
                wf = WeightFunction(name.lower(), float(kernel_radius))
    if wf is None:
        raise TypeError(f"unsupported weight_function {weight_function!r}: only Sarracen's cubic, quartic "
                        "and quintic kernels and their column tables run on the B200 backend")
    _resolved[key] = (weight_function, wf)
    return wf


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("B200Backend needs a CUDA device; there is no CPU fallback "
                           "(use backend='cpu' for Sarracen's own CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def _common_dtype(arrays) -> torch.dtype:
    """float32 only if every column is float32; anything else computes in float64
    (the reference promotes mixed and integer columns to float64 as well)."""
    for a in arrays:
        dt = a.dtype if isinstance(a, torch.Tensor) else np.asarray(a).dtype
        if dt not in (np.float32, torch.float32):
            return torch.float64
    return torch.float32


def _from_numpy(arr: np.ndarray) -> torch.Tensor:
    """Host tensor sharing memory with `arr` where possible.  Read-only arrays (pandas hands them out
    under copy-on-write) are wrapped without the copy torch's warning suggests: the columns are only
    ever read."""
    arr = np.ascontiguousarray(arr)
    if arr.dtype.byteorder not in ("=", "|"):
        arr = arr.astype(arr.dtype.newbyteorder("="))
    if not arr.flags.writeable:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)
            return torch.from_numpy(arr)
    return torch.from_numpy(arr)


def _to_device(arrays, device):
    dtype = _common_dtype(arrays)
    out = []
    for a in arrays:
        if isinstance(a, torch.Tensor):
            t = a.to(device=device, dtype=dtype).contiguous().reshape(-1)
        else:
            t = _from_numpy(np.asarray(a)).to(device=device, dtype=dtype).reshape(-1)
        out.append(t)
    return out


# ---- streamed uploads -------------------------------------------------------------------------
# Host columns of a large particle set are not uploaded in one piece: they go up in chunks on a
# copy stream, double-buffered, while the scatter pipeline of the previous chunk runs; every chunk
# accumulates into the same raster.  At 2^27 particles the upload (5.4 GB, ~100 ms over PCIe)
# costs twice the kernels, so hiding one behind the other is worth more than any kernel tuning
# for a host-array caller.  Chunk sizes grow geometrically -- n/16 first (the only upload nothing
# hides), then 4x each up to 2^24 particles (2^24 Plummer particles: 3x -> 4 chunks 90.2 ms, 4x -> 3
# chunks 88.9 ms, one upload 96.5 ms) -- because a projection renders a few big chunks more
# efficiently than many small ones (fewer, longer tile segments), while an upload-bound voxel grid
# only needs the chunks to be short next to the whole.  Sets below 2^22 go up in one piece.
_STREAM_MIN = 1 << 22
_CHUNK_FIRST_MIN, _CHUNK_FIRST_MAX, _CHUNK_GROWTH, _CHUNK_MAX = 1 << 20, 1 << 23, 4, 1 << 24


def _chunk_bounds(n: int):
    size = min(max(n // 16, _CHUNK_FIRST_MIN), _CHUNK_FIRST_MAX)
    bounds, lo = [], 0
    while lo < n:
        hi = min(lo + size, n)
        bounds.append((lo, hi))
        lo, size = hi, min(size * _CHUNK_GROWTH, _CHUNK_MAX)
    return bounds


_copy_streams = {}

# Pageable host columns (plain NumPy arrays -- what Sarracen passes) copy at ~10 GB/s through the
# driver's own staging; here they are staged through small page-locked slots instead, one worker
# thread per column (the host memcpy releases the GIL), each slot sent on as soon as it is filled.
_STAGE_ELEMS, _STAGE_SLOTS = 1 << 21, 3
_stage_pool = None


class _StageRing:
    """Page-locked staging slots of one worker (one column at a time), reused round-robin."""

    def __init__(self):
        self.slots = {}   # dtype -> [[buffer, event or None], ...]
        self.turn = {}

    def next(self, dtype):
        ring = self.slots.get(dtype)
        if ring is None:
            ring = self.slots[dtype] = [[torch.empty(_STAGE_ELEMS, dtype=dtype, pin_memory=True), None]
                                        for _ in range(_STAGE_SLOTS)]
            self.turn[dtype] = 0
        slot = ring[self.turn[dtype]]
        self.turn[dtype] = (self.turn[dtype] + 1) % _STAGE_SLOTS
        if slot[1] is not None:
            slot[1].synchronize()   # the copy that last read this slot has finished
        return slot


_stage_rings = {}


def release_host_buffers() -> None:
    """Free the page-locked staging slots (up to 3 x 16 MB per column worker) kept between calls."""
    _stage_rings.clear()


def _stage_column(j, src, dst, copy_stream):
    """Worker: pageable `src` (host) -> `dst` (device) through the calling worker thread's own
    page-locked slots (`j` is only a label: the ring belongs to the thread, so any number of
    columns can be in flight on the pool)."""
    import threading
    key = threading.get_ident()
    ring = _stage_rings.get(key)
    if ring is None:
        ring = _stage_rings[key] = _StageRing()
    n = src.numel()
    with torch.cuda.stream(copy_stream):
        for off in range(0, n, _STAGE_ELEMS):
            m = min(_STAGE_ELEMS, n - off)
            slot = ring.next(src.dtype)
            slot[0][:m].copy_(src[off:off + m])
            dst[off:off + m].copy_(slot[0][:m], non_blocking=True)
            slot[1] = torch.cuda.Event()
            slot[1].record(copy_stream)


def _host_columns(arrays):
    """1-D host tensors sharing memory with the inputs, or None if any column is not host data."""
    cols = []
    for a in arrays:
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return None
            t = a.detach().contiguous().reshape(-1)
        else:
            arr = np.asarray(a)
            if arr.dtype.kind not in "fiu":
                return None
            t = _from_numpy(arr).reshape(-1)
        cols.append(t)
    n = cols[0].numel()
    return cols if all(c.numel() == n for c in cols) else None


def streamed_scatter(arrays, dev, render):
    """Run `render(device_columns, out, accumulate) -> [rasters]` over a particle set given as host
    (or device) columns; returns the device rasters.  `device_columns` are 1-D tensors of one
    dtype, in the order of `arrays`.

    Small or device-resident sets: one call.  Large host sets: chunked as described above.  Public
    because multi-GPU callers use it per rank on their shard before the NCCL combine
    (INTEGRATION.md section 4)."""
    cols = _host_columns(arrays)
    n = cols[0].numel() if cols else 0
    if cols is None or n < _STREAM_MIN:
        return render(_to_device(arrays, dev), None, False)
    dtype = _common_dtype(cols)
    bounds = _chunk_bounds(n)
    largest = max(hi - lo for lo, hi in bounds)
    copy_stream = _copy_streams.get(dev.index)
    if copy_stream is None:
        copy_stream = _copy_streams[dev.index] = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    bufs = [[torch.empty(largest, dtype=c.dtype, device=dev) for c in cols] for _ in range(2)]
    ready, free = [None, None], [None, None]

    pageable = [j for j, c in enumerate(cols) if not c.is_pinned()]
    global _stage_pool
    if pageable and _stage_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _stage_pool = ThreadPoolExecutor(max_workers=8, thread_name_prefix="sphb-stage")

    def upload(k):
        b, (lo, hi) = k % 2, bounds[k]
        with torch.cuda.stream(copy_stream):
            if free[b] is not None:
                copy_stream.wait_event(free[b])
            jobs = []
            for j, (buf, c) in enumerate(zip(bufs[b], cols)):
                if c.is_pinned():
                    buf[:hi - lo].copy_(c[lo:hi], non_blocking=True)
                else:
                    jobs.append(_stage_pool.submit(_stage_column, j % 8, c[lo:hi], buf[:hi - lo], copy_stream))
            for job in jobs:
                job.result()
            ready[b] = torch.cuda.Event()
            ready[b].record(copy_stream)

    # the buffers come from the main stream's caching allocator: whatever used their blocks last (the
    # previous call's kernels, other work of the caller) was ordered on the main stream, so the copies wait
    copy_stream.wait_stream(main)
    upload(0)
    out = None
    for k, (lo, hi) in enumerate(bounds):
        b = k % 2
        main.wait_event(ready[b])
        views = [buf[:hi - lo] if buf.dtype == dtype else buf[:hi - lo].to(dtype) for buf in bufs[b]]
        out = render(views, out, k > 0)
        free[b] = torch.cuda.Event()
        free[b].record(main)
        # after the launch: the host side of the next upload (staging copies of pageable columns)
        # then overlaps this chunk's kernels
        if k + 1 < len(bounds):
            upload(k + 1)
    # the buffers go back to the caching allocator of the main stream: make it wait for the copies
    main.wait_stream(copy_stream)
    return out


_PINNED_MIN, _PINNED_MAX = 8 << 20, 4 << 30


def _host(t: torch.Tensor) -> np.ndarray:
    """Device raster -> new host ndarray.  Large rasters land in page-locked memory from torch's
    caching host allocator (a plain ndarray to the caller; the block is recycled when the array
    is released): the copy then runs at PCIe speed instead of through pageable bounce buffers."""
    nbytes = t.numel() * t.element_size()
    if _PINNED_MIN <= nbytes <= _PINNED_MAX:
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t)
        return host.numpy()
    return t.cpu().numpy()


def _shard(arrays):
    """Multi-GPU mode (sarracen_b200.parallel.enable): this rank's slice of the columns."""
    mode = parallel.active()
    if mode is None:
        return list(arrays), None
    return parallel.shard_columns([a if isinstance(a, torch.Tensor) else np.asarray(a) for a in arrays], mode), mode


def _images(rasters, mode):
    """Device rasters of one call -> host arrays; in multi-GPU mode summed over the ranks first
    (all components packed into one collective)."""
    rasters = list(rasters)
    if mode is not None:
        rasters = parallel.finish_images(rasters, mode)
    return [_host(r) for r in rasters]


class B200Backend:
    """SPH interpolation on a B200 behind Sarracen's backend interface.

    With `sarracen_b200.parallel.enable()` (one process per GPU under torch.distributed) every
    method renders this rank's shard of the particles and combines the partial rasters over NCCL:
    all ranks pass the same columns and receive the combined result."""

    @staticmethod
    def interpolate_2d_render(x, y, weight, h, weight_function, kernel_radius, x_pixels, y_pixels,
                              x_min, x_max, y_min, y_max, exact) -> np.ndarray:
        dev = _device()
        r = ops.Raster(x_pixels, y_pixels, x_min, x_max, y_min, y_max)
        (x, y, weight, h), mode = _shard([x, y, weight, h])
        if exact:
            return _images(streamed_scatter([x, y, weight, h], dev, lambda c, out, acc: ops.exact2d(
                c[0], c[1], c[3], [c[2]], r, out=out, accumulate=acc)), mode)[0]
        wf = resolve_weight_function(weight_function, kernel_radius)
        return _images(streamed_scatter([x, y, weight, h], dev, lambda c, out, acc: ops.scatter2d(
            c[0], c[1], c[3], [c[2]], wf, kernel_radius, 2, r, out=out, accumulate=acc)), mode)[0]

    @staticmethod
    def interpolate_2d_render_vec(x, y, weight_x, weight_y, h, weight_function, kernel_radius, x_pixels,
                                  y_pixels, x_min, x_max, y_min, y_max,
                                  exact) -> Tuple[np.ndarray, np.ndarray]:
        dev = _device()
        r = ops.Raster(x_pixels, y_pixels, x_min, x_max, y_min, y_max)
        (x, y, weight_x, weight_y, h), mode = _shard([x, y, weight_x, weight_y, h])
        if exact:
            a, b = _images(streamed_scatter([x, y, weight_x, weight_y, h], dev, lambda c, out, acc: ops.exact2d(
                c[0], c[1], c[4], [c[2], c[3]], r, out=out, accumulate=acc)), mode)
        else:
            wf = resolve_weight_function(weight_function, kernel_radius)
            a, b = _images(streamed_scatter([x, y, weight_x, weight_y, h], dev, lambda c, out, acc: ops.scatter2d(
                c[0], c[1], c[4], [c[2], c[3]], wf, kernel_radius, 2, r, out=out, accumulate=acc)), mode)
        return a, b

    @staticmethod
    def interpolate_2d_line(x, y, weight, h, weight_function, kernel_radius, pixels, x1, x2, y1,
                            y2) -> np.ndarray:
        dev = _device()
        (x, y, weight, h), mode = _shard([x, y, weight, h])
        xd, yd, wd, hd = _to_device([x, y, weight, h], dev)
        wf = resolve_weight_function(weight_function, kernel_radius)
        return _images(ops.line2d(xd, yd, hd, [wd], wf, kernel_radius, pixels, x1, x2, y1, y2), mode)[0]

    @staticmethod
    def interpolate_3d_line(x, y, z, weight, h, weight_function, kernel_radius, pixels, x1, x2, y1, y2,
                            z1, z2) -> np.ndarray:
        dev = _device()
        (x, y, z, weight, h), mode = _shard([x, y, z, weight, h])
        xd, yd, zd, wd, hd = _to_device([x, y, z, weight, h], dev)
        wf = resolve_weight_function(weight_function, kernel_radius)
        return _images(ops.line3d(xd, yd, zd, hd, [wd], wf, kernel_radius, pixels, x1, x2, y1, y2, z1, z2), mode)[0]

    @staticmethod
    def interpolate_3d_projection(x, y, weight, h, weight_function, kernel_radius, x_pixels, y_pixels,
                                  x_min, x_max, y_min, y_max, exact) -> np.ndarray:
        dev = _device()
        r = ops.Raster(x_pixels, y_pixels, x_min, x_max, y_min, y_max)
        (x, y, weight, h), mode = _shard([x, y, weight, h])
        if exact:
            return _images(streamed_scatter([x, y, weight, h], dev, lambda c, out, acc: ops.exact3d_proj(
                c[0], c[1], c[3], [c[2]], r, out=out, accumulate=acc)), mode)[0]
        wf = resolve_weight_function(weight_function, kernel_radius)
        return _images(streamed_scatter([x, y, weight, h], dev, lambda c, out, acc: ops.scatter2d(
            c[0], c[1], c[3], [c[2]], wf, kernel_radius, 2, r, out=out, accumulate=acc)), mode)[0]

    @staticmethod
    def interpolate_3d_projection_vec(x, y, weight_x, weight_y, h, weight_function, kernel_radius,
                                      x_pixels, y_pixels, x_min, x_max, y_min, y_max,
                                      exact) -> Tuple[np.ndarray, np.ndarray]:
        dev = _device()
        r = ops.Raster(x_pixels, y_pixels, x_min, x_max, y_min, y_max)
        (x, y, weight_x, weight_y, h), mode = _shard([x, y, weight_x, weight_y, h])
        if exact:
            a, b = _images(streamed_scatter([x, y, weight_x, weight_y, h], dev, lambda c, out, acc: ops.exact3d_proj(
                c[0], c[1], c[4], [c[2], c[3]], r, out=out, accumulate=acc)), mode)
        else:
            wf = resolve_weight_function(weight_function, kernel_radius)
            a, b = _images(streamed_scatter([x, y, weight_x, weight_y, h], dev, lambda c, out, acc: ops.scatter2d(
                c[0], c[1], c[4], [c[2], c[3]], wf, kernel_radius, 2, r, out=out, accumulate=acc)), mode)
        return a, b

    @staticmethod
    def interpolate_3d_cross(x, y, z, z_slice, weight, h, weight_function, kernel_radius, x_pixels,
                             y_pixels, x_min, x_max, y_min, y_max) -> np.ndarray:
        dev = _device()
        r = ops.Raster(x_pixels, y_pixels, x_min, x_max, y_min, y_max)
        wf = resolve_weight_function(weight_function, kernel_radius)
        zs = float(z_slice)
        (x, y, z, weight, h), mode = _shard([x, y, z, weight, h])
        return _images(streamed_scatter([x, y, z, weight, h], dev, lambda c, out, acc: ops.scatter2d(
            c[0], c[1], c[4], [c[3]], wf, kernel_radius, 3, r, z=c[2], z_slice=zs, out=out, accumulate=acc)),
            mode)[0]

    @staticmethod
    def interpolate_3d_cross_vec(x, y, z, z_slice, weight_x, weight_y, h, weight_function, kernel_radius,
                                 x_pixels, y_pixels, x_min, x_max, y_min,
                                 y_max) -> Tuple[np.ndarray, np.ndarray]:
        dev = _device()
        r = ops.Raster(x_pixels, y_pixels, x_min, x_max, y_min, y_max)
        wf = resolve_weight_function(weight_function, kernel_radius)
        zs = float(z_slice)
        (x, y, z, weight_x, weight_y, h), mode = _shard([x, y, z, weight_x, weight_y, h])
        a, b = _images(streamed_scatter([x, y, z, weight_x, weight_y, h], dev, lambda c, out, acc: ops.scatter2d(
            c[0], c[1], c[5], [c[3], c[4]], wf, kernel_radius, 3, r, z=c[2], z_slice=zs, out=out, accumulate=acc)),
            mode)
        return a, b

    @staticmethod
    def interpolate_3d_grid(x, y, z, weight, h, weight_function, kernel_radius, x_pixels, y_pixels,
                            z_pixels, x_min, x_max, y_min, y_max, z_min, z_max) -> np.ndarray:
        dev = _device()
        r = ops.Raster(x_pixels, y_pixels, x_min, x_max, y_min, y_max, z_pixels, z_min, z_max)
        wf = resolve_weight_function(weight_function, kernel_radius)
        (x, y, z, weight, h), mode = _shard([x, y, z, weight, h])
        if mode is None:
            return _host(streamed_scatter([x, y, z, weight, h], dev, lambda c, out, acc: ops.scatter3d(
                c[0], c[1], c[2], c[4], [c[3]], wf, kernel_radius, r, out=out, accumulate=acc))[0])
        # multi-GPU: this rank's shard -> partial grid, summed along z onto the slab owners.  A shard that
        # goes up in one piece is rendered slab by slab with the reduction of each finished slab overlapping
        # the render of the next ones; a shard streamed in chunks adds to every slab until its last chunk,
        # so its reduction starts when the render is over.
        cols = _host_columns([x, y, z, weight, h])
        if cols is None or cols[0].numel() < _STREAM_MIN:
            c = _to_device([x, y, z, weight, h], dev)
            slab = parallel.scatter3d_sharded(c[0], c[1], c[2], c[4], [c[3]], wf, kernel_radius, r,
                                              group=mode.group, chunks=mode.chunks)[0]
        else:
            part = streamed_scatter([x, y, z, weight, h], dev, lambda c, out, acc: ops.scatter3d(
                c[0], c[1], c[2], c[4], [c[3]], wf, kernel_radius, r, out=out, accumulate=acc))[0]
            slab = parallel.reduce_along_z(part, mode.group, mode.chunks)
        if mode.grids == "slab":
            return _host(slab)
        return _host(parallel.gather_slabs(slab, int(z_pixels), mode))


def install() -> type:
    """Route Sarracen's backend='gpu' to `B200Backend` (see module docstring).  The class installed is
    `B200Backend` re-based on Sarracen's own `BaseBackend` (base_backend.py:7), so `isinstance` /
    `issubclass` checks against the reference interface hold; it is returned."""
    import sarracen.interpolate
    import sarracen.interpolate.interpolate as front
    from sarracen.interpolate.base_backend import BaseBackend
    cls = type("B200Backend", (B200Backend, BaseBackend), {"__doc__": B200Backend.__doc__})
    front.GPUBackend = cls
    sarracen.interpolate.GPUBackend = cls
    return cls
