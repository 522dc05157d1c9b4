"""Device-level operators: torch CUDA tensors in, torch CUDA tensors out.

Thin host code over the C ABI of libsphb200.so (include/sphb200.h).  PyTorch is
used for device memory, streams and nothing else: every function here hands raw
device pointers to a hand-written CUDA kernel.  There is no CPU path -- a
missing library or a non-CUDA tensor raises.
"""
import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import SphbKernel, SphbParticles, SphbRaster, SphbStats

_KIND = {"cubic": _lib.SPHB_CUBIC, "quartic": _lib.SPHB_QUARTIC, "quintic": _lib.SPHB_QUINTIC,
         "lut": _lib.SPHB_COLUMN_LUT}

# Workspace handling: one growing byte buffer per device.  A call whose bin
# list would not fit in `max_workspace_fraction` of the free memory is split
# into particle batches that accumulate into the same output.  The buffer is
# shared by every call on that device: like the reference's backends, the
# operators are synchronous and meant to be called from one host thread per
# device (one process per GPU under torch.distributed).
_workspaces = {}
max_workspace_fraction = 0.6
last_stats = {}
time_phases = False  # bench/profiling: bracket the pipeline phases with CUDA events (adds a sync)


@dataclass
class Raster:
    nx: int
    ny: int
    x_min: float
    x_max: float
    y_min: float
    y_max: float
    nz: int = 1
    z_min: float = 0.0
    z_max: float = 1.0

    def c(self) -> SphbRaster:
        return SphbRaster(int(self.nx), int(self.ny), int(self.nz), float(self.x_min), float(self.x_max),
                          float(self.y_min), float(self.y_max), float(self.z_min), float(self.z_max))


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (sarracen_b200 has no CPU path)")


def _workspace(device, nbytes: int) -> torch.Tensor:
    key = str(device)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            del _workspaces[key]
            del ws
        ws = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def release_workspaces() -> None:
    _workspaces.clear()


def _particles(x, y, z, h, weights: Sequence[torch.Tensor], lo: int = 0, hi: Optional[int] = None):
    """Pack column tensors (one dtype, contiguous, same device) into the C struct."""
    cols = [x, y, h] + list(weights) + ([z] if z is not None else [])
    for i, t in enumerate(cols):
        _require_cuda(t, f"particle column {i}")
    dtype, device, n = x.dtype, x.device, x.numel()
    if dtype not in (torch.float32, torch.float64):
        raise TypeError("particle columns must be float32 or float64")
    for t in cols:
        if t.dtype != dtype or t.device != device or t.numel() != n or not t.is_contiguous() or t.dim() != 1:
            raise ValueError("particle columns must be contiguous 1-D tensors of one dtype, length and device")
    if not 1 <= len(weights) <= _lib.SPHB_MAX_WEIGHTS:
        raise ValueError(f"1..{_lib.SPHB_MAX_WEIGHTS} weight columns are supported per call")
    hi = n if hi is None else hi
    item = x.element_size()
    p = SphbParticles()
    p.n = hi - lo
    p.dtype = _lib.SPHB_F64 if dtype == torch.float64 else _lib.SPHB_F32
    p.x = x.data_ptr() + lo * item
    p.y = y.data_ptr() + lo * item
    p.z = (z.data_ptr() + lo * item) if z is not None else None
    p.h = h.data_ptr() + lo * item
    p.n_weights = len(weights)
    for i, w in enumerate(weights):
        p.w[i] = w.data_ptr() + lo * item
    return p


def _kernel(weight_function, radius: float, ndim: int, device):
    kind = getattr(weight_function, "kind", None)
    if kind not in _KIND:
        # one of Sarracen's own numba dispatchers (kernel.w, or the column-table closure of
        # BaseKernel.get_column_kernel_func): map it to a device code path, or raise TypeError
        from .backend import resolve_weight_function
        weight_function = resolve_weight_function(weight_function, radius)
        kind = weight_function.kind
    k = SphbKernel()
    k.kind = _KIND[kind]
    k.ndim = int(ndim)
    k.radius = float(radius)
    keep = None
    if kind == "lut":
        keep = weight_function.lut_on(device)
        k.lut = keep.data_ptr()
        k.lut_samples = keep.numel()
        k.ndim = 2
    return k, keep


def _out_ptrs(outs: Sequence[torch.Tensor]):
    arr = (ctypes.c_void_p * len(outs))()
    for i, o in enumerate(outs):
        arr[i] = o.data_ptr()
    return arr


def column_kernel(kind: str, samples: int, device=None) -> torch.Tensor:
    """A4: column-integrated kernel table on the device (base_kernel.py:39-66)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty(int(samples), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.load().sphb_column_kernel(_KIND[kind], int(samples), ctypes.c_void_p(out.data_ptr()),
                                                  _stream(device)))
    return out


def _scatter(dim3: bool, x, y, z, h, weights, weight_function, radius, ndim, raster: Raster, use_z=False,
             z_slice=0.0, out: Optional[List[torch.Tensor]] = None, accumulate=False, stage_planes=None,
             on_stage=None) -> List[torch.Tensor]:
    lib = _lib.load()
    device = x.device
    n = x.numel()
    shape = (raster.nz, raster.ny, raster.nx) if dim3 else (raster.ny, raster.nx)
    if out is None:
        out = [torch.empty(shape, dtype=torch.float64, device=device) for _ in weights]
        accumulate = False
    for o in out:
        _require_cuda(o, "output")
        if o.dtype != torch.float64 or tuple(o.shape) != shape or not o.is_contiguous():
            raise ValueError(f"outputs must be contiguous float64 tensors of shape {shape}")
    kern, _keep = _kernel(weight_function, radius, ndim, device)
    rast = raster.c()
    optr = _out_ptrs(out)
    stats = SphbStats()
    stats.time_phases = 1 if time_phases else 0
    total_pairs = 0
    total_direct = 0
    ms = [0.0, 0.0, 0.0, 0.0, 0.0]
    launches = 0
    # z-slab stages (voxel grids only): the library calls back right after the kernels of stage s have
    # been enqueued, and `on_stage(s)` runs there, on this thread, so that an event recorded by it sits
    # between stage s and stage s + 1 in the stream.  A particle set that has to be split into batches
    # (workspace) cannot be staged: every batch adds to all slabs, so the stages are reported at the end.
    staged = dim3 and stage_planes is not None and len(stage_planes) > 1
    split = False
    if staged:
        n_stages = len(stage_planes)
        planes_c = (ctypes.c_int * n_stages)(*[int(v) for v in stage_planes])
        stage_cb = _lib.STAGE_FN((lambda _user, s: on_stage(s)) if on_stage is not None else (lambda _user, s: None))
    with torch.cuda.device(device):
        stream = _stream(device)
        budget = None  # worked out only if the workspace has to grow
        todo = [(0, n)]
        done_any = False
        while todo:
            lo, hi = todo.pop()
            part = _particles(x, y, z, h, weights, lo, hi)
            # the first batch clears the outputs (inside the library), later ones add to them
            acc = 1 if (accumulate or done_any) else 0
            need = ctypes.c_size_t(0)
            ws = _workspaces.get(str(device))
            while True:
                wptr = ctypes.c_void_p(ws.data_ptr()) if ws is not None else None
                wlen = ws.numel() if ws is not None else 0
                if staged and not split:
                    rc = lib.sphb_scatter3d_staged(ctypes.byref(part), ctypes.byref(kern), ctypes.byref(rast), acc,
                                                   optr, wptr, wlen, ctypes.byref(need), ctypes.byref(stats),
                                                   n_stages, planes_c, stage_cb, None, stream)
                elif dim3:
                    rc = lib.sphb_scatter3d(ctypes.byref(part), ctypes.byref(kern), ctypes.byref(rast), acc, optr,
                                            wptr, wlen, ctypes.byref(need), ctypes.byref(stats), stream)
                else:
                    rc = lib.sphb_scatter2d(ctypes.byref(part), ctypes.byref(kern), ctypes.byref(rast),
                                            1 if use_z else 0, float(z_slice), acc, optr, wptr, wlen,
                                            ctypes.byref(need), ctypes.byref(stats), stream)
                if rc == _lib.SPHB_ERR_WORKSPACE and budget is None:
                    free, _total = torch.cuda.mem_get_info(device)
                    held = _workspaces[str(device)].numel() if str(device) in _workspaces else 0
                    budget = int((free + held) * max_workspace_fraction)
                if rc == _lib.SPHB_ERR_WORKSPACE and need.value <= budget:
                    ws = None  # drop our reference so the old buffer is freed before growing
                    ws = _workspace(device, need.value)
                    continue
                break
            if rc in (_lib.SPHB_ERR_WORKSPACE, _lib.SPHB_ERR_TOO_MANY) and hi - lo > 1:
                mid = (lo + hi) // 2  # bin list too large for this device: halve the batch
                todo.append((mid, hi))
                todo.append((lo, mid))
                split = True
                continue
            _lib.check(rc)
            done_any = True
            total_pairs += stats.n_pairs
            total_direct += stats.n_direct
            launches += stats.launches
            ms = [ms[0] + stats.ms_bin, ms[1] + stats.ms_sort, ms[2] + stats.ms_render, ms[3] + stats.ms_direct,
                  ms[4] + stats.ms_count]
        if staged and split and on_stage is not None:
            for s_done in range(n_stages):
                on_stage(s_done)
        if not staged and on_stage is not None and stage_planes is not None:
            for s_done in range(len(stage_planes)):   # a single stage: the whole call
                on_stage(s_done)
    last_stats.clear()
    last_stats.update(n_pairs=int(total_pairs), n_direct=int(total_direct), n_tiles=int(stats.n_tiles),
                      tile=(int(stats.tile_x), int(stats.tile_y), int(stats.tile_z)),
                      ms_bin=ms[0], ms_sort=ms[1], ms_render=ms[2], ms_direct=ms[3], ms_count=ms[4],
                      launches=launches)
    return out


class Plan:
    """Bin totals of one particle set on one raster (`sphb_plan`), in page-locked host memory."""

    def __init__(self):
        self.buf = torch.zeros(8, dtype=torch.int64).pin_memory()
        self.c = _lib.SphbPlan.from_address(self.buf.data_ptr())

    def __repr__(self):
        return (f"Plan(n_small={self.c.n_small}, n_pairs={self.c.n_pairs}, n_large={self.c.n_large}, "
                f"max_small_edge={self.c.max_small_edge})")


def _plan_args(dim3, x, y, z, h, weights, weight_function, radius, ndim, raster, z_slice):
    device = x.device
    use_z = (not dim3) and z is not None and z_slice is not None
    kern, keep = _kernel(weight_function, radius, 3 if dim3 else ndim, device)
    part = _particles(x, y, z if (dim3 or use_z) else None, h, weights)
    return device, use_z, kern, keep, part, raster.c()


def scatter_plan(x, y, h, weights, weight_function, radius, ndim, raster: Raster, z=None, z_slice=None,
                 plan: Optional[Plan] = None) -> Plan:
    """Phase 1 of the plan / execute form (`sphb_scatter_plan`): the count pass, asynchronous; the totals
    land in `plan` once the stream has got there (synchronise before reading or using it).  A voxel
    raster (raster.nz > 1) plans the grid scatter."""
    dim3 = raster.nz > 1
    device, use_z, kern, _keep, part, rast = _plan_args(dim3, x, y, z, h, weights, weight_function, radius, ndim,
                                                        raster, z_slice)
    plan = plan or Plan()
    lib = _lib.load()
    need = ctypes.c_size_t(0)
    with torch.cuda.device(device):
        for _attempt in range(2):
            ws = _workspaces.get(str(device))
            rc = lib.sphb_scatter_plan(ctypes.byref(part), ctypes.byref(kern), ctypes.byref(rast), 1 if dim3 else 0,
                                       1 if use_z else 0, float(z_slice or 0.0),
                                       ctypes.c_void_p(ws.data_ptr()) if ws is not None else None,
                                       ws.numel() if ws is not None else 0, ctypes.byref(need),
                                       ctypes.c_void_p(plan.buf.data_ptr()), _stream(device))
            if rc != _lib.SPHB_ERR_WORKSPACE:
                break
            ws = None
            _workspace(device, need.value)
        _lib.check(rc)
    return plan


def scatter_planned(plan: Plan, x, y, h, weights, weight_function, radius, ndim, raster: Raster, out, z=None,
                    z_slice=None, accumulate=False) -> List[torch.Tensor]:
    """Phase 2 (`sphb_scatter_planned`): the whole pipeline with every list sized from `plan` -- enqueues
    kernels and returns, no synchronisation, CUDA-graph capturable once the workspace is large enough
    (`scatter_planned_workspace` grows it outside the capture).  `out`: float64 rasters to fill."""
    dim3 = raster.nz > 1
    device, use_z, kern, _keep, part, rast = _plan_args(dim3, x, y, z, h, weights, weight_function, radius, ndim,
                                                        raster, z_slice)
    ws = _workspaces.get(str(device))
    need = ctypes.c_size_t(0)
    with torch.cuda.device(device):
        rc = _lib.load().sphb_scatter_planned(
            ctypes.byref(part), ctypes.byref(kern), ctypes.byref(rast), 1 if dim3 else 0, 1 if use_z else 0,
            float(z_slice or 0.0), 1 if accumulate else 0, _out_ptrs(out),
            ctypes.c_void_p(ws.data_ptr()) if ws is not None else None, ws.numel() if ws is not None else 0,
            ctypes.byref(need), ctypes.byref(plan.c), _stream(device))
    if rc == _lib.SPHB_ERR_WORKSPACE:
        raise _lib.SphbError(rc, f"workspace of {need.value} bytes needed: call scatter_planned_workspace first")
    _lib.check(rc)
    return out


def scatter_planned_workspace(plan: Plan, x, y, h, weights, weight_function, radius, ndim, raster: Raster, z=None,
                              z_slice=None) -> int:
    """Grow the device workspace to what `scatter_planned` needs for `plan`; returns the byte count."""
    dim3 = raster.nz > 1
    device, use_z, kern, _keep, part, rast = _plan_args(dim3, x, y, z, h, weights, weight_function, radius, ndim,
                                                        raster, z_slice)
    need = ctypes.c_size_t(0)
    dummy = (ctypes.c_void_p * len(weights))(*[8] * len(weights))
    with torch.cuda.device(device):
        rc = _lib.load().sphb_scatter_planned(ctypes.byref(part), ctypes.byref(kern), ctypes.byref(rast),
                                              1 if dim3 else 0, 1 if use_z else 0, float(z_slice or 0.0), 1, dummy,
                                              None, 0, ctypes.byref(need), ctypes.byref(plan.c), _stream(device))
    if rc not in (_lib.SPHB_ERR_WORKSPACE, _lib.SPHB_OK):
        _lib.check(rc)
    _workspace(device, need.value)
    return int(need.value)


def scatter_overflowed(device=None) -> bool:
    """True if the last planned call on `device` met more particles / pairs than its plan (synchronises)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ws = _workspaces[str(device)]
    flag = ctypes.c_int(0)
    with torch.cuda.device(device):
        _lib.check(_lib.load().sphb_scatter_overflowed(ctypes.c_void_p(ws.data_ptr()), ctypes.byref(flag),
                                                       _stream(device)))
    return bool(flag.value)


def scatter2d(x, y, h, weights, weight_function, radius, ndim, raster: Raster, z=None, z_slice=None,
              out=None, accumulate=False) -> List[torch.Tensor]:
    """B3: sampled scatter onto an image (cpu_backend.py:226-313).

    z/z_slice given -> cross-section (cull |z_slice - z| >= R h, distance
    includes dz); omitted -> 2-D render or column projection.  One pass
    accumulates every weight column.  Returns float64 (ny, nx) tensors.
    """
    use_z = z is not None and z_slice is not None
    return _scatter(False, x, y, z if use_z else None, h, weights, weight_function, radius, ndim, raster,
                    use_z, 0.0 if z_slice is None else z_slice, out, accumulate)


def scatter3d(x, y, z, h, weights, weight_function, radius, raster: Raster, out=None,
              accumulate=False, stage_planes=None, on_stage=None) -> List[torch.Tensor]:
    """B2/B3: sampled scatter onto a voxel grid in one pass (cpu_backend.py:193-220).

    `stage_planes` (ascending z-plane indices, the last one >= nz) renders the grid in z-slab
    stages: stage s finishes every plane below stage_planes[s], and `on_stage(s)` is called right
    after its kernels have been enqueued on the current stream (sarracen_b200.parallel records an
    event there and reduces that slab across GPUs while the later stages render)."""
    return _scatter(True, x, y, z, h, weights, weight_function, radius, 3, raster, False, 0.0, out, accumulate,
                    stage_planes, on_stage)


def line2d(x, y, h, weights, weight_function, radius, pixels, x1, x2, y1, y2) -> List[torch.Tensor]:
    """B6: line cut through 2-D data (cpu_backend.py:450-538)."""
    device = x.device
    out = [torch.empty(int(pixels), dtype=torch.float64, device=device) for _ in weights]
    kern, _keep = _kernel(weight_function, radius, 2, device)
    part = _particles(x, y, None, h, weights)
    with torch.cuda.device(device):
        _lib.check(_lib.load().sphb_line2d(ctypes.byref(part), ctypes.byref(kern), int(pixels), float(x1), float(x2),
                                           float(y1), float(y2), 0, _out_ptrs(out), _stream(device)))
    return out


def line3d(x, y, z, h, weights, weight_function, radius, pixels, x1, x2, y1, y2, z1, z2) -> List[torch.Tensor]:
    """B7: line cut through 3-D data (cpu_backend.py:540-609)."""
    device = x.device
    out = [torch.empty(int(pixels), dtype=torch.float64, device=device) for _ in weights]
    kern, _keep = _kernel(weight_function, radius, 3, device)
    part = _particles(x, y, z, h, weights)
    with torch.cuda.device(device):
        _lib.check(_lib.load().sphb_line3d(ctypes.byref(part), ctypes.byref(kern), int(pixels), float(x1), float(x2),
                                           float(y1), float(y2), float(z1), float(z2), 0, _out_ptrs(out),
                                           _stream(device)))
    return out


def _exact(fn_name, x, y, h, weights, raster: Raster, out=None, accumulate=False) -> List[torch.Tensor]:
    device = x.device
    if out is None:
        if accumulate:
            raise ValueError("accumulate=True needs the rasters to add to (out=)")
        out = [torch.empty((raster.ny, raster.nx), dtype=torch.float64, device=device) for _ in weights]
    else:
        out = list(out)
        for o in out:
            _require_cuda(o, "out")
            if o.dtype != torch.float64 or tuple(o.shape) != (raster.ny, raster.nx) or not o.is_contiguous():
                raise ValueError("out rasters must be contiguous float64 of shape (ny, nx)")
        if len(out) != len(weights):
            raise ValueError("one output raster per weight column")
    part = _particles(x, y, None, h, weights)
    rast = raster.c()
    with torch.cuda.device(device):
        fn = getattr(_lib.load(), fn_name)
        need = ctypes.c_size_t(0)
        ws = _workspaces.get(str(device))
        for _attempt in range(3):   # exact2d sizes its table of line terms after a first count
            rc = fn(ctypes.byref(part), ctypes.byref(rast), 1 if accumulate else 0, _out_ptrs(out),
                    ctypes.c_void_p(ws.data_ptr()) if ws is not None else None,
                    ws.numel() if ws is not None else 0, ctypes.byref(need), _stream(device))
            if rc != _lib.SPHB_ERR_WORKSPACE:
                break
            ws = None
            ws = _workspace(device, need.value)
        _lib.check(rc)
    return out


def exact2d(x, y, h, weights, raster: Raster, out=None, accumulate=False) -> List[torch.Tensor]:
    """B4: exact pixel integrals of the 2-D cubic spline (cpu_backend.py:317-447)."""
    return _exact("sphb_exact2d", x, y, h, weights, raster, out, accumulate)


def exact3d_proj(x, y, h, weights, raster: Raster, out=None, accumulate=False) -> List[torch.Tensor]:
    """B5: exact column integrals of the 3-D cubic spline (cpu_backend.py:611-720)."""
    return _exact("sphb_exact3d_proj", x, y, h, weights, raster, out, accumulate)


def count_contributions(x, y, h, radius, raster: Raster, z=None, z_slice=None) -> int:
    """Exact number of (particle, pixel) pairs with q <= R (SURVEY 8(d) 'pixel-contribution')."""
    device = x.device
    use_z = z is not None and z_slice is not None and raster.nz == 1
    part = _particles(x, y, z if (use_z or raster.nz > 1) else None, h, [h])
    out = torch.zeros(1, dtype=torch.int64, device=device)
    rast = raster.c()
    with torch.cuda.device(device):
        _lib.check(_lib.load().sphb_count_contributions(ctypes.byref(part), float(radius), ctypes.byref(rast),
                                                        1 if use_z else 0, float(z_slice or 0.0),
                                                        ctypes.c_void_p(out.data_ptr()), _stream(device)))
    return int(out.item())


def normalize_(num: torch.Tensor, den: torch.Tensor) -> torch.Tensor:
    """N5: num <- nan_to_num(num / den) in place on the device (interpolate.py:594)."""
    _require_cuda(num, "num")
    _require_cuda(den, "den")
    if num.dtype != torch.float64 or den.dtype != torch.float64 or num.shape != den.shape \
            or not num.is_contiguous() or not den.is_contiguous():
        raise ValueError("normalize_ needs contiguous float64 tensors of one shape")
    with torch.cuda.device(num.device):
        _lib.check(_lib.load().sphb_normalize(ctypes.c_void_p(num.data_ptr()), ctypes.c_void_p(den.data_ptr()),
                                              num.numel(), _stream(num.device)))
    return num


def sum_slots_(dst: torch.Tensor, own: torch.Tensor, slots: torch.Tensor, skip: int) -> torch.Tensor:
    """dst = own + sum over the rows s != skip of `slots` (n_slots, n): the receive side of the peer-memory
    exchange along z (sarracen_b200.parallel).  `slots` may be a view into peer-mapped (symmetric) memory."""
    for t, name in ((dst, "dst"), (own, "own"), (slots, "slots")):
        _require_cuda(t, name)
        if t.dtype != torch.float64 or not t.is_contiguous():
            raise ValueError("sum_slots_ needs contiguous float64 tensors")
    n = dst.numel()
    if own.numel() != n or slots.dim() != 2 or slots.shape[1] != n:
        raise ValueError("sum_slots_: dst and own of n elements, slots of shape (n_slots, n)")
    with torch.cuda.device(dst.device):
        _lib.check(_lib.load().sphb_sum_slots(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(own.data_ptr()),
                                              ctypes.c_void_p(slots.data_ptr()), int(slots.shape[0]), n, int(skip), n,
                                              _stream(dst.device)))
    return dst
