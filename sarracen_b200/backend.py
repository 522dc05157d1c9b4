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
from typing import Tuple

import numpy as np
import torch

from . import ops
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


def _to_device(arrays, device):
    dtype = _common_dtype(arrays)
    out = []
    for a in arrays:
        if isinstance(a, torch.Tensor):
            t = a.to(device=device, dtype=dtype).contiguous().reshape(-1)
        else:
            arr = np.ascontiguousarray(np.asarray(a))
            if arr.dtype.byteorder not in ("=", "|") or not arr.flags.writeable:
                arr = arr.astype(arr.dtype.newbyteorder("="), copy=True)
            t = torch.from_numpy(arr).to(device=device, dtype=dtype).reshape(-1)
        out.append(t)
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


class B200Backend:
    """SPH interpolation on a B200 behind Sarracen's backend interface."""

    @staticmethod
    def interpolate_2d_render(x, y, weight, h, weight_function, kernel_radius, x_pixels, y_pixels,
                              x_min, x_max, y_min, y_max, exact) -> np.ndarray:
        dev = _device()
        xd, yd, wd, hd = _to_device([x, y, weight, h], dev)
        r = ops.Raster(x_pixels, y_pixels, x_min, x_max, y_min, y_max)
        if exact:
            return _host(ops.exact2d(xd, yd, hd, [wd], r)[0])
        wf = resolve_weight_function(weight_function, kernel_radius)
        return _host(ops.scatter2d(xd, yd, hd, [wd], wf, kernel_radius, 2, r)[0])

    @staticmethod
    def interpolate_2d_render_vec(x, y, weight_x, weight_y, h, weight_function, kernel_radius, x_pixels,
                                  y_pixels, x_min, x_max, y_min, y_max,
                                  exact) -> Tuple[np.ndarray, np.ndarray]:
        dev = _device()
        xd, yd, wx, wy, hd = _to_device([x, y, weight_x, weight_y, h], dev)
        r = ops.Raster(x_pixels, y_pixels, x_min, x_max, y_min, y_max)
        if exact:
            a, b = ops.exact2d(xd, yd, hd, [wx, wy], r)
        else:
            wf = resolve_weight_function(weight_function, kernel_radius)
            a, b = ops.scatter2d(xd, yd, hd, [wx, wy], wf, kernel_radius, 2, r)
        return _host(a), _host(b)

    @staticmethod
    def interpolate_2d_line(x, y, weight, h, weight_function, kernel_radius, pixels, x1, x2, y1,
                            y2) -> np.ndarray:
        dev = _device()
        xd, yd, wd, hd = _to_device([x, y, weight, h], dev)
        wf = resolve_weight_function(weight_function, kernel_radius)
        return _host(ops.line2d(xd, yd, hd, [wd], wf, kernel_radius, pixels, x1, x2, y1, y2)[0])

    @staticmethod
    def interpolate_3d_line(x, y, z, weight, h, weight_function, kernel_radius, pixels, x1, x2, y1, y2,
                            z1, z2) -> np.ndarray:
        dev = _device()
        xd, yd, zd, wd, hd = _to_device([x, y, z, weight, h], dev)
        wf = resolve_weight_function(weight_function, kernel_radius)
        return _host(ops.line3d(xd, yd, zd, hd, [wd], wf, kernel_radius, pixels, x1, x2, y1, y2, z1, z2)[0])

    @staticmethod
    def interpolate_3d_projection(x, y, weight, h, weight_function, kernel_radius, x_pixels, y_pixels,
                                  x_min, x_max, y_min, y_max, exact) -> np.ndarray:
        dev = _device()
        xd, yd, wd, hd = _to_device([x, y, weight, h], dev)
        r = ops.Raster(x_pixels, y_pixels, x_min, x_max, y_min, y_max)
        if exact:
            return _host(ops.exact3d_proj(xd, yd, hd, [wd], r)[0])
        wf = resolve_weight_function(weight_function, kernel_radius)
        return _host(ops.scatter2d(xd, yd, hd, [wd], wf, kernel_radius, 2, r)[0])

    @staticmethod
    def interpolate_3d_projection_vec(x, y, weight_x, weight_y, h, weight_function, kernel_radius,
                                      x_pixels, y_pixels, x_min, x_max, y_min, y_max,
                                      exact) -> Tuple[np.ndarray, np.ndarray]:
        dev = _device()
        xd, yd, wx, wy, hd = _to_device([x, y, weight_x, weight_y, h], dev)
        r = ops.Raster(x_pixels, y_pixels, x_min, x_max, y_min, y_max)
        if exact:
            a, b = ops.exact3d_proj(xd, yd, hd, [wx, wy], r)
        else:
            wf = resolve_weight_function(weight_function, kernel_radius)
            a, b = ops.scatter2d(xd, yd, hd, [wx, wy], wf, kernel_radius, 2, r)
        return _host(a), _host(b)

    @staticmethod
    def interpolate_3d_cross(x, y, z, z_slice, weight, h, weight_function, kernel_radius, x_pixels,
                             y_pixels, x_min, x_max, y_min, y_max) -> np.ndarray:
        dev = _device()
        xd, yd, zd, wd, hd = _to_device([x, y, z, weight, h], dev)
        r = ops.Raster(x_pixels, y_pixels, x_min, x_max, y_min, y_max)
        wf = resolve_weight_function(weight_function, kernel_radius)
        return _host(ops.scatter2d(xd, yd, hd, [wd], wf, kernel_radius, 3, r, z=zd, z_slice=float(z_slice))[0])

    @staticmethod
    def interpolate_3d_cross_vec(x, y, z, z_slice, weight_x, weight_y, h, weight_function, kernel_radius,
                                 x_pixels, y_pixels, x_min, x_max, y_min,
                                 y_max) -> Tuple[np.ndarray, np.ndarray]:
        dev = _device()
        xd, yd, zd, wx, wy, hd = _to_device([x, y, z, weight_x, weight_y, h], dev)
        r = ops.Raster(x_pixels, y_pixels, x_min, x_max, y_min, y_max)
        wf = resolve_weight_function(weight_function, kernel_radius)
        a, b = ops.scatter2d(xd, yd, hd, [wx, wy], wf, kernel_radius, 3, r, z=zd, z_slice=float(z_slice))
        return _host(a), _host(b)

    @staticmethod
    def interpolate_3d_grid(x, y, z, weight, h, weight_function, kernel_radius, x_pixels, y_pixels,
                            z_pixels, x_min, x_max, y_min, y_max, z_min, z_max) -> np.ndarray:
        dev = _device()
        xd, yd, zd, wd, hd = _to_device([x, y, z, weight, h], dev)
        r = ops.Raster(x_pixels, y_pixels, x_min, x_max, y_min, y_max, z_pixels, z_min, z_max)
        wf = resolve_weight_function(weight_function, kernel_radius)
        return _host(ops.scatter3d(xd, yd, zd, hd, [wd], wf, kernel_radius, r)[0])


def install() -> None:
    """Route Sarracen's backend='gpu' to `B200Backend` (see module docstring)."""
    import sarracen.interpolate
    import sarracen.interpolate.interpolate as front
    front.GPUBackend = B200Backend
    sarracen.interpolate.GPUBackend = B200Backend
