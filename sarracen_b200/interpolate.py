"""Device-side interpolation front-end.

Same functions, argument names, defaults, error behaviour and return types as
`sarracen.interpolate` (sarracen/interpolate/interpolate.py:493-1637), restated
for a GPU-resident pipeline:

* a dataframe column is uploaded once per call and everything after that --
  weights `A m / rho` (interpolate.py:448-472), the `hmin` clamp (:475-490),
  rotation about an origin (:258-384), default bounds (:81-122) -- is computed
  on the device;
* the target, the second vector component and the normalisation weight share
  ONE footprint walk (the reference makes one backend call per weight and a
  second round for `normalize=True`, e.g. :1068-1081, :1208-1224);
* `nan_to_num(grid / norm_grid)` (:594) runs on the device (`sphb_normalize`).

`data` is a `SarracenDataFrame` or a `sarracen_b200.ParticleFrame` (anything with
`xcol/ycol/zcol/hcol/mcol/rhocol`, `params`, `kernel`, `get_dim()`).  Only
`backend='gpu'` (or None) is served here; `'cpu'` raises -- this package has no
CPU path -- and any other code raises `ValueError("Invalid backend")` like
`get_backend` (:1640-1658).
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .kernels import BaseKernel

try:  # scipy is only needed to turn a Rotation / Euler angles into a 3x3 matrix
    from scipy.spatial.transform import Rotation
except ImportError:  # pragma: no cover
    Rotation = None


# ----------------------------------------------------------------------------- host-side checks
def _default_xy(data, x, y):
    """interpolate.py:15-40."""
    if x is None:
        x = data.xcol if not y == data.xcol else data.ycol
    if y is None:
        y = data.ycol if not x == data.ycol else data.xcol
    return x, y


def _default_xyz(data, x, y, z):
    """interpolate.py:43-78."""
    xcol, ycol, zcol = data.xcol, data.ycol, data.zcol
    if x is None:
        x = xcol if not y == xcol and not z == xcol else ycol if not y == ycol and not z == ycol else zcol
    if y is None:
        y = ycol if not x == ycol and not z == ycol else xcol if not x == xcol and not z == xcol else zcol
    if z is None:
        z = zcol if not x == zcol and not y == zcol else ycol if not x == ycol and not y == ycol else xcol
    return x, y, z


def _set_pixels(x_pixels, y_pixels, xlim, ylim):
    """interpolate.py:125-161: default 512, aspect-preserving rint."""
    dx, dy = xlim[1] - xlim[0], ylim[1] - ylim[0]
    if y_pixels is None:
        if x_pixels is None:
            x_pixels = 512
        y_pixels = int(np.rint(x_pixels * (dy / dx)))
    elif x_pixels is None:
        x_pixels = int(np.rint(y_pixels * (dx / dy)))
    return x_pixels, y_pixels


def _verify_columns(data, x, y):
    """interpolate.py:164-197."""
    if x not in data.columns:
        raise KeyError(f"x-directional column '{x}' does not exist in the provided dataset.")
    if y not in data.columns:
        raise KeyError(f"y-directional column '{y}' does not exist in the provided dataset.")
    if data.hcol is None:
        raise KeyError("Smoothing length column does not exist in the provided dataset.")
    if data.mcol is None and "mass" not in data.params:
        raise KeyError("Missing particle mass data in this SarracenDataFrame.")
    if data.rhocol is None and "hfact" not in data.params:
        raise KeyError("Density cannot be derived from the columns in this SarracenDataFrame.")


def _check_boundaries(x_pixels, y_pixels, xlim, ylim):
    """interpolate.py:200-231."""
    if xlim[1] - xlim[0] <= 0:
        raise ValueError("`xlim` max must be greater than min!")
    if ylim[1] - ylim[0] <= 0:
        raise ValueError("`ylim` max must be greater than min!")
    if x_pixels <= 0:
        raise ValueError("`x_pixels` must be greater than zero!")
    if y_pixels <= 0:
        raise ValueError("`y_pixels` must be greater than zero!")


def _check_dimension(data, dim):
    """interpolate.py:234-255."""
    if dim not in [2, 3]:
        raise ValueError("`dim` must be 2 or 3.")
    if data.get_dim() != dim:
        raise ValueError(f"Dataset is not {dim}-dimensional.")


def _check_backend(backend, data):
    backend = backend if backend is not None else getattr(data, "backend", "gpu")
    if backend == "cpu":
        raise ValueError("sarracen_b200 has no CPU path: use Sarracen's own backend='cpu'")
    if backend != "gpu":
        raise ValueError("Invalid backend")


def _corotate(corotation, rotation):
    """interpolate.py:387-428 (does not modify its argument)."""
    c0 = [float(v) for v in corotation[0]]
    c1 = [float(corotation[1][i]) - c0[i] for i in range(3)]
    angle = -np.arctan2(c1[1], c1[0])
    if rotation is None:
        rotation = np.array([angle * 180 / np.pi, 0, 0])
    else:
        if Rotation is not None and isinstance(rotation, Rotation):
            rotation = rotation.as_rotvec(degrees=True)
        rotation = np.array([angle * 180 / np.pi + rotation[0], rotation[1], rotation[2]])
        rotation = Rotation.from_euler("zyx", rotation, degrees=True)
    return rotation, c0


# ----------------------------------------------------------------------------- device-side prep
class _Device:
    """Columns of one call on the device, in the call's compute dtype."""

    def __init__(self, data, columns: Sequence[str]):
        if not torch.cuda.is_available():
            raise RuntimeError("sarracen_b200.interpolate needs a CUDA device; there is no CPU fallback")
        self.data = data
        self.dev = torch.device("cuda", torch.cuda.current_device())
        flat = []
        for c in columns:  # a target may be a list of labels
            flat.extend(c if isinstance(c, (list, tuple)) else [c])
        if hasattr(data, "column_dtype"):   # DeviceFrame: no host copies
            dtypes = [data.column_dtype(c) for c in flat if isinstance(c, str) and c in data.columns]
        else:
            dtypes = [np.asarray(data[c]).dtype for c in flat if isinstance(c, str) and c in data.columns]
        self.dtype = torch.float32 if dtypes and all(dt == np.float32 for dt in dtypes) else torch.float64
        self.cache = {}
        self.n = len(data)

    def col(self, name: str) -> torch.Tensor:
        if name not in self.cache:
            resident = getattr(self.data, "_device_cache", None)
            if resident and name in resident and resident[name].device == self.dev:
                self.cache[name] = resident[name].to(self.dtype).contiguous()  # no copy if dtype matches
                return self.cache[name]
            arr = np.ascontiguousarray(self.data[name].to_numpy())
            if not arr.flags.writeable:  # pandas copy-on-write views; torch wants a writable buffer
                arr = arr.copy()
            self.cache[name] = torch.from_numpy(arr).to(self.dev).to(self.dtype).contiguous()
        return self.cache[name]

    def raw(self, name: str) -> torch.Tensor:
        """A column on the device in ITS OWN dtype (analysis helpers that do their arithmetic in
        double whatever the column precision)."""
        key = ("raw", name)
        if key not in self.cache:
            resident = getattr(self.data, "_device_cache", None)
            if resident and name in resident and resident[name].device == self.dev:
                self.cache[key] = resident[name].contiguous()
            else:
                arr = np.ascontiguousarray(self.data[name].to_numpy())
                if not arr.flags.writeable:
                    arr = arr.copy()
                self.cache[key] = torch.from_numpy(arr).to(self.dev).contiguous()
        return self.cache[key]

    def mass(self):
        """interpolate.py:431-435."""
        if self.data.mcol is None:
            return float(self.data.params["mass"])
        return self.col(self.data.mcol)

    def density(self) -> torch.Tensor:
        """interpolate.py:438-445: rho column, else m (hfact / h)^dim."""
        if self.data.rhocol is None:
            hfact = float(self.data.params["hfact"])
            return (hfact / self.col(self.data.hcol)) ** self.data.get_dim() * self.mass()
        return self.col(self.data.rhocol)

    def weight(self, target, dens_weight: bool) -> torch.Tensor:
        """interpolate.py:448-472; `target` is a column label, a device tensor or None (unity)."""
        key = ("w", target if (target is None or isinstance(target, str)) else id(target), bool(dens_weight))
        if key not in self.cache:
            self.cache[key] = self._weight(target, dens_weight)
        return self.cache[key]

    def _weight(self, target, dens_weight: bool) -> torch.Tensor:
        if target is None:
            tdata = None
        elif isinstance(target, str):
            if target == "rho":
                tdata = self.density()
            else:
                if target not in self.data.columns:
                    raise KeyError(f"Target column '{target}' does not exist in provided dataset.")
                tdata = self.col(target)
        elif isinstance(target, torch.Tensor):
            tdata = target
        elif isinstance(target, np.ndarray):
            tdata = torch.from_numpy(np.ascontiguousarray(target)).to(self.dev).to(self.dtype)
        else:
            raise KeyError(f"Target must be of type str or ndarray. Found: '{type(target)}'")
        m = self.mass()
        w = m if tdata is None else tdata * m
        if not dens_weight:
            w = w / self.density()
        if not isinstance(w, torch.Tensor):
            w = torch.full((self.n,), float(w), dtype=self.dtype, device=self.dev)
        return w.to(self.dtype).contiguous()

    def smoothing(self, hmin: bool, pix_size: Optional[float]) -> torch.Tensor:
        """interpolate.py:475-490: h, clamped below by half a pixel when hmin."""
        h = self.col(self.data.hcol)
        if hmin:
            h = torch.clamp(h, min=0.5 * float(pix_size))
        return h.contiguous()

    def rotate(self, vx, vy, vz, rotation, rot_origin):
        """interpolate.py:258-329: (v - o) R^T + o with the reference's choice of o."""
        if rotation is None:
            return vx, vy, vz
        if Rotation is None:
            raise RuntimeError("scipy is required for rotated views")
        rot = rotation if isinstance(rotation, Rotation) else Rotation.from_euler("zyx", rotation, degrees=True)
        v = torch.stack([vx, vy, vz], dim=1).to(torch.float64)
        if rot_origin is None:
            origin = torch.zeros(3, dtype=torch.float64, device=self.dev)
        elif isinstance(rot_origin, str):
            if rot_origin == "com":
                origin = torch.tensor([float(c) for c in self.data.centre_of_mass()], dtype=torch.float64,
                                      device=self.dev)
            elif rot_origin == "midpoint":
                origin = (v.min(0).values + v.max(0).values) / 2
            else:
                raise ValueError("rot_origin should be an [x, y, z] point or 'com' or 'midpoint'")
        elif isinstance(rot_origin, (list, np.ndarray)) or hasattr(rot_origin, "to_numpy"):
            if len(rot_origin) != 3:
                raise ValueError("rot_origin should specify [x, y, z] point.")
            origin = torch.tensor(np.asarray(rot_origin, dtype=np.float64), device=self.dev)
        else:
            raise ValueError("rot_origin should be an [x, y, z] point or 'com' or 'midpoint'")
        mat = torch.tensor(rot.as_matrix(), dtype=torch.float64, device=self.dev)
        v = (v - origin) @ mat.T + origin
        v = v.to(self.dtype)
        return v[:, 0].contiguous(), v[:, 1].contiguous(), v[:, 2].contiguous()

    def rotate_xyz(self, x, y, z, rotation, rot_origin):
        """interpolate.py:332-384: rotate the global position columns, then hand back
        whichever of them (or any other column) the caller named."""
        d = self.data
        if rotation is None:
            return self.col(x), self.col(y), self.col(z)
        rx, ry, rz = self.rotate(self.col(d.xcol), self.col(d.ycol), self.col(d.zcol), rotation, rot_origin)

        def pick(c):
            return rx if c == d.xcol else ry if c == d.ycol else rz if c == d.zcol else self.col(c)
        return pick(x), pick(y), pick(z)


def _bounds(x_t: torch.Tensor, y_t: torch.Tensor, xlim, ylim):
    """interpolate.py:81-122: missing limits default to the data minimum / maximum."""
    def one(t, lim):
        lo = lim[0] if lim is not None and lim[0] is not None else None
        hi = lim[1] if lim is not None and lim[1] is not None else None
        lo = float(t.min().item()) if lo is None else lo
        hi = float(t.max().item()) if hi is None else hi
        return lo, hi
    return one(x_t, xlim), one(y_t, ylim)


def _finish(grids: List[torch.Tensor], norm: Optional[torch.Tensor]) -> List[np.ndarray]:
    """Optional on-device normalisation, then one D2H copy per grid."""
    if norm is not None:
        for g in grids:
            ops.normalize_(g, norm)
    from .backend import _host
    return [_host(g) for g in grids]


def _target_weights(dv, target, dens_weight, normalize):
    """Weight columns of one call.  `target` is a column label as in the reference, or -- an
    extension -- a list / tuple of up to three labels (four without normalisation) that share the
    bin, sort and footprint walk (several columns rendered on the same geometry)."""
    many = isinstance(target, (list, tuple))
    names = list(target) if many else [target]
    limit = ops._lib.SPHB_MAX_WEIGHTS - (1 if normalize else 0)
    if not 1 <= len(names) <= limit:
        raise ValueError(f"between 1 and {limit} targets can share one call")
    weights = [dv.weight(t, dens_weight) for t in names]
    if normalize:
        weights.append(dv.weight(None, dens_weight))
    return weights, many


def _finish_targets(out, normalize, many):
    grids = _finish(out[:-1] if normalize else out, out[-1] if normalize else None)
    return grids if many else grids[0]


def _image(dv, x_t, y_t, h_t, weights, wf, radius, ndim, raster, exact, exact_fn, z_t=None, z_slice=None):
    if exact:
        return exact_fn(x_t, y_t, h_t, weights, raster)
    return ops.scatter2d(x_t, y_t, h_t, weights, wf, radius, ndim, raster, z=z_t, z_slice=z_slice)


# ----------------------------------------------------------------------------- public functions
def interpolate_2d(data, target, x=None, y=None, kernel: Optional[BaseKernel] = None, x_pixels=None,
                   y_pixels=None, xlim=None, ylim=None, exact=False, backend=None, dens_weight=False,
                   normalize=True, hmin=False) -> np.ndarray:
    """2-D data onto a 2-D grid (interpolate.py:493-596)."""
    _check_dimension(data, 2)
    x, y = _default_xy(data, x, y)
    _verify_columns(data, x, y)
    _check_backend(backend, data)
    dv = _Device(data, [x, y, data.hcol, data.mcol, data.rhocol, target])
    xlim, ylim = _bounds(dv.col(x), dv.col(y), xlim, ylim)
    x_pixels, y_pixels = _set_pixels(x_pixels, y_pixels, xlim, ylim)
    _check_boundaries(x_pixels, y_pixels, xlim, ylim)
    weights, many = _target_weights(dv, target, dens_weight, normalize)
    kernel = kernel if kernel is not None else data.kernel
    pix = max((xlim[1] - xlim[0]) / x_pixels, (ylim[1] - ylim[0]) / y_pixels)
    r = ops.Raster(x_pixels, y_pixels, xlim[0], xlim[1], ylim[0], ylim[1])
    out = _image(dv, dv.col(x), dv.col(y), dv.smoothing(hmin, pix), weights, kernel.w, kernel.get_radius(), 2, r,
                 exact, ops.exact2d)
    return _finish_targets(out, normalize, many)


def interpolate_2d_vec(data, target_x, target_y, x=None, y=None, kernel: Optional[BaseKernel] = None,
                       x_pixels=None, y_pixels=None, xlim=None, ylim=None, exact=False, backend=None,
                       dens_weight=False, normalize=True, hmin=False) -> Tuple[np.ndarray, np.ndarray]:
    """2-D vector data onto two 2-D grids (interpolate.py:599-708)."""
    _check_dimension(data, 2)
    x, y = _default_xy(data, x, y)
    _verify_columns(data, x, y)
    _check_backend(backend, data)
    dv = _Device(data, [x, y, data.hcol, data.mcol, data.rhocol, target_x, target_y])
    xlim, ylim = _bounds(dv.col(x), dv.col(y), xlim, ylim)
    x_pixels, y_pixels = _set_pixels(x_pixels, y_pixels, xlim, ylim)
    _check_boundaries(x_pixels, y_pixels, xlim, ylim)
    weights = [dv.weight(target_x, dens_weight), dv.weight(target_y, dens_weight)]
    if normalize:
        weights.append(dv.weight(None, dens_weight))
    kernel = kernel if kernel is not None else data.kernel
    pix = max((xlim[1] - xlim[0]) / x_pixels, (ylim[1] - ylim[0]) / y_pixels)
    r = ops.Raster(x_pixels, y_pixels, xlim[0], xlim[1], ylim[0], ylim[1])
    out = _image(dv, dv.col(x), dv.col(y), dv.smoothing(hmin, pix), weights, kernel.w, kernel.get_radius(), 2, r,
                 exact, ops.exact2d)
    gx, gy = _finish(out[:2], out[2] if normalize else None)
    return gx, gy


def _line_limits(lim):
    return (lim, lim) if isinstance(lim, (float, int)) else lim


def interpolate_2d_line(data, target, x=None, y=None, kernel: Optional[BaseKernel] = None, pixels=None,
                        xlim=None, ylim=None, backend=None, dens_weight=False, normalize=True,
                        hmin=False) -> np.ndarray:
    """Line cut through 2-D data (interpolate.py:711-820)."""
    _check_dimension(data, 2)
    x, y = _default_xy(data, x, y)
    _verify_columns(data, x, y)
    _check_backend(backend, data)
    dv = _Device(data, [x, y, data.hcol, data.mcol, data.rhocol, target])
    weights, many = _target_weights(dv, target, dens_weight, normalize)
    xlim, ylim = _bounds(dv.col(x), dv.col(y), _line_limits(xlim), _line_limits(ylim))
    if xlim[0] == xlim[1] and ylim[0] == ylim[1]:
        raise ValueError("Zero length cross section!")
    kernel = kernel if kernel is not None else data.kernel
    pixels = pixels if pixels is not None else 512
    if pixels <= 0:
        raise ValueError("pixcount must be greater than zero!")
    pix = np.sqrt((xlim[1] - xlim[0]) ** 2 + (ylim[1] - ylim[0]) ** 2) / pixels
    out = ops.line2d(dv.col(x), dv.col(y), dv.smoothing(hmin, pix), weights, kernel.w, kernel.get_radius(), pixels,
                     xlim[0], xlim[1], ylim[0], ylim[1])
    return _finish_targets(out, normalize, many)


def interpolate_3d_line(data, target, x=None, y=None, z=None, kernel: Optional[BaseKernel] = None, pixels=None,
                        xlim=None, ylim=None, zlim=None, backend=None, dens_weight=False, normalize=True,
                        hmin=False) -> np.ndarray:
    """Line cut through 3-D data (interpolate.py:823-943)."""
    _check_dimension(data, 3)
    x, y, z = _default_xyz(data, x, y, z)
    _verify_columns(data, x, y)
    _check_backend(backend, data)
    dv = _Device(data, [x, y, z, data.hcol, data.mcol, data.rhocol, target])
    weights, many = _target_weights(dv, target, dens_weight, normalize)
    xlim, ylim, zlim = _line_limits(xlim), _line_limits(ylim), _line_limits(zlim)
    zmin = float(dv.col(z).min().item())  # both defaults are the minimum (interpolate.py:905-906)
    zlim = (zmin if zlim is None or zlim[0] is None else zlim[0], zmin if zlim is None or zlim[1] is None else zlim[1])
    xlim, ylim = _bounds(dv.col(x), dv.col(y), xlim, ylim)
    if ylim[1] == ylim[0] and xlim[1] == xlim[0] and zlim[1] == zlim[0]:
        raise ValueError("Zero length cross section!")
    kernel = kernel if kernel is not None else data.kernel
    pixels = pixels if pixels is not None else 512
    if pixels <= 0:
        raise ValueError("pixcount must be greater than zero!")
    pix = np.sqrt((xlim[1] - xlim[0]) ** 2 + (ylim[1] - ylim[0]) ** 2 + (zlim[1] - zlim[0]) ** 2) / pixels
    out = ops.line3d(dv.col(x), dv.col(y), dv.col(z), dv.smoothing(hmin, pix), weights, kernel.w,
                     kernel.get_radius(), pixels, xlim[0], xlim[1], ylim[0], ylim[1], zlim[0], zlim[1])
    return _finish_targets(out, normalize, many)


def interpolate_3d_proj(data, target, x=None, y=None, kernel: Optional[BaseKernel] = None, integral_samples=1000,
                        corotation=None, rotation=None, rot_origin=None, x_pixels=None, y_pixels=None, xlim=None,
                        ylim=None, exact=False, backend=None, dens_weight=None, normalize=True,
                        hmin=False, _dv=None) -> np.ndarray:
    """Column-integrated view of 3-D data (interpolate.py:946-1083)."""
    _check_dimension(data, 3)
    x, y, z = _default_xyz(data, x, y, None)
    _verify_columns(data, x, y)
    _check_backend(backend, data)
    if dens_weight is None:
        flags = {t != "rho" for t in (target if isinstance(target, (list, tuple)) else [target])}
        if len(flags) != 1:
            raise ValueError("pass dens_weight explicitly when 'rho' is rendered together with other targets")
        dens_weight = flags.pop()
    dv = _dv if _dv is not None else _Device(data, [data.xcol, data.ycol, data.zcol, x, y, data.hcol, data.mcol,
                                                    data.rhocol, target])
    weights, many = _target_weights(dv, target, dens_weight, normalize)
    if corotation is not None:
        rotation, rot_origin = _corotate(corotation, rotation)
    x_t, y_t, _ = dv.rotate_xyz(x, y, z, rotation, rot_origin)
    xlim, ylim = _bounds(x_t, y_t, xlim, ylim)
    x_pixels, y_pixels = _set_pixels(x_pixels, y_pixels, xlim, ylim)
    _check_boundaries(x_pixels, y_pixels, xlim, ylim)
    kernel = kernel if kernel is not None else data.kernel
    wf = kernel.get_column_kernel_func(integral_samples)
    pix = max((xlim[1] - xlim[0]) / x_pixels, (ylim[1] - ylim[0]) / y_pixels)
    r = ops.Raster(x_pixels, y_pixels, xlim[0], xlim[1], ylim[0], ylim[1])
    out = _image(dv, x_t, y_t, dv.smoothing(hmin, pix), weights, wf, kernel.get_radius(), 2, r, exact,
                 ops.exact3d_proj)
    return _finish_targets(out, normalize, many)


def interpolate_3d_vec(data, target_x, target_y, target_z, x=None, y=None, kernel: Optional[BaseKernel] = None,
                       integral_samples=1000, rotation=None, rot_origin=None, x_pixels=None, y_pixels=None,
                       xlim=None, ylim=None, exact=False, backend=None, dens_weight=False, normalize=True,
                       hmin=False) -> Tuple[np.ndarray, np.ndarray]:
    """Column-integrated view of 3-D vector data (interpolate.py:1086-1226)."""
    _check_dimension(data, 3)
    x, y, z = _default_xyz(data, x, y, None)
    _verify_columns(data, x, y)
    _check_backend(backend, data)
    dv = _Device(data, [data.xcol, data.ycol, data.zcol, x, y, data.hcol, data.mcol, data.rhocol, target_x,
                        target_y, target_z])
    x_t, y_t, _ = dv.rotate_xyz(x, y, z, rotation, rot_origin)
    xlim, ylim = _bounds(x_t, y_t, xlim, ylim)
    x_pixels, y_pixels = _set_pixels(x_pixels, y_pixels, xlim, ylim)
    _check_boundaries(x_pixels, y_pixels, xlim, ylim)
    for lab, t in (("x", target_x), ("y", target_y), ("z", target_z)):
        if t not in data.columns:
            raise KeyError(f"{lab}-directional target column '{t}' does not exist in the provided dataset.")
    tx, ty, _ = dv.rotate(dv.col(target_x), dv.col(target_y), dv.col(target_z), rotation, rot_origin)
    weights = [dv.weight(tx, dens_weight), dv.weight(ty, dens_weight)]
    if normalize:
        weights.append(dv.weight(None, dens_weight))
    kernel = kernel if kernel is not None else data.kernel
    wf = kernel.get_column_kernel_func(integral_samples)
    pix = max((xlim[1] - xlim[0]) / x_pixels, (ylim[1] - ylim[0]) / y_pixels)
    r = ops.Raster(x_pixels, y_pixels, xlim[0], xlim[1], ylim[0], ylim[1])
    out = _image(dv, x_t, y_t, dv.smoothing(hmin, pix), weights, wf, kernel.get_radius(), 2, r, exact,
                 ops.exact3d_proj)
    gx, gy = _finish(out[:2], out[2] if normalize else None)
    return gx, gy


def interpolate_3d_cross(data, target, x=None, y=None, z=None, z_slice=None, kernel: Optional[BaseKernel] = None,
                         corotation=None, rotation=None, rot_origin=None, x_pixels=None, y_pixels=None,
                         xlim=None, ylim=None, backend=None, dens_weight=False, normalize=True,
                         hmin=False) -> np.ndarray:
    """Cross-section of 3-D data at `z_slice` (interpolate.py:1229-1361).

    Extension: `z_slice` may be a sequence of K depths; the result is then a
    (K, ny, nx) stack.  Evenly spaced depths are rendered in ONE pass -- one upload, one bin, one
    walk over the particles -- as the K planes of a voxel grid (the reference's grid *is* the loop of
    cross-sections, cpu_backend.py:213-218); other sequences share the upload, the rotation and the
    weights and make one pass per depth."""
    _check_dimension(data, 3)
    x, y, z = _default_xyz(data, x, y, z)
    _verify_columns(data, x, y)
    _check_backend(backend, data)
    dv = _Device(data, [data.xcol, data.ycol, data.zcol, x, y, z, data.hcol, data.mcol, data.rhocol, target])
    if z_slice is None:
        z_slice = float(dv.col(z).to(torch.float64).mean().item())  # of the unrotated column (:1327-1328)
    depths = None
    if isinstance(z_slice, (list, tuple, np.ndarray)):
        depths = [float(v) for v in np.asarray(z_slice, dtype=np.float64).ravel()]
        if not depths:
            raise ValueError("`z_slice` is empty")
    weights, many = _target_weights(dv, target, dens_weight, normalize)
    kernel = kernel if kernel is not None else data.kernel
    if corotation is not None:
        rotation, rot_origin = _corotate(corotation, rotation)
    x_t, y_t, z_t = dv.rotate_xyz(x, y, z, rotation, rot_origin)
    xlim, ylim = _bounds(x_t, y_t, xlim, ylim)
    x_pixels, y_pixels = _set_pixels(x_pixels, y_pixels, xlim, ylim)
    _check_boundaries(x_pixels, y_pixels, xlim, ylim)
    pix = max((xlim[1] - xlim[0]) / x_pixels, (ylim[1] - ylim[0]) / y_pixels)
    r = ops.Raster(x_pixels, y_pixels, xlim[0], xlim[1], ylim[0], ylim[1])
    h_t = dv.smoothing(hmin, pix)
    if depths is None:
        out = ops.scatter2d(x_t, y_t, h_t, weights, kernel.w, kernel.get_radius(), 3, r, z=z_t, z_slice=float(z_slice))
        return _finish_targets(out, normalize, many)
    k = len(depths)
    step = (depths[-1] - depths[0]) / (k - 1) if k > 1 else 0.0
    even = k > 1 and step > 0 and all(abs(depths[i] - (depths[0] + i * step)) <= 1e-12 * max(1.0, abs(depths[i]))
                                      for i in range(k))
    if even:
        # plane i of the grid has its centre at z_min + (i + 1/2) step = depths[i]
        r3 = ops.Raster(x_pixels, y_pixels, xlim[0], xlim[1], ylim[0], ylim[1], k, depths[0] - 0.5 * step,
                        depths[-1] + 0.5 * step)
        out = ops.scatter3d(x_t, y_t, z_t, h_t, weights, kernel.w, kernel.get_radius(), r3)
        return _finish_targets(out, normalize, many)
    stacks = None
    for i, d in enumerate(depths):
        out = ops.scatter2d(x_t, y_t, h_t, weights, kernel.w, kernel.get_radius(), 3, r, z=z_t, z_slice=d)
        res = _finish_targets(out, normalize, many)
        res = res if many else [res]
        if stacks is None:
            stacks = [np.empty((k,) + g.shape) for g in res]
        for s_, g in zip(stacks, res):
            s_[i] = g
    return stacks if many else stacks[0]


def interpolate_3d_proj_views(data, target, rotations, **kwargs) -> np.ndarray:
    """Extension: the column-integrated view of `target` for each rotation of
    `rotations` (the `render()` call pattern of a turntable, render.py:401-422), as a (K, ny, x) stack.
    The columns are uploaded, and the weights m / rho A built, ONCE; each view rotates, bins and renders on
    the device.  Give `xlim` / `ylim` (or the pixel counts) if all views are to share one raster; every
    other keyword is that of `interpolate_3d_proj`."""
    if "rotation" in kwargs:
        raise TypeError("pass `rotations`, not `rotation`")
    x, y, _ = _default_xyz(data, kwargs.get("x"), kwargs.get("y"), None)
    dv = _Device(data, [data.xcol, data.ycol, data.zcol, x, y, data.hcol, data.mcol, data.rhocol, target])
    views = [interpolate_3d_proj(data, target, rotation=rot, _dv=dv, **kwargs) for rot in rotations]
    if isinstance(views[0], list):
        return [np.stack([v[i] for v in views]) for i in range(len(views[0]))]
    if len({v.shape for v in views}) != 1:
        raise ValueError("the views have different rasters: give xlim / ylim and the pixel counts")
    return np.stack(views)


def interpolate_3d_cross_vec(data, target_x, target_y, target_z, z_slice=None, x=None, y=None, z=None,
                             kernel: Optional[BaseKernel] = None, rotation=None, rot_origin=None, x_pixels=None,
                             y_pixels=None, xlim=None, ylim=None, backend=None, dens_weight=False, normalize=True,
                             hmin=False) -> Tuple[np.ndarray, np.ndarray]:
    """Cross-section of 3-D vector data (interpolate.py:1364-1501)."""
    _check_dimension(data, 3)
    x, y, z = _default_xyz(data, x, y, z)
    _verify_columns(data, x, y)
    _check_backend(backend, data)
    dv = _Device(data, [data.xcol, data.ycol, data.zcol, x, y, z, data.hcol, data.mcol, data.rhocol, target_x,
                        target_y, target_z])
    if z_slice is None:
        z_slice = float(dv.col(z).to(torch.float64).mean().item())
    x_t, y_t, z_t = dv.rotate_xyz(x, y, z, rotation, rot_origin)
    xlim, ylim = _bounds(x_t, y_t, xlim, ylim)
    x_pixels, y_pixels = _set_pixels(x_pixels, y_pixels, xlim, ylim)
    _check_boundaries(x_pixels, y_pixels, xlim, ylim)
    for lab, t in (("x", target_x), ("y", target_y), ("z", target_z)):
        if t not in data.columns:
            raise KeyError(f"{lab}-directional target column '{t}' does not exist in the provided dataset.")
    tx, ty, _ = dv.rotate(dv.col(target_x), dv.col(target_y), dv.col(target_z), rotation, rot_origin)
    weights = [dv.weight(tx, dens_weight), dv.weight(ty, dens_weight)]
    if normalize:
        weights.append(dv.weight(None, dens_weight))
    kernel = kernel if kernel is not None else data.kernel
    pix = max((xlim[1] - xlim[0]) / x_pixels, (ylim[1] - ylim[0]) / y_pixels)
    r = ops.Raster(x_pixels, y_pixels, xlim[0], xlim[1], ylim[0], ylim[1])
    out = ops.scatter2d(x_t, y_t, dv.smoothing(hmin, pix), weights, kernel.w, kernel.get_radius(), 3, r, z=z_t,
                        z_slice=float(z_slice))
    gx, gy = _finish(out[:2], out[2] if normalize else None)
    return gx, gy


def interpolate_3d_grid(data, target, x=None, y=None, z=None, kernel: Optional[BaseKernel] = None, rotation=None,
                        rot_origin=None, x_pixels=None, y_pixels=None, z_pixels=None, xlim=None, ylim=None,
                        zlim=None, backend=None, dens_weight=False, normalize=True, hmin=False) -> np.ndarray:
    """3-D data onto a voxel grid (interpolate.py:1504-1637)."""
    _check_dimension(data, 3)
    x, y, z = _default_xyz(data, x, y, z)
    _verify_columns(data, x, y)
    _check_backend(backend, data)
    dv = _Device(data, [data.xcol, data.ycol, data.zcol, x, y, z, data.hcol, data.mcol, data.rhocol, target])
    weights, many = _target_weights(dv, target, dens_weight, normalize)
    x_t, y_t, z_t = dv.rotate_xyz(x, y, z, rotation, rot_origin)
    xlim, ylim = _bounds(x_t, y_t, xlim if xlim else (None, None), ylim if ylim else (None, None))
    zlim = zlim if zlim else (float(z_t.min().item()), float(z_t.max().item()))
    x_pixels, y_pixels = _set_pixels(x_pixels, y_pixels, xlim, ylim)
    if z_pixels is None:
        z_pixels = int(np.rint(x_pixels * ((zlim[1] - zlim[0]) / (xlim[1] - xlim[0]))))
    _check_boundaries(x_pixels, y_pixels, xlim, ylim)
    if zlim[1] - zlim[0] <= 0:
        raise ValueError("`z_max` must be greater than `z_min`!")
    if z_pixels <= 0:
        raise ValueError("`z_pixels` must be greater than zero!")
    kernel = kernel if kernel is not None else data.kernel
    pix = max((xlim[1] - xlim[0]) / x_pixels, (ylim[1] - ylim[0]) / y_pixels)  # x / y only (:1619-1620)
    r = ops.Raster(x_pixels, y_pixels, xlim[0], xlim[1], ylim[0], ylim[1], z_pixels, zlim[0], zlim[1])
    out = ops.scatter3d(x_t, y_t, z_t, dv.smoothing(hmin, pix), weights, kernel.w, kernel.get_radius(), r)
    return _finish_targets(out, normalize, many)
