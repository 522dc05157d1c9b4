"""ctypes front for oracle/sph_oracle.c.

TEST INFRASTRUCTURE ONLY.  ``OracleBackend`` mirrors the static-method
interface of the reference's ``BaseBackend``
(sarracen/interpolate/base_backend.py:7-177) so that parity tests can call the
oracle and the B200 backend with the same arguments.  ``weight_function`` is
any object with ``kind`` ('cubic' | 'quartic' | 'quintic' | 'lut'), ``radius``
and, for 'lut', ``lut`` (the column-integrated table) -- ``OracleKernel`` here,
or the product's ``sarracen_b200.kernels.WeightFunction`` (duck-typed; the
oracle does not import the product).
"""
import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsph_oracle.so")
_KINDS = {"cubic": 0, "quartic": 1, "quintic": 2, "lut": 3}
_RADIUS = {"cubic": 2.0, "quartic": 2.5, "quintic": 3.0}
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)


class _CKernel(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("ndim", ctypes.c_int), ("radius", ctypes.c_double),
                ("lut", _dp), ("samples", ctypes.c_int)]


def build(force: bool = False) -> str:
    """Compile sph_oracle.c with gcc (-O2 -fopenmp) into oracle/_build/."""
    src = os.path.join(_HERE, "sph_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-std=c11", "-shared",
                               "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.orc_w.restype = ctypes.c_double
        lib.orc_w.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_int]
        lib.orc_kernel_eval.restype = ctypes.c_double
        lib.orc_kernel_eval.argtypes = [ctypes.POINTER(_CKernel), ctypes.c_double]
        lib.orc_line_int.restype = ctypes.c_double
        lib.orc_line_int.argtypes = [ctypes.c_double] * 4
        lib.orc_surface_int.restype = ctypes.c_double
        lib.orc_surface_int.argtypes = [ctypes.c_double] * 8
        lib.orc_column_kernel.argtypes = [ctypes.c_int, ctypes.c_int, _dp]
        lib.orc_set_num_threads.argtypes = [ctypes.c_int]
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(_load().orc_num_threads())


def set_num_threads(n: int) -> None:
    _load().orc_set_num_threads(int(n))


def kernel_w(kind: str, q: float, ndim: int) -> float:
    """w(q, ndim) of the analytic kernels (sarracen/kernels/*_spline.py)."""
    return float(_load().orc_w(_KINDS[kind], float(q), int(ndim)))


def column_kernel(kind: str, samples: int = 1000) -> np.ndarray:
    """Column-integrated table (sarracen/kernels/base_kernel.py:39-66)."""
    out = np.zeros(int(samples))
    _load().orc_column_kernel(_KINDS[kind], int(samples), out.ctypes.data_as(_dp))
    return out


def line_int(r0, d1, d2, h) -> float:
    return float(_load().orc_line_int(r0, d1, d2, h))


def surface_int(r0, x1, y1, x2, y2, wx, wy, h) -> float:
    return float(_load().orc_surface_int(r0, x1, y1, x2, y2, wx, wy, h))


class OracleKernel:
    """A weight function handle: an analytic kernel or a column table."""

    def __init__(self, kind: str, lut: Optional[np.ndarray] = None, radius: Optional[float] = None):
        self.kind = kind
        self.lut = None if lut is None else np.ascontiguousarray(lut, dtype=np.float64)
        self.radius = float(radius if radius is not None else _RADIUS[kind])

    @staticmethod
    def column(kind: str, samples: int = 1000) -> "OracleKernel":
        return OracleKernel("lut", column_kernel(kind, samples), _RADIUS[kind])

    def __call__(self, q: float, ndim: int) -> float:
        ck, _keep = _ckernel(self, self.radius, ndim)
        return float(_load().orc_kernel_eval(ctypes.byref(ck), float(q)))


def _ckernel(wf, radius, ndim):
    kind = wf.kind
    lut = getattr(wf, "lut", None)
    if kind == "lut":
        lut = np.ascontiguousarray(np.asarray(lut, dtype=np.float64))
        ck = _CKernel(_KINDS[kind], int(ndim), float(radius), lut.ctypes.data_as(_dp), lut.size)
    else:
        ck = _CKernel(_KINDS[kind], int(ndim), float(radius), None, 0)
    return ck, lut


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a), dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


class OracleBackend:
    """CPU oracle behind the reference's backend interface
    (sarracen/interpolate/base_backend.py:7-177; dispatch as in
    sarracen/interpolate/cpu_backend.py:14-220)."""

    last_contributions = 0

    @staticmethod
    def _fast_2d(x, y, z, z_slice, w, h, wf, radius, nx, ny, x_min, x_max, y_min, y_max, n_dims):
        x, y, w, h = _f64(x), _f64(y), _f64(w), _f64(h)
        z = None if z is None else _f64(z)
        ck, _keep = _ckernel(wf, radius, n_dims)
        out = np.zeros((int(ny), int(nx)))
        cnt = ctypes.c_int64(0)
        rc = _load().orc_fast_2d(_p(x), _p(y), None if z is None else _p(z), ctypes.c_double(z_slice),
                                 _p(w), _p(h), ctypes.c_int64(x.size), ctypes.byref(ck), int(nx), int(ny),
                                 ctypes.c_double(x_min), ctypes.c_double(x_max), ctypes.c_double(y_min),
                                 ctypes.c_double(y_max), int(n_dims), _p(out), ctypes.byref(cnt))
        if rc:
            raise MemoryError("oracle allocation failed")
        OracleBackend.last_contributions = cnt.value
        return out

    @staticmethod
    def _exact(fn, x, y, w, h, nx, ny, x_min, x_max, y_min, y_max):
        x, y, w, h = _f64(x), _f64(y), _f64(w), _f64(h)
        out = np.zeros((int(ny), int(nx)))
        cnt = ctypes.c_int64(0)
        rc = fn(_p(x), _p(y), _p(w), _p(h), ctypes.c_int64(x.size), int(nx), int(ny),
                ctypes.c_double(x_min), ctypes.c_double(x_max), ctypes.c_double(y_min),
                ctypes.c_double(y_max), _p(out), ctypes.byref(cnt))
        if rc:
            raise MemoryError("oracle allocation failed")
        OracleBackend.last_contributions = cnt.value
        return out

    @staticmethod
    def interpolate_2d_render(x, y, weight, h, weight_function, kernel_radius, x_pixels, y_pixels,
                              x_min, x_max, y_min, y_max, exact) -> np.ndarray:
        if exact:
            return OracleBackend._exact(_load().orc_exact_2d, x, y, weight, h, x_pixels, y_pixels,
                                        x_min, x_max, y_min, y_max)
        return OracleBackend._fast_2d(x, y, None, 0.0, weight, h, weight_function, kernel_radius,
                                      x_pixels, y_pixels, x_min, x_max, y_min, y_max, 2)

    @staticmethod
    def interpolate_2d_render_vec(x, y, weight_x, weight_y, h, weight_function, kernel_radius,
                                  x_pixels, y_pixels, x_min, x_max, y_min, y_max,
                                  exact) -> Tuple[np.ndarray, np.ndarray]:
        f = OracleBackend.interpolate_2d_render
        return (f(x, y, weight_x, h, weight_function, kernel_radius, x_pixels, y_pixels, x_min, x_max,
                  y_min, y_max, exact),
                f(x, y, weight_y, h, weight_function, kernel_radius, x_pixels, y_pixels, x_min, x_max,
                  y_min, y_max, exact))

    @staticmethod
    def interpolate_2d_line(x, y, weight, h, weight_function, kernel_radius, pixels, x1, x2, y1,
                            y2) -> np.ndarray:
        x, y, w, h = _f64(x), _f64(y), _f64(weight), _f64(h)
        ck, _keep = _ckernel(weight_function, kernel_radius, 2)
        out = np.zeros(int(pixels))
        rc = _load().orc_fast_2d_line(_p(x), _p(y), _p(w), _p(h), ctypes.c_int64(x.size),
                                      ctypes.byref(ck), int(pixels), ctypes.c_double(x1),
                                      ctypes.c_double(x2), ctypes.c_double(y1), ctypes.c_double(y2),
                                      _p(out))
        if rc:
            raise MemoryError("oracle allocation failed")
        return out

    @staticmethod
    def interpolate_3d_line(x, y, z, weight, h, weight_function, kernel_radius, pixels, x1, x2, y1,
                            y2, z1, z2) -> np.ndarray:
        x, y, z, w, h = _f64(x), _f64(y), _f64(z), _f64(weight), _f64(h)
        ck, _keep = _ckernel(weight_function, kernel_radius, 3)
        out = np.zeros(int(pixels))
        rc = _load().orc_fast_3d_line(_p(x), _p(y), _p(z), _p(w), _p(h), ctypes.c_int64(x.size),
                                      ctypes.byref(ck), int(pixels), ctypes.c_double(x1),
                                      ctypes.c_double(x2), ctypes.c_double(y1), ctypes.c_double(y2),
                                      ctypes.c_double(z1), ctypes.c_double(z2), _p(out))
        if rc:
            raise MemoryError("oracle allocation failed")
        return out

    @staticmethod
    def interpolate_3d_projection(x, y, weight, h, weight_function, kernel_radius, x_pixels,
                                  y_pixels, x_min, x_max, y_min, y_max, exact) -> np.ndarray:
        if exact:
            return OracleBackend._exact(_load().orc_exact_3d_project, x, y, weight, h, x_pixels,
                                        y_pixels, x_min, x_max, y_min, y_max)
        return OracleBackend._fast_2d(x, y, None, 0.0, weight, h, weight_function, kernel_radius,
                                      x_pixels, y_pixels, x_min, x_max, y_min, y_max, 2)

    @staticmethod
    def interpolate_3d_projection_vec(x, y, weight_x, weight_y, h, weight_function, kernel_radius,
                                      x_pixels, y_pixels, x_min, x_max, y_min, y_max,
                                      exact) -> Tuple[np.ndarray, np.ndarray]:
        f = OracleBackend.interpolate_3d_projection
        return (f(x, y, weight_x, h, weight_function, kernel_radius, x_pixels, y_pixels, x_min, x_max,
                  y_min, y_max, exact),
                f(x, y, weight_y, h, weight_function, kernel_radius, x_pixels, y_pixels, x_min, x_max,
                  y_min, y_max, exact))

    @staticmethod
    def interpolate_3d_cross(x, y, z, z_slice, weight, h, weight_function, kernel_radius, x_pixels,
                             y_pixels, x_min, x_max, y_min, y_max) -> np.ndarray:
        return OracleBackend._fast_2d(x, y, z, z_slice, weight, h, weight_function, kernel_radius,
                                      x_pixels, y_pixels, x_min, x_max, y_min, y_max, 3)

    @staticmethod
    def interpolate_3d_cross_vec(x, y, z, z_slice, weight_x, weight_y, h, weight_function,
                                 kernel_radius, x_pixels, y_pixels, x_min, x_max, y_min,
                                 y_max) -> Tuple[np.ndarray, np.ndarray]:
        f = OracleBackend.interpolate_3d_cross
        return (f(x, y, z, z_slice, weight_x, h, weight_function, kernel_radius, x_pixels, y_pixels,
                  x_min, x_max, y_min, y_max),
                f(x, y, z, z_slice, weight_y, h, weight_function, kernel_radius, x_pixels, y_pixels,
                  x_min, x_max, y_min, y_max))

    @staticmethod
    def interpolate_3d_grid(x, y, z, weight, h, weight_function, kernel_radius, x_pixels, y_pixels,
                            z_pixels, x_min, x_max, y_min, y_max, z_min, z_max) -> np.ndarray:
        x, y, z, w, h = _f64(x), _f64(y), _f64(z), _f64(weight), _f64(h)
        ck, _keep = _ckernel(weight_function, kernel_radius, 3)
        out = np.zeros((int(z_pixels), int(y_pixels), int(x_pixels)))
        cnt = ctypes.c_int64(0)
        rc = _load().orc_grid_3d(_p(x), _p(y), _p(z), _p(w), _p(h), ctypes.c_int64(x.size),
                                 ctypes.byref(ck), int(x_pixels), int(y_pixels), int(z_pixels),
                                 ctypes.c_double(x_min), ctypes.c_double(x_max), ctypes.c_double(y_min),
                                 ctypes.c_double(y_max), ctypes.c_double(z_min), ctypes.c_double(z_max),
                                 _p(out), ctypes.byref(cnt))
        if rc:
            raise MemoryError("oracle allocation failed")
        OracleBackend.last_contributions = cnt.value
        return out
