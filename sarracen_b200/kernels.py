"""Smoothing kernels: host-side handles for the device weight functions.

Mirrors the public surface of sarracen/kernels (``BaseKernel.w``,
``get_radius``, ``get_column_kernel``, ``get_column_kernel_func``;
base_kernel.py:8-99, cubic_spline.py, quartic_spline.py, quintic_spline.py).
In the reference ``kernel.w`` and the column function are numba dispatchers
that the backend calls per pixel; here they are ``WeightFunction`` handles
naming a device code path (``kind``) plus, for the column kernel, the table
that the device interpolates in.  The table itself is integrated on the GPU
(sphb_column_kernel).
"""
from typing import Optional

import numpy as np

_RADIUS = {"cubic": 2.0, "quartic": 2.5, "quintic": 3.0}
_NORM = {
    "cubic": (2 / 3, 10 / (7 * np.pi), 1 / np.pi),
    "quartic": (1 / 24, 96 / (1199 * np.pi), 1 / (20 * np.pi)),
    "quintic": (1 / 120, 7 / (478 * np.pi), 1 / (120 * np.pi)),
}


def _shape(kind: str, q):
    q = np.asarray(q, dtype=np.float64)
    if kind == "cubic":
        return np.where(q < 1, 1 - 1.5 * q ** 2 + 0.75 * q ** 3, 0.25 * np.clip(2 - q, 0, None) ** 3) * (q >= 0)
    if kind == "quartic":
        return (np.clip(2.5 - q, 0, None) ** 4 - 5 * np.clip(1.5 - q, 0, None) ** 4
                + 10 * np.clip(0.5 - q, 0, None) ** 4) * (q >= 0)
    return (np.clip(3 - q, 0, None) ** 5 - 6 * np.clip(2 - q, 0, None) ** 5
            + 15 * np.clip(1 - q, 0, None) ** 5) * (q >= 0)


class WeightFunction:
    """What the backend receives as ``weight_function``.

    kind: 'cubic' | 'quartic' | 'quintic' (analytic, evaluated on the device
    from q) or 'lut' (column-integrated table, linear interpolation).
    Calling it evaluates w(q, ndim) on the host for scalar/array q -- a
    convenience for tests and plotting, never used by the scatter path.
    """

    def __init__(self, kind: str, radius: float, lut: Optional[np.ndarray] = None):
        if kind not in ("cubic", "quartic", "quintic", "lut"):
            raise ValueError(f"unknown weight function kind '{kind}'")
        self.kind = kind
        self.radius = float(radius)
        self.lut = None if lut is None else np.ascontiguousarray(lut, dtype=np.float64)
        self._lut_device = {}

    def __call__(self, q, ndim: int):
        if self.kind == "lut":
            s = self.lut.size
            t = np.asarray(q, dtype=np.float64) * (s - 1) / self.radius
            i0 = np.clip(np.floor(t), 0, s - 1).astype(np.int64)
            i1 = np.clip(np.ceil(t), 0, s - 1).astype(np.int64)
            out = self.lut[i0] + (t - i0) * (self.lut[i1] - self.lut[i0])
        else:
            out = _NORM[self.kind][int(ndim) - 1] * _shape(self.kind, q)
        return float(out) if np.ndim(out) == 0 else out

    def lut_on(self, device):
        """The table as a float64 tensor on `device` (cached)."""
        import torch
        key = str(device)
        if key not in self._lut_device:
            self._lut_device[key] = torch.from_numpy(self.lut).to(device)
        return self._lut_device[key]


class BaseKernel:
    """A smoothing kernel (sarracen/kernels/base_kernel.py:8-99)."""
    kind = None

    def __init__(self) -> None:
        self._column_cache = {}
        self._func_cache = {}

    @classmethod
    def get_radius(cls) -> float:
        return _RADIUS[cls.kind]

    def get_column_kernel(self, samples: int = 1000) -> np.ndarray:
        """Integrate the 3-D kernel along z on the device
        (base_kernel.py:39-66): ndarray of `samples` column weights for
        q_xy = linspace(0, radius, samples)."""
        samples = int(samples)
        if samples not in self._column_cache:
            from . import ops
            self._column_cache[samples] = ops.column_kernel(self.kind, samples).cpu().numpy()
        return self._column_cache[samples]

    def get_column_kernel_func(self, samples: int) -> WeightFunction:
        """Weight function interpolating in the column table
        (base_kernel.py:68-99)."""
        samples = int(samples)
        if samples not in self._func_cache:
            self._func_cache[samples] = WeightFunction("lut", self.get_radius(),
                                                       self.get_column_kernel(samples))
        return self._func_cache[samples]


def _make(kind_name: str, doc: str):
    wf = WeightFunction(kind_name, _RADIUS[kind_name])
    return type(kind_name.capitalize() + "SplineKernel", (BaseKernel,),
                {"kind": kind_name, "w": wf, "__doc__": doc})


CubicSplineKernel = _make("cubic", "Cubic spline, radius 2 (sarracen/kernels/cubic_spline.py).")
QuarticSplineKernel = _make("quartic", "Quartic spline, radius 2.5 (sarracen/kernels/quartic_spline.py).")
QuinticSplineKernel = _make("quintic", "Quintic spline, radius 3 (sarracen/kernels/quintic_spline.py).")
