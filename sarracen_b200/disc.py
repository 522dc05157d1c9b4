"""Disc analysis on the GPU: the radial-binning profiles of `sarracen.disc`
(sarracen/disc/surface_density.py:10-459, sarracen/disc/utils.py:25-105) with the same names,
arguments, defaults and return values -- `surface_density`, `azimuthal_average`,
`angular_momentum`, `scale_height`, `honH`.

The reference bins with `pandas.cut` + `groupby` on the host.  Here the particle columns stay on
the device (a `DeviceFrame`, a `ParticleFrame.to_device()` or a plain frame uploaded once), the
ring of every particle and the ring populations come from one hand-written kernel
(`sphb_radial_index`: radius, binary search on the same edge array pandas would use, right-closed
intervals) and the per-ring sums from another (`sphb_bin_sums`: shared-memory private histograms);
only the `bins`-long profiles return to the host.  There is no CPU path.
"""
import ctypes
import sys
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .interpolate import _Device


def _edges(r_min: float, r_max: float, r_in, r_out, bins: int, log: bool) -> np.ndarray:
    """utils.py:75-83."""
    if r_in is None:
        r_in = r_min - sys.float_info.epsilon
    if r_out is None:
        r_out = r_max + sys.float_info.epsilon
    if log:
        return np.logspace(np.log10(r_in), np.log10(r_out), bins + 1)
    return np.linspace(r_in, r_out, bins + 1)


def _midpoints(edges: np.ndarray, log: bool) -> np.ndarray:
    """utils.py:89-105."""
    if log:
        return np.sqrt(edges[:-1] * edges[1:])
    return 0.5 * (edges[1:] - edges[:-1]) + edges[:-1]


class _Rings:
    """Ring index of every particle (device int32, -1 = outside), populations and edges."""

    def __init__(self, data, r_in, r_out, bins: int, log: bool, geometry: str, origin: Optional[Sequence[float]]):
        if geometry not in ("cylindrical", "spherical"):
            raise ValueError("geometry should be either 'cylindrical' or 'spherical'")
        self.origin = [0.0, 0.0, 0.0] if origin is None else [float(v) for v in origin]
        self.bins = int(bins)
        spherical = geometry == "spherical"
        names = [data.xcol, data.ycol] + ([data.zcol] if spherical else [])
        self.dv = _Device(data, names)
        dv = self.dv
        x, y = dv.raw(data.xcol), dv.raw(data.ycol)
        z = dv.raw(data.zcol) if spherical else None
        if not (x.dtype == y.dtype and (z is None or z.dtype == x.dtype) and x.dtype in (torch.float32, torch.float64)):
            x, y, z = x.double(), y.double(), (z.double() if z is not None else None)
        if r_in is None or r_out is None:
            # the data range, with the radius rounded like the binning kernel (and NumPy) rounds it
            r2 = (x.double() - self.origin[0]) ** 2 + (y.double() - self.origin[1]) ** 2
            if spherical:
                r2 = r2 + (z.double() - self.origin[2]) ** 2
            r = torch.sqrt(r2)
            r_min, r_max = float(r.min()), float(r.max())
            del r, r2
        else:
            r_min = r_max = 0.0
        self.edges = _edges(r_min, r_max, r_in, r_out, self.bins, log)
        self.log = log
        dev = x.device
        self.index = torch.empty(x.numel(), dtype=torch.int32, device=dev)
        self.counts = torch.empty(self.bins, dtype=torch.int64, device=dev)
        edges_d = torch.from_numpy(self.edges).to(dev)
        org = (ctypes.c_double * 3)(*self.origin)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().sphb_radial_index(
                _lib.SPHB_F64 if x.dtype == torch.float64 else _lib.SPHB_F32, x.data_ptr(), y.data_ptr(),
                z.data_ptr() if z is not None else None, x.numel(), org, 1 if spherical else 0, edges_d.data_ptr(),
                self.bins, self.index.data_ptr(), self.counts.data_ptr(),
                ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))

    def sums(self, values: List[torch.Tensor]) -> List[torch.Tensor]:
        """Per-ring sums of up to four per-particle value columns (float64 results on the device)."""
        dev = self.index.device
        dtype = torch.float64 if any(v.dtype == torch.float64 for v in values) else torch.float32
        vals = [v.to(dtype).contiguous() for v in values]
        out = [torch.empty(self.bins, dtype=torch.float64, device=dev) for _ in vals]
        vp = (ctypes.c_void_p * len(vals))(*[v.data_ptr() for v in vals])
        op = (ctypes.c_void_p * len(vals))(*[o.data_ptr() for o in out])
        with torch.cuda.device(dev):
            _lib.check(_lib.load().sphb_bin_sums(_lib.SPHB_F64 if dtype == torch.float64 else _lib.SPHB_F32,
                                                 self.index.data_ptr(), len(vals), vp, self.index.numel(), self.bins,
                                                 op, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return out

    def midpoints(self) -> np.ndarray:
        return _midpoints(self.edges, self.log)


def _mass(data, dv: _Device):
    """utils.py:8-15: the mass column, or the global particle mass."""
    if data.mcol is None:
        if "mass" not in data.params:
            raise KeyError("'mass' column does not exist in this SarracenDataFrame.")
        return float(data.params["mass"])
    return dv.raw(data.mcol)


def azimuthal_average(data, target: str, r_in=None, r_out=None, bins: int = 300, log: bool = False,
                      geometry: str = "cylindrical", origin=None, retbins: bool = False):
    """Mean of `target` per ring (surface_density.py:10-77); empty rings are NaN."""
    rings = _Rings(data, r_in, r_out, bins, log, geometry, origin)
    total = rings.sums([rings.dv.raw(target)])[0]
    result = (total / rings.counts.double()).cpu().numpy()
    return (result, rings.midpoints()) if retbins else result


def surface_density(data, r_in=None, r_out=None, bins: int = 300, log: bool = False, geometry: str = "cylindrical",
                    origin=None, retbins: bool = False):
    """Mass per ring over the ring's area (surface_density.py:80-157)."""
    rings = _Rings(data, r_in, r_out, bins, log, geometry, origin)
    areas = np.pi * (rings.edges[1:] ** 2 - rings.edges[:-1] ** 2)
    mass = _mass(data, rings.dv)
    if isinstance(mass, torch.Tensor):
        sigma = rings.sums([mass])[0].cpu().numpy()
    else:
        sigma = rings.counts.cpu().numpy() * mass
    return (sigma / areas, rings.midpoints()) if retbins else sigma / areas


def _angular_momentum(data, rings: _Rings, unit_vector: bool):
    """surface_density.py:160-201: L = sum m (r - origin) x v per ring, on the device."""
    dv = rings.dv
    o = rings.origin
    x, y, z = (dv.raw(c).double() - o[i] for i, c in enumerate((data.xcol, data.ycol, data.zcol)))
    vx, vy, vz = (dv.raw(c).double() for c in (data.vxcol, data.vycol, data.vzcol))
    mass = _mass(data, dv)
    m = mass.double() if isinstance(mass, torch.Tensor) else mass
    lx, ly, lz = rings.sums([m * (y * vz - z * vy), m * (z * vx - x * vz), m * (x * vy - y * vx)])
    if unit_vector:
        inv = 1.0 / torch.sqrt(lx ** 2 + ly ** 2 + lz ** 2)
        lx, ly, lz = lx * inv, ly * inv, lz * inv
    return lx, ly, lz


def angular_momentum(data, r_in=None, r_out=None, bins: int = 300, log: bool = False,
                     geometry: str = "cylindrical", origin=None, retbins: bool = False, unit_vector: bool = True):
    """Angular momentum profile (surface_density.py:204-283)."""
    rings = _Rings(data, r_in, r_out, bins, log, geometry, origin)
    lx, ly, lz = (t.cpu().numpy() for t in _angular_momentum(data, rings, unit_vector))
    return (lx, ly, lz, rings.midpoints()) if retbins else (lx, ly, lz)


def _scale_height(data, rings: _Rings) -> torch.Tensor:
    """surface_density.py:286-312: per ring, the sample standard deviation of the particles' position
    along the ring's unit angular momentum vector (two passes: mean, then squared deviations)."""
    dv = rings.dv
    lx, ly, lz = _angular_momentum(data, rings, True)
    idx = rings.index.long().clamp(min=0)
    zdash = lx[idx] * dv.raw(data.xcol).double() + ly[idx] * dv.raw(data.ycol).double() \
        + lz[idx] * dv.raw(data.zcol).double()
    n = rings.counts.double()
    mean = rings.sums([zdash])[0] / n
    dev2 = rings.sums([(zdash - mean[idx]) ** 2])[0]
    return torch.sqrt(dev2 / (n - 1.0))


def scale_height(data, r_in=None, r_out=None, bins: int = 300, log: bool = False, geometry: str = "cylindrical",
                 origin=None, retbins: bool = False):
    """H / R of the disc (surface_density.py:315-390)."""
    rings = _Rings(data, r_in, r_out, bins, log, geometry, origin)
    mid = rings.midpoints()
    h = _scale_height(data, rings).cpu().numpy() / mid
    return (h, mid) if retbins else h


def honH(data, r_in=None, r_out=None, bins: int = 300, log: bool = False, geometry: str = "cylindrical",
         origin=None, retbins: bool = False):
    """<h> / H per ring (surface_density.py:393-459)."""
    rings = _Rings(data, r_in, r_out, bins, log, geometry, origin)
    big_h = _scale_height(data, rings)
    mean_h = rings.sums([rings.dv.raw(data.hcol)])[0] / rings.counts.double()
    result = (mean_h / big_h).cpu().numpy()
    return (result, rings.midpoints()) if retbins else result
