"""`ParticleFrame`: the slice of `SarracenDataFrame` the interpolation front-end needs.

The front-end (`sarracen_b200.interpolate`) only reads a handful of attributes of
the dataframe it is given: the special column labels, `params`, `kernel`,
`backend`, `get_dim()` and `centre_of_mass()`
(sarracen/sarracen_dataframe.py:130-195, :286-311, :573-760).  A real
`SarracenDataFrame` provides them; `ParticleFrame` provides the same surface for
users (and tests) that do not have Sarracen installed.  It is pandas glue, not
compute: readers, writers, plotting and unit handling are out of scope.
"""
from typing import Dict, Optional

import numpy as np
import pandas as pd

from .kernels import BaseKernel, CubicSplineKernel


class ParticleFrame(pd.DataFrame):
    _metadata = ["params", "xcol", "ycol", "zcol", "hcol", "mcol", "rhocol", "vxcol", "vycol", "vzcol",
                 "_kernel", "backend"]

    def __init__(self, data=None, params: Optional[dict] = None, *args, **kwargs):
        super().__init__(data, *args, **kwargs)
        self.params = dict(params or {})
        self.xcol = self.ycol = self.zcol = None
        self.hcol = self.mcol = self.rhocol = None
        self.vxcol = self.vycol = self.vzcol = None
        self._identify_special_columns()
        self._kernel = CubicSplineKernel()
        self.backend = "gpu"

    @property
    def _constructor(self):
        return ParticleFrame

    def _identify_special_columns(self) -> None:
        """Same detection rules as sarracen_dataframe.py:130-195."""
        cols = list(self.columns)
        self.xcol = "x" if "x" in cols else "rx" if "rx" in cols else (str(cols[0]) if cols else None)
        self.ycol = "y" if "y" in cols else "ry" if "ry" in cols else (str(cols[1]) if len(cols) > 1 else None)
        self.zcol = "z" if "z" in cols else "rz" if "rz" in cols else None
        self.hcol = "h" if "h" in cols else None
        self.mcol = "m" if "m" in cols else "mass" if "mass" in cols else None
        self.rhocol = "rho" if "rho" in cols else "density" if "density" in cols else None
        self.vxcol = "vx" if "vx" in cols else None
        self.vycol = "vy" if "vy" in cols else None
        self.vzcol = "vz" if "vz" in cols else None

    @property
    def kernel(self) -> BaseKernel:
        return self._kernel

    @kernel.setter
    def kernel(self, new_kernel: BaseKernel) -> None:
        if not isinstance(new_kernel, BaseKernel):
            raise TypeError("Kernel is an incorrect type.")
        self._kernel = new_kernel

    # ---- optional device residency ----------------------------------------------------------
    def to_device(self, device=None, dtype=None) -> "ParticleFrame":
        """Upload every numeric column once and keep it on the GPU: later `interpolate_*` calls on
        this frame skip the host -> device copy (at 10^8 particles that copy costs more than the
        kernels).  The cache is a snapshot -- call `release_device()` (or `to_device()` again)
        after changing column values."""
        import numpy as np
        import torch
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        cache = {}
        for name in self.columns:
            arr = self[name].to_numpy()
            if arr.dtype.kind not in "fiu":
                continue
            arr = np.ascontiguousarray(arr)
            if not arr.flags.writeable:
                arr = arr.copy()
            t = torch.from_numpy(arr).to(dev)
            cache[name] = t if dtype is None else t.to(dtype)
        object.__setattr__(self, "_device_cache", cache)
        return self

    def release_device(self) -> None:
        object.__setattr__(self, "_device_cache", None)

    def get_dim(self) -> int:
        return 3 if self.zcol is not None else 2

    def centre_of_mass(self) -> list:
        """sarracen_dataframe.py:286-311."""
        mass = self[self.mcol] if self.mcol in self.columns else self.params["mass"]
        com = [(self[c] * mass).sum() for c in (self.xcol, self.ycol, self.zcol)]
        inv = 1.0 / mass.sum() if isinstance(mass, pd.Series) else 1.0 / (len(self) * mass)
        return [c * inv for c in com]


def _special_columns(cols):
    """Detection rules of sarracen_dataframe.py:130-195 on a list of labels."""
    out = {}
    out["xcol"] = "x" if "x" in cols else "rx" if "rx" in cols else (str(cols[0]) if cols else None)
    out["ycol"] = "y" if "y" in cols else "ry" if "ry" in cols else (str(cols[1]) if len(cols) > 1 else None)
    out["zcol"] = "z" if "z" in cols else "rz" if "rz" in cols else None
    out["hcol"] = "h" if "h" in cols else None
    out["mcol"] = "m" if "m" in cols else "mass" if "mass" in cols else None
    out["rhocol"] = "rho" if "rho" in cols else "density" if "density" in cols else None
    for v in ("vx", "vy", "vz"):
        out[v + "col"] = v if v in cols else None
    return out


class DeviceFrame:
    """Particle columns that live on the GPU and nowhere else: the same surface as
    `ParticleFrame` for `sarracen_b200.interpolate` -- column labels, `params`, `kernel`, `backend`,
    `get_dim()`, `centre_of_mass()` -- without pandas.  `frame[name]` brings ONE column to the host as
    a pandas Series (convenience; the front-end never does that), `to_pandas()` all of them."""

    def __init__(self, columns: Dict[str, "object"], params: Optional[dict] = None):
        lengths = {int(t.numel()) for t in columns.values()}
        if len(lengths) > 1:
            raise ValueError("all columns must have the same length")
        self._device_cache = {k: v.contiguous() for k, v in columns.items()}
        self.columns = list(self._device_cache)
        self.params = dict(params or {})
        for k, v in _special_columns(self.columns).items():
            setattr(self, k, v)
        self._kernel = CubicSplineKernel()
        self.backend = "gpu"

    def __len__(self) -> int:
        return int(next(iter(self._device_cache.values())).numel()) if self._device_cache else 0

    def __contains__(self, name) -> bool:
        return name in self._device_cache

    def __getitem__(self, name):
        return pd.Series(self._device_cache[name].cpu().numpy(), name=name)

    def column_dtype(self, name) -> np.dtype:
        """dtype of a column without bringing it to the host."""
        return np.dtype(str(self._device_cache[name].dtype).replace("torch.", ""))

    @property
    def kernel(self) -> BaseKernel:
        return self._kernel

    @kernel.setter
    def kernel(self, new_kernel: BaseKernel) -> None:
        if not isinstance(new_kernel, BaseKernel):
            raise TypeError("Kernel is an incorrect type.")
        self._kernel = new_kernel

    def get_dim(self) -> int:
        return 3 if self.zcol is not None else 2

    def centre_of_mass(self) -> list:
        """sarracen_dataframe.py:286-311, evaluated on the device."""
        import torch
        c = self._device_cache
        if self.mcol in c:
            m = c[self.mcol].to(torch.float64)
            inv = 1.0 / float(m.sum())
            return [float((c[k].to(torch.float64) * m).sum()) * inv for k in (self.xcol, self.ycol, self.zcol)]
        return [float(c[k].to(torch.float64).mean()) for k in (self.xcol, self.ycol, self.zcol)]

    def to_pandas(self) -> ParticleFrame:
        return ParticleFrame({k: v.cpu().numpy() for k, v in self._device_cache.items()}, params=self.params)
