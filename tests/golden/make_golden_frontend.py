"""Golden vectors for the interpolation FRONT-END, from the unmodified reference.

    python tests/golden/make_golden_frontend.py      (authoring container only)

Builds two seeded particle sets (2-D and 3-D), calls the reference's public
`sarracen.interpolate.interpolate_*` functions with `backend='cpu'` for a list of
argument combinations (normalize / dens_weight / hmin / rotation / rot_origin /
corotation / axis choice / exact / default bounds and pixel counts ...), and
stores columns, the call list (JSON) and every result in
tests/golden/frontend.npz.  tests/test_gpu_frontend.py replays the same calls
through `sarracen_b200.interpolate` on the GPU.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_shim import load_reference  # noqa: E402

sar = load_reference()
from sarracen.interpolate import interpolate as ref  # noqa: E402
from sarracen.kernels import CubicSplineKernel, QuarticSplineKernel, QuinticSplineKernel  # noqa: E402

KERNELS = {"cubic": CubicSplineKernel, "quartic": QuarticSplineKernel, "quintic": QuinticSplineKernel}

rng = np.random.default_rng(3001)
n2, n3 = 1500, 2500
f2 = dict(x=rng.uniform(0, 1, n2), y=rng.uniform(0, 0.8, n2), h=np.exp(rng.uniform(np.log(0.01), np.log(0.12), n2)),
          m=rng.uniform(0.5, 1.5, n2) / n2, rho=rng.uniform(0.5, 2.0, n2), A=rng.uniform(0, 1, n2),
          vx=rng.normal(size=n2), vy=rng.normal(size=n2))
f3 = dict(x=rng.uniform(-1, 1, n3), y=rng.uniform(-0.8, 0.9, n3), z=rng.uniform(-0.5, 0.6, n3),
          h=np.exp(rng.uniform(np.log(0.03), np.log(0.3), n3)), m=rng.uniform(0.5, 1.5, n3) / n3,
          rho=rng.uniform(0.5, 2.0, n3), A=rng.uniform(0, 1, n3), vx=rng.normal(size=n3), vy=rng.normal(size=n3),
          vz=rng.normal(size=n3))

CALLS = [
    # (frame, function, args, kwargs)
    ("f2", "interpolate_2d", ["A"], dict(x_pixels=40)),
    ("f2", "interpolate_2d", ["A"], dict(x_pixels=31, y_pixels=17, normalize=False, dens_weight=True, hmin=True,
                                          xlim=[0.1, None], ylim=[None, 0.7])),
    ("f2", "interpolate_2d", ["rho"], dict(x_pixels=25, kernel="quintic")),
    ("f2", "interpolate_2d", ["A"], dict(x_pixels=20, exact=True, xlim=[0.0, 1.0], ylim=[0.0, 0.8])),
    ("f2", "interpolate_2d", ["A"], dict(x="y", y="x", x_pixels=22, normalize=False)),
    ("f2", "interpolate_2d_vec", ["vx", "vy"], dict(x_pixels=28)),
    ("f2", "interpolate_2d_vec", ["vx", "vy"], dict(x_pixels=18, normalize=False, dens_weight=True, kernel="quartic")),
    ("f2", "interpolate_2d_line", ["A"], dict(pixels=33, xlim=[0.1, 0.9], ylim=[0.2, 0.7])),
    ("f2", "interpolate_2d_line", ["A"], dict(pixels=12, hmin=True, normalize=False)),
    ("f2nomrho", "interpolate_2d", ["A"], dict(x_pixels=24)),
    ("f3", "interpolate_3d_proj", ["A"], dict(x_pixels=36)),
    ("f3", "interpolate_3d_proj", ["rho"], dict(x_pixels=30, normalize=False)),
    ("f3", "interpolate_3d_proj", ["A"], dict(x_pixels=26, rotation=[30, 20, 10], rot_origin="com")),
    ("f3", "interpolate_3d_proj", ["A"], dict(x_pixels=26, rotation=[50, -35, 70], rot_origin="midpoint",
                                               kernel="quintic", integral_samples=64)),
    ("f3", "interpolate_3d_proj", ["A"], dict(x_pixels=24, rotation=[15, 0, 40], rot_origin=[0.1, -0.2, 0.05],
                                               dens_weight=False, hmin=True)),
    ("f3", "interpolate_3d_proj", ["A"], dict(x_pixels=24, corotation=[[0.1, 0.2, 0.0], [0.6, -0.3, 0.0]])),
    ("f3", "interpolate_3d_proj", ["A"], dict(x="y", y="z", x_pixels=21)),
    ("f3", "interpolate_3d_proj", ["A"], dict(x_pixels=14, exact=True, xlim=[-1, 1], ylim=[-0.8, 0.9])),
    ("f3", "interpolate_3d_vec", ["vx", "vy", "vz"], dict(x_pixels=27, rotation=[25, 40, -15], rot_origin="com")),
    ("f3", "interpolate_3d_vec", ["vx", "vy", "vz"], dict(x_pixels=19, normalize=False, dens_weight=True)),
    ("f3", "interpolate_3d_cross", ["A"], dict(x_pixels=32)),
    ("f3", "interpolate_3d_cross", ["A"], dict(x_pixels=23, z_slice=0.1, rotation=[20, 10, 5], hmin=True,
                                                kernel="quartic")),
    ("f3", "interpolate_3d_cross_vec", ["vx", "vy", "vz"], dict(x_pixels=21, z_slice=-0.05,
                                                                rotation=[10, 20, 30])),
    ("f3", "interpolate_3d_grid", ["A"], dict(x_pixels=12)),
    ("f3", "interpolate_3d_grid", ["A"], dict(x_pixels=10, y_pixels=9, z_pixels=7, rotation=[45, 10, -20],
                                               normalize=False, dens_weight=True, zlim=[-0.4, 0.5])),
    ("f3", "interpolate_3d_line", ["A"], dict(pixels=29, xlim=[-0.8, 0.7], ylim=[-0.5, 0.6], zlim=[-0.3, 0.4])),
    ("f3", "interpolate_3d_line", ["A"], dict(pixels=15, xlim=[-0.5, 0.5], ylim=[0.0, 0.0], zlim=[0.0, 0.0],
                                               hmin=True, normalize=False)),
    ("f3nomrho", "interpolate_3d_cross", ["A"], dict(x_pixels=18)),
]


def frame(name):
    if name == "f2":
        return sar.SarracenDataFrame(f2, params={"hfact": 1.2})
    if name == "f3":
        return sar.SarracenDataFrame(f3, params={"hfact": 1.2})
    base = f2 if name.startswith("f2") else f3
    cols = {k: v for k, v in base.items() if k not in ("m", "rho")}
    return sar.SarracenDataFrame(cols, params={"hfact": 1.2, "mass": 1.0 / len(base["x"])})


out = {f"f2_{k}": v for k, v in f2.items()}
out.update({f"f3_{k}": v for k, v in f3.items()})
for i, (fr, fn, args, kw) in enumerate(CALLS):
    kwargs = dict(kw)
    if "kernel" in kwargs:
        kwargs["kernel"] = KERNELS[kwargs["kernel"]]()
    for lim in ("xlim", "ylim", "zlim"):
        if lim in kwargs:
            kwargs[lim] = tuple(kwargs[lim])
    if "corotation" in kwargs:
        kwargs["corotation"] = [list(c) for c in kwargs["corotation"]]
    res = getattr(ref, fn)(frame(fr), *args, backend="cpu", **kwargs)
    if isinstance(res, tuple):
        for j, r in enumerate(res):
            out[f"res_{i}_{j}"] = r
    else:
        out[f"res_{i}_0"] = res
    print(i, fn, kw, [r.shape for r in (res if isinstance(res, tuple) else (res,))], flush=True)
out["calls"] = np.array(json.dumps(CALLS))
path = os.path.join(HERE, "frontend.npz")
np.savez_compressed(path, **out)
print(f"frontend.npz {os.path.getsize(path) / 1024:.1f} KiB")
