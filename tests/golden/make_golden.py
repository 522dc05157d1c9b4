"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference and numba):

    python tests/golden/make_golden.py

Every array written here is either a seeded synthetic input or the output of
the reference's own code (sarracen 1.3.1: ``CPUBackend`` at the backend
boundary, ``kernels.*`` and ``kernels.cubic_spline_exact``) on that input.
The .npz files are committed; the GPU box has no reference, so the parity
tests there compare against these files and against the oracle that is itself
pinned to them (tests/test_oracle_golden.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_shim import load_reference  # noqa: E402

load_reference()
from sarracen.interpolate import CPUBackend  # noqa: E402
from sarracen.kernels import (CubicSplineKernel, QuarticSplineKernel,  # noqa: E402
                              QuinticSplineKernel)
from sarracen.kernels.cubic_spline_exact import line_int, surface_int  # noqa: E402

KERNELS = {"cubic": CubicSplineKernel, "quartic": QuarticSplineKernel,
           "quintic": QuinticSplineKernel}


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.1f} KiB")


def particles(rng, n, lo=-0.1, hi=1.1, hmin=0.004, hmax=0.25):
    """Positions partly outside the unit view, log-uniform smoothing lengths."""
    x = rng.uniform(lo, hi, n)
    y = rng.uniform(lo, hi, n)
    z = rng.uniform(lo, hi, n)
    h = np.exp(rng.uniform(np.log(hmin), np.log(hmax), n))
    w = rng.uniform(0.1, 1.0, n) * h ** 3
    return x, y, z, h, w


def golden_kernels():
    qs = np.concatenate([np.linspace(-0.5, 3.5, 161), [0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0]])
    out = {"q": qs}
    for name, K in KERNELS.items():
        for d in (1, 2, 3):
            out[f"w_{name}_{d}"] = np.array([K.w(q, d) for q in qs])
        out[f"column_{name}_1000"] = K().get_column_kernel(1000)
        out[f"column_{name}_64"] = K().get_column_kernel(64)
        out[f"column_{name}_2"] = K().get_column_kernel(2)
        f = K().get_column_kernel_func(1000)
        out[f"columnfunc_{name}_1000"] = np.array([f(q, 3) for q in qs])
    save("kernels", **out)


def golden_sampled():
    rng = np.random.default_rng(2001)
    n = 3000
    x, y, z, h, w = particles(rng, n)
    nx, ny = 37, 23
    lim = (0.0, 1.0, 0.05, 0.9)
    out = dict(x=x, y=y, z=z, h=h, w=w, w2=rng.normal(size=n) * h ** 3, nx=nx, ny=ny,
               lim=np.array(lim), z_slice=0.45)
    for name, K in KERNELS.items():
        k = K()
        R = k.get_radius()
        out[f"render2d_{name}"] = CPUBackend.interpolate_2d_render(
            x, y, w, h, k.w, R, nx, ny, *lim, False)
        out[f"cross_{name}"] = CPUBackend.interpolate_3d_cross(
            x, y, z, out["z_slice"], w, h, k.w, R, nx, ny, *lim)
        f = k.get_column_kernel_func(1000)
        out[f"proj_{name}"] = CPUBackend.interpolate_3d_projection(
            x, y, w, h, f, R, nx, ny, *lim, False)
    k = CubicSplineKernel()
    f = k.get_column_kernel_func(1000)
    a, b = CPUBackend.interpolate_3d_projection_vec(x, y, w, out["w2"], h, f, 2, nx, ny, *lim, False)
    out["projvec_a"], out["projvec_b"] = a, b
    a, b = CPUBackend.interpolate_3d_cross_vec(x, y, z, out["z_slice"], w, out["w2"], h, k.w, 2,
                                               nx, ny, *lim)
    out["crossvec_a"], out["crossvec_b"] = a, b
    a, b = CPUBackend.interpolate_2d_render_vec(x, y, w, out["w2"], h, k.w, 2, nx, ny, *lim, False)
    out["render2dvec_a"], out["render2dvec_b"] = a, b
    # float32 inputs handed to the reference as float32
    x32, y32, z32, h32, w32 = (a.astype(np.float32) for a in (x, y, z, h, w))
    out["render2d_cubic_f32"] = CPUBackend.interpolate_2d_render(
        x32, y32, w32, h32, k.w, 2, nx, ny, *lim, False)
    out["cross_cubic_f32"] = CPUBackend.interpolate_3d_cross(
        x32, y32, z32, np.float32(out["z_slice"]), w32, h32, k.w, 2, nx, ny, *lim)
    fq = QuinticSplineKernel().get_column_kernel_func(1000)
    out["proj_quintic_f32"] = CPUBackend.interpolate_3d_projection(
        x32, y32, w32, h32, fq, 3, nx, ny, *lim, False)
    save("sampled", **out)


def golden_small_h():
    """h well below the pixel size: most particles miss every pixel centre."""
    rng = np.random.default_rng(2002)
    n = 20000
    x, y = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    h = np.full(n, 1.2 / np.sqrt(n))
    w = rng.uniform(0.5, 1.5, n) * h ** 2
    nx, ny = 64, 48
    k = CubicSplineKernel()
    img = CPUBackend.interpolate_2d_render(x, y, w, h, k.w, 2, nx, ny, 0.0, 1.0, 0.0, 1.0, False)
    save("small_h", x=x, y=y, h=h, w=w, nx=nx, ny=ny, render2d_cubic=img)


def golden_grid():
    rng = np.random.default_rng(2003)
    n = 1500
    x, y, z, h, w = particles(rng, n, hmin=0.02, hmax=0.2)
    nx, ny, nz = 14, 11, 9
    lim = (0.0, 1.0, 0.1, 0.9, 0.05, 1.0)
    out = dict(x=x, y=y, z=z, h=h, w=w, nx=nx, ny=ny, nz=nz, lim=np.array(lim))
    for name, K in KERNELS.items():
        out[f"grid_{name}"] = CPUBackend.interpolate_3d_grid(
            x, y, z, w, h, K.w, K.get_radius(), nx, ny, nz, *lim)
    save("grid", **out)


def golden_lines():
    rng = np.random.default_rng(2004)
    n = 3000
    x, y, z, h, w = particles(rng, n, hmin=0.01, hmax=0.2)
    out = dict(x=x, y=y, z=z, h=h, w=w, pixels=41)
    cuts2 = [(0.0, 1.0, 0.1, 0.8), (0.2, 0.9, 0.9, 0.3), (0.1, 0.95, 0.5, 0.5)]
    cuts3 = [(0.0, 1.0, 0.1, 0.8, 0.2, 0.7), (0.9, 0.1, 0.5, 0.5, 0.0, 1.0),
             (0.5, 0.5, 0.5, 0.5, 0.0, 1.0)]
    out["cuts2"] = np.array(cuts2)
    out["cuts3"] = np.array(cuts3)
    for name, K in KERNELS.items():
        R = K.get_radius()
        for i, c in enumerate(cuts2):
            out[f"line2d_{name}_{i}"] = CPUBackend.interpolate_2d_line(
                x, y, w, h, K.w, R, 41, *c)
        for i, c in enumerate(cuts3):
            out[f"line3d_{name}_{i}"] = CPUBackend.interpolate_3d_line(
                x, y, z, w, h, K.w, R, 41, *c)
    save("lines", **out)


def golden_exact():
    rng = np.random.default_rng(2005)
    n = 400
    nx, ny = 20, 14
    lim = (0.0, 1.0, 0.0, 0.7)
    pw = 1.0 / nx
    # centres strictly inside the image (SURVEY note N3), h from pw/20 to 6 pw
    x = rng.uniform(0.02, 0.98, n)
    y = rng.uniform(0.02, 0.68, n)
    h = np.exp(rng.uniform(np.log(pw / 20), np.log(6 * pw), n))
    w = rng.uniform(0.1, 1.0, n) * h ** 2
    k = CubicSplineKernel()
    out = dict(x=x, y=y, h=h, w=w, nx=nx, ny=ny, lim=np.array(lim))
    out["exact2d"] = CPUBackend.interpolate_2d_render(x, y, w, h, k.w, 2, nx, ny, *lim, True)
    out["exact3dproj"] = CPUBackend.interpolate_3d_projection(
        x, y, w, h, k.get_column_kernel_func(1000), 2, nx, ny, *lim, True)
    # scalar integrals at random arguments
    m = 400
    la = np.column_stack([rng.uniform(-3, 3, m), rng.uniform(-3, 3, m), rng.uniform(-3, 3, m),
                          rng.uniform(0.3, 2.0, m)])
    out["line_int_args"] = la
    out["line_int"] = np.array([line_int(*row) for row in la])
    sa = np.column_stack([rng.uniform(-2.5, 2.5, m), rng.uniform(-1, 1, m), rng.uniform(-1, 1, m),
                          rng.uniform(-1, 1, m), rng.uniform(-1, 1, m), rng.uniform(0.1, 2, m),
                          rng.uniform(0.1, 2, m), rng.uniform(0.3, 1.5, m)])
    out["surface_int_args"] = sa
    out["surface_int"] = np.array([surface_int(*row) for row in sa])
    save("exact", **out)


if __name__ == "__main__":
    golden_kernels()
    golden_sampled()
    golden_small_h()
    golden_grid()
    golden_lines()
    golden_exact()
