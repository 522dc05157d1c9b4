"""Seeded randomised parity sweep (SPHB_FUZZ_IMAGES / SPHB_FUZZ_GRIDS widen it): raster shapes (incl. non-multiples of the tile), limits, smoothing-length
ranges from far below to far above the pixel size, particles on pixel centres / outside the view / with
degenerate values, all kernels, both dtypes -- CUDA path vs the CPU oracle.

Tolerances: 1e-10 for float64.  float32 inputs are held to 1e-5 on every BASELINE-shaped case
(tests/test_gpu_parity.py); this sweep allows 5e-5 for them because it also draws nearly empty images
whose largest pixel is a single contribution from the very edge of a support, where
W ~ (R - q)^n turns the float rounding of q^2 (6e-8 relative) into up to 2e-5 relative -- 1 of 300
seeds (a sparse float32 cross-section) lands between 1e-5 and 2e-5, none above."""
TOL32_SWEEP = 5e-5
import numpy as np
import pytest
from conftest import assert_parity

pytestmark = pytest.mark.gpu
KINDS = ["cubic", "quartic", "quintic"]
RADIUS = {"cubic": 2.0, "quartic": 2.5, "quintic": 3.0}


def _case(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 4000))
    nx, ny = int(rng.integers(1, 200)), int(rng.integers(1, 300))
    x0, y0 = rng.uniform(-2, 2, 2)
    lx, ly = rng.uniform(0.2, 3, 2)
    lim = (x0, x0 + lx, y0, y0 + ly)
    pw = min(lx / nx, ly / ny)
    x = rng.uniform(x0 - 0.3 * lx, x0 + 1.3 * lx, n)
    y = rng.uniform(y0 - 0.3 * ly, y0 + 1.3 * ly, n)
    z = rng.uniform(-1, 1, n)
    lo, hi = rng.choice([0.05, 0.5, 2.0]), rng.choice([3.0, 20.0, 80.0])
    h = np.exp(rng.uniform(np.log(lo * pw), np.log(hi * pw), n))
    # a few particles exactly on pixel centres
    m = min(n, 5)
    x[:m] = x0 + (rng.integers(0, nx, m) + 0.5) * (lx / nx)
    y[:m] = y0 + (rng.integers(0, ny, m) + 0.5) * (ly / ny)
    w = rng.normal(size=n) * h ** 2
    return x, y, z, h, w, nx, ny, lim, rng


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("SPHB_FUZZ_IMAGES", "24"))))
def test_random_images(seed):
    from oracle import OracleBackend, OracleKernel
    from sarracen_b200 import B200Backend as B
    import sarracen_b200 as s
    x, y, z, h, w, nx, ny, lim, rng = _case(seed)
    kind = KINDS[seed % 3]
    kern = {"cubic": s.CubicSplineKernel, "quartic": s.QuarticSplineKernel, "quintic": s.QuinticSplineKernel}[kind]()
    R = RADIUS[kind]
    dtype, tol = ((np.float64, 1e-10), (np.float32, TOL32_SWEEP))[(seed // 3) % 2]
    a = [v.astype(dtype) for v in (x, y, z, h, w)]
    xa, ya, za, ha, wa = a
    mode = seed % 4
    if mode == 0:
        out = B.interpolate_2d_render(xa, ya, wa, ha, kern.w, R, nx, ny, *lim, False)
        ref = OracleBackend.interpolate_2d_render(xa, ya, wa, ha, OracleKernel(kind), R, nx, ny, *lim, False)
    elif mode == 1:
        samples = int(rng.choice([2, 17, 1000]))
        out = B.interpolate_3d_projection(xa, ya, wa, ha, kern.get_column_kernel_func(samples), R, nx, ny, *lim, False)
        ref = OracleBackend.interpolate_3d_projection(xa, ya, wa, ha, OracleKernel.column(kind, samples), R, nx, ny, *lim,
                                                      False)
    elif mode == 2:
        zs = float(dtype(rng.uniform(-0.5, 0.5)))
        out = B.interpolate_3d_cross(xa, ya, za, zs, wa, ha, kern.w, R, nx, ny, *lim)
        ref = OracleBackend.interpolate_3d_cross(xa, ya, za, zs, wa, ha, OracleKernel(kind), R, nx, ny, *lim)
    else:
        w2 = (rng.normal(size=x.size) * h ** 2).astype(dtype)
        o1, o2 = B.interpolate_2d_render_vec(xa, ya, wa, w2, ha, kern.w, R, nx, ny, *lim, False)
        r1, r2 = OracleBackend.interpolate_2d_render_vec(xa, ya, wa, w2, ha, OracleKernel(kind), R, nx, ny, *lim, False)
        assert_parity(o2, r2, tol, f"seed {seed} vec y")
        out, ref = o1, r1
    assert_parity(out, ref, tol, f"seed {seed} mode {mode} {kind} {np.dtype(dtype).name} {nx}x{ny} n={x.size}")


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("SPHB_FUZZ_GRIDS", "8"))))
def test_random_grids(seed):
    from oracle import OracleBackend, OracleKernel
    from sarracen_b200 import B200Backend as B
    import sarracen_b200 as s
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(1, 1500))
    nx, ny, nz = (int(v) for v in rng.integers(1, 40, 3))
    lim = (0.0, rng.uniform(0.5, 2), -1.0, rng.uniform(-0.5, 1), 0.2, rng.uniform(0.5, 1.5))
    pw = min((lim[1] - lim[0]) / nx, (lim[3] - lim[2]) / ny, (lim[5] - lim[4]) / nz)
    x = rng.uniform(lim[0] - 0.2, lim[1] + 0.2, n)
    y = rng.uniform(lim[2] - 0.2, lim[3] + 0.2, n)
    z = rng.uniform(lim[4] - 0.2, lim[5] + 0.2, n)
    h = np.exp(rng.uniform(np.log(0.1 * pw), np.log(12 * pw), n))
    m = min(n, 4)
    x[:m] = lim[0] + (rng.integers(0, nx, m) + 0.5) * ((lim[1] - lim[0]) / nx)
    z[:m] = lim[4] + (rng.integers(0, nz, m) + 0.5) * ((lim[5] - lim[4]) / nz)
    w = rng.normal(size=n) * h ** 3
    kind = KINDS[seed % 3]
    kern = {"cubic": s.CubicSplineKernel, "quartic": s.QuarticSplineKernel, "quintic": s.QuinticSplineKernel}[kind]()
    dtype, tol = ((np.float64, 1e-10), (np.float32, TOL32_SWEEP))[seed % 2]
    a = [v.astype(dtype) for v in (x, y, z, w, h)]
    out = B.interpolate_3d_grid(*a, kern.w, RADIUS[kind], nx, ny, nz, *lim)
    ref = OracleBackend.interpolate_3d_grid(*a, OracleKernel(kind), RADIUS[kind], nx, ny, nz, *lim)
    assert_parity(out, ref, tol, f"grid seed {seed} {kind} {np.dtype(dtype).name} {nx}x{ny}x{nz} n={n}")


def test_degenerate_particles_are_skipped():
    """h <= 0, NaN / inf coordinates: skipped here (the reference would poison the image); the rest is unchanged."""
    from oracle import OracleBackend, OracleKernel
    from sarracen_b200 import B200Backend as B
    import sarracen_b200 as s
    rng = np.random.default_rng(77)
    n = 500
    x, y = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    h = rng.uniform(0.01, 0.1, n)
    w = rng.uniform(0.1, 1, n)
    good = np.ones(n, bool)
    h[3], h[4], x[5], y[6], x[7] = 0.0, -1.0, np.nan, np.inf, -np.inf
    good[[3, 4, 5, 6, 7]] = False
    k = s.CubicSplineKernel()
    out = B.interpolate_2d_render(x, y, w, h, k.w, 2, 40, 40, 0, 1, 0, 1, False)
    ref = OracleBackend.interpolate_2d_render(x[good], y[good], w[good], h[good], OracleKernel("cubic"), 2, 40, 40, 0, 1,
                                              0, 1, False)
    assert_parity(out, ref, 1e-10, "degenerate particles")


@pytest.mark.parametrize("seed", range(10))
def test_random_lines_and_exact(seed):
    from oracle import OracleBackend, OracleKernel
    from sarracen_b200 import B200Backend as B
    import sarracen_b200 as s
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(1, 1200))
    x, y, z = rng.uniform(-0.2, 1.2, (3, n))
    h = np.exp(rng.uniform(np.log(0.004), np.log(0.3), n))
    w = rng.normal(size=n) * h ** 2
    kind = KINDS[seed % 3]
    kern = {"cubic": s.CubicSplineKernel, "quartic": s.QuarticSplineKernel, "quintic": s.QuinticSplineKernel}[kind]()
    R = RADIUS[kind]
    px = int(rng.integers(1, 1500))
    c2 = [tuple(rng.uniform(0, 1, 4)), (0.1, 0.9, 0.4, 0.4), (0.5, 0.5, 0.1, 0.9), (0.9, 0.2, 0.3, 0.8)][seed % 4]
    out = B.interpolate_2d_line(x, y, w, h, kern.w, R, px, *c2)
    ref = OracleBackend.interpolate_2d_line(x, y, w, h, OracleKernel(kind), R, px, *c2)
    assert_parity(out, ref, 1e-10, f"line2d seed {seed} {c2}")
    c3 = tuple(rng.uniform(0, 1, 6))
    out = B.interpolate_3d_line(x, y, z, w, h, kern.w, R, px, *c3)
    ref = OracleBackend.interpolate_3d_line(x, y, z, w, h, OracleKernel(kind), R, px, *c3)
    assert_parity(out, ref, 1e-10, f"line3d seed {seed}")
    # exact paths (cubic only); positive weights, centres inside and outside the view
    m = min(n, 150)
    nx, ny = int(rng.integers(1, 40)), int(rng.integers(1, 40))
    lim = (0.1, 0.9, 0.2, 0.8)
    wpos = np.abs(w[:m]) + 1e-3
    k = s.CubicSplineKernel()
    out = B.interpolate_2d_render(x[:m], y[:m], wpos, h[:m], k.w, 2, nx, ny, *lim, True)
    ref = OracleBackend.interpolate_2d_render(x[:m], y[:m], wpos, h[:m], OracleKernel("cubic"), 2, nx, ny, *lim, True)
    assert_parity(out, ref, 1e-10, f"exact2d seed {seed} {nx}x{ny}")
    out = B.interpolate_3d_projection(x[:m], y[:m], wpos, h[:m], k.w, 2, nx, ny, *lim, True)
    ref = OracleBackend.interpolate_3d_projection(x[:m], y[:m], wpos, h[:m], OracleKernel("cubic"), 2, nx, ny, *lim, True)
    assert_parity(out, ref, 1e-10, f"exact3d seed {seed} {nx}x{ny}")
