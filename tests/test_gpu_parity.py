"""Parity of the CUDA path (through the C ABI) with the reference.

Every test calls `B200Backend` -- ctypes -> libsphb200.so -> sm_100a kernels --
and compares with (a) the committed golden outputs of the unmodified reference
CPU backend and (b) the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): relative L2 and max error of 1e-10 for
float64 inputs and 1e-5 for float32 inputs (the reference given the same
float32 arrays), per function.
"""
import numpy as np
import pytest
from conftest import assert_parity, load_golden

pytestmark = pytest.mark.gpu

TOL64 = 1e-10
TOL32 = 1e-5
KINDS = ["cubic", "quartic", "quintic"]


@pytest.fixture(scope="module")
def B():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    from sarracen_b200 import B200Backend
    return B200Backend


def _kernel(kind):
    import sarracen_b200 as s
    return {"cubic": s.CubicSplineKernel, "quartic": s.QuarticSplineKernel,
            "quintic": s.QuinticSplineKernel}[kind]()


def test_library_is_loaded():
    from sarracen_b200 import _lib
    assert _lib.load().sphb_version() == _lib.SPHB_VERSION


def test_column_table_on_device():
    g = load_golden("kernels")
    for kind in KINDS:
        for s in (1000, 64, 2):
            tab = _kernel(kind).get_column_kernel(s)
            np.testing.assert_allclose(tab, g[f"column_{kind}_{s}"], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("kind", KINDS)
def test_sampled_paths_float64(B, kind):
    g = load_golden("sampled")
    x, y, z, h, w = (g[k] for k in "xyzhw")
    nx, ny, lim, zs = int(g["nx"]), int(g["ny"]), g["lim"], float(g["z_slice"])
    k = _kernel(kind)
    R = k.get_radius()
    out = B.interpolate_2d_render(x, y, w, h, k.w, R, nx, ny, *lim, False)
    assert_parity(out, g[f"render2d_{kind}"], TOL64, "render2d")
    out = B.interpolate_3d_cross(x, y, z, zs, w, h, k.w, R, nx, ny, *lim)
    assert_parity(out, g[f"cross_{kind}"], TOL64, "cross")
    out = B.interpolate_3d_projection(x, y, w, h, k.get_column_kernel_func(1000), R, nx, ny, *lim, False)
    assert_parity(out, g[f"proj_{kind}"], TOL64, "proj")


def test_vec_paths_share_one_walk(B):
    g = load_golden("sampled")
    x, y, z, h, w, w2 = (g[k] for k in ("x", "y", "z", "h", "w", "w2"))
    nx, ny, lim, zs = int(g["nx"]), int(g["ny"]), g["lim"], float(g["z_slice"])
    k = _kernel("cubic")
    a, b = B.interpolate_3d_projection_vec(x, y, w, w2, h, k.get_column_kernel_func(1000), 2, nx, ny, *lim,
                                           False)
    assert_parity(a, g["projvec_a"], TOL64)
    assert_parity(b, g["projvec_b"], TOL64)
    a, b = B.interpolate_3d_cross_vec(x, y, z, zs, w, w2, h, k.w, 2, nx, ny, *lim)
    assert_parity(a, g["crossvec_a"], TOL64)
    assert_parity(b, g["crossvec_b"], TOL64)
    a, b = B.interpolate_2d_render_vec(x, y, w, w2, h, k.w, 2, nx, ny, *lim, False)
    assert_parity(a, g["render2dvec_a"], TOL64)
    assert_parity(b, g["render2dvec_b"], TOL64)


def test_sampled_paths_float32(B):
    g = load_golden("sampled")
    x, y, z, h, w = (g[k].astype(np.float32) for k in "xyzhw")
    nx, ny, lim, zs = int(g["nx"]), int(g["ny"]), g["lim"], float(np.float32(g["z_slice"]))
    out = B.interpolate_2d_render(x, y, w, h, _kernel("cubic").w, 2, nx, ny, *lim, False)
    assert_parity(out, g["render2d_cubic_f32"], TOL32, "render2d f32")
    out = B.interpolate_3d_cross(x, y, z, zs, w, h, _kernel("cubic").w, 2, nx, ny, *lim)
    assert_parity(out, g["cross_cubic_f32"], TOL32, "cross f32")
    out = B.interpolate_3d_projection(x, y, w, h, _kernel("quintic").get_column_kernel_func(1000), 3, nx, ny,
                                      *lim, False)
    assert_parity(out, g["proj_quintic_f32"], TOL32, "proj f32")


def test_small_smoothing_lengths(B):
    g = load_golden("small_h")
    out = B.interpolate_2d_render(g["x"], g["y"], g["w"], g["h"], _kernel("cubic").w, 2, int(g["nx"]),
                                  int(g["ny"]), 0.0, 1.0, 0.0, 1.0, False)
    assert_parity(out, g["render2d_cubic"], TOL64)


@pytest.mark.parametrize("kind", KINDS)
def test_grid(B, kind):
    g = load_golden("grid")
    k = _kernel(kind)
    out = B.interpolate_3d_grid(g["x"], g["y"], g["z"], g["w"], g["h"], k.w, k.get_radius(), int(g["nx"]),
                                int(g["ny"]), int(g["nz"]), *g["lim"])
    assert_parity(out, g[f"grid_{kind}"], TOL64)
    out = B.interpolate_3d_grid(*(g[c].astype(np.float32) for c in "xyzwh"), k.w, k.get_radius(),
                                int(g["nx"]), int(g["ny"]), int(g["nz"]), *g["lim"])
    assert_parity(out, g[f"grid_{kind}"], TOL32, "grid f32")


@pytest.mark.parametrize("kind", KINDS)
def test_lines(B, kind):
    g = load_golden("lines")
    x, y, z, h, w = (g[k] for k in "xyzhw")
    px = int(g["pixels"])
    k = _kernel(kind)
    for i, c in enumerate(g["cuts2"]):
        out = B.interpolate_2d_line(x, y, w, h, k.w, k.get_radius(), px, *c)
        assert_parity(out, g[f"line2d_{kind}_{i}"], TOL64, f"line2d {i}")
    for i, c in enumerate(g["cuts3"]):
        out = B.interpolate_3d_line(x, y, z, w, h, k.w, k.get_radius(), px, *c)
        assert_parity(out, g[f"line3d_{kind}_{i}"], TOL64, f"line3d {i}")


def test_exact_paths(B):
    g = load_golden("exact")
    x, y, h, w = (g[k] for k in "xyhw")
    nx, ny, lim = int(g["nx"]), int(g["ny"]), g["lim"]
    k = _kernel("cubic")
    out = B.interpolate_2d_render(x, y, w, h, k.w, 2, nx, ny, *lim, True)
    assert_parity(out, g["exact2d"], TOL64, "exact2d")
    out = B.interpolate_3d_projection(x, y, w, h, k.w, 2, nx, ny, *lim, True)
    assert_parity(out, g["exact3dproj"], TOL64, "exact3dproj")


# ---- larger seeded cases against the oracle (sizes it finishes in seconds) ----

def _plummer(rng, n, a=1.0, rmax=10.0):
    u = rng.uniform(0, (1 + a * a / rmax ** 2) ** -1.5, n)
    r = a / np.sqrt(u ** (-2 / 3) - 1)
    ct = rng.uniform(-1, 1, n)
    ph = rng.uniform(0, 2 * np.pi, n)
    st = np.sqrt(1 - ct * ct)
    rho = 3 / (4 * np.pi * a ** 3) * (1 + r * r / (a * a)) ** -2.5
    h = 1.2 * ((1.0 / n) / rho) ** (1 / 3)
    return r * st * np.cos(ph), r * st * np.sin(ph), r * ct, h, rho


@pytest.mark.parametrize("dtype,tol", [(np.float64, TOL64), (np.float32, TOL32)])
def test_projection_plummer_vs_oracle(B, dtype, tol):
    """BASELINE config 2 at reduced N: Plummer sphere, quintic column table."""
    from oracle import OracleBackend, OracleKernel
    rng = np.random.default_rng(1002)
    n = 20000
    x, y, z, h, rho = _plummer(rng, n)
    w = rng.uniform(0, 1, n) * (1.0 / n) / rho
    x, y, h, w = (a.astype(dtype) for a in (x, y, h, w))
    nx = ny = 256
    lim = (-2.0, 2.0, -2.0, 2.0)
    k = _kernel("quintic")
    out = B.interpolate_3d_projection(x, y, w, h, k.get_column_kernel_func(1000), 3, nx, ny, *lim, False)
    ref = OracleBackend.interpolate_3d_projection(x, y, w, h, OracleKernel.column("quintic", 1000), 3, nx, ny,
                                                  *lim, False)
    assert_parity(out, ref, tol, "plummer projection")


@pytest.mark.parametrize("dtype,tol", [(np.float64, TOL64), (np.float32, TOL32)])
def test_grid_uniform_vs_oracle(B, dtype, tol):
    """BASELINE config 4 at reduced N (same 2h / voxel = 2.44)."""
    from oracle import OracleBackend, OracleKernel
    rng = np.random.default_rng(1004)
    n = 1 << 15
    x, y, z = rng.uniform(0, 1, (3, n))
    h = np.full(n, 1.2 * n ** (-1 / 3))
    w = rng.uniform(0, 1, n) / n
    x, y, z, h, w = (a.astype(dtype) for a in (x, y, z, h, w))
    m = 32
    k = _kernel("cubic")
    out = B.interpolate_3d_grid(x, y, z, w, h, k.w, 2, m, m, m, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    ref = OracleBackend.interpolate_3d_grid(x, y, z, w, h, OracleKernel("cubic"), 2, m, m, m, 0.0, 1.0, 0.0, 1.0,
                                            0.0, 1.0)
    assert_parity(out, ref, tol, "uniform grid")


def test_single_particle_every_function(B):
    """After sarracen/tests/interpolate/test_interpolate.py:36-147: one particle,
    every pixel equals w / h^d * W(r / h)."""
    from oracle import OracleKernel, kernel_w
    k = _kernel("cubic")
    x, y, z = np.array([0.0]), np.array([0.0]), np.array([0.0])
    h, w = np.array([0.9]), np.array([0.35])
    n = 25
    lim = (-1.0, 1.0, -1.0, 1.0)
    c = -1.0 + (np.arange(n) + 0.5) * (2.0 / n)
    r2 = c[None, :] ** 2 + c[:, None] ** 2
    img = B.interpolate_2d_render(x, y, w, h, k.w, 2, n, n, *lim, False)
    exp = np.vectorize(lambda q: kernel_w("cubic", q, 2))(np.sqrt(r2) / 0.9) * 0.35 / 0.9 ** 2
    assert_parity(img, exp, 1e-11)
    col = OracleKernel.column("cubic", 1000)
    img = B.interpolate_3d_projection(x, y, w, h, k.get_column_kernel_func(1000), 2, n, n, *lim, False)
    exp = np.vectorize(lambda q: col(q, 3))(np.sqrt(r2) / 0.9) * 0.35 / 0.9 ** 2
    assert_parity(img, exp, 1e-11)
    img = B.interpolate_3d_cross(x, y, z, 0.0, w, h, k.w, 2, n, n, *lim)
    exp = np.vectorize(lambda q: kernel_w("cubic", q, 3))(np.sqrt(r2) / 0.9) * 0.35 / 0.9 ** 3
    assert_parity(img, exp, 1e-11)
    grid = B.interpolate_3d_grid(x, y, z, w, h, k.w, 2, n, n, n, *lim, -1.0, 1.0)
    r3 = np.sqrt(r2[None, :, :] + c[:, None, None] ** 2)
    exp = np.vectorize(lambda q: kernel_w("cubic", q, 3))(r3 / 0.9) * 0.35 / 0.9 ** 3
    assert_parity(grid, exp, 1e-11)


def test_repeated_particle_has_no_races(B):
    """After test_interpolate.py:150-270: the same particle 10 000 times."""
    k = _kernel("quintic")
    n = 10000
    x, y = np.zeros(n), np.zeros(n)
    h, w = np.full(n, 0.7), np.full(n, 1e-4)
    one = B.interpolate_2d_render(x[:1], y[:1], w[:1], h[:1], k.w, 3, 31, 29, -1, 1, -1, 1, False)
    many = B.interpolate_2d_render(x, y, w, h, k.w, 3, 31, 29, -1, 1, -1, 1, False)
    np.testing.assert_allclose(many, one * n, rtol=1e-11)


def test_empty_and_out_of_view(B):
    k = _kernel("cubic")
    e = np.zeros(0)
    assert not B.interpolate_2d_render(e, e, e, e, k.w, 2, 8, 6, 0, 1, 0, 1, False).any()
    assert B.interpolate_3d_grid(e, e, e, e, e, k.w, 2, 4, 5, 6, 0, 1, 0, 1, 0, 1).shape == (6, 5, 4)
    far = np.array([50.0])
    img = B.interpolate_2d_render(far, far, np.ones(1), np.array([0.1]), k.w, 2, 8, 6, 0, 1, 0, 1, False)
    assert img.shape == (6, 8) and not img.any()


def test_four_weights_in_one_walk_and_accumulate(B):
    """ops level: up to 4 weight columns per footprint walk; accumulate adds to the raster."""
    import torch
    from sarracen_b200 import ops
    from oracle import OracleBackend, OracleKernel
    rng = np.random.default_rng(11)
    n = 4000
    x, y = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    h = np.exp(rng.uniform(np.log(0.004), np.log(0.2), n))
    ws = [rng.normal(size=n) * h ** 2 for _ in range(4)]
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    r = ops.Raster(70, 45, 0.0, 1.0, 0.1, 0.8)
    k = _kernel("quartic")
    outs = ops.scatter2d(t(x), t(y), t(h), [t(w) for w in ws], k.w, 2.5, 2, r)
    for w, o in zip(ws, outs):
        ref = OracleBackend.interpolate_2d_render(x, y, w, h, OracleKernel("quartic"), 2.5, 70, 45, 0.0, 1.0, 0.1, 0.8,
                                                  False)
        assert_parity(o.cpu().numpy(), ref, TOL64, "4 weights")
    # second call accumulates into the first output
    ops.scatter2d(t(x), t(y), t(h), [t(ws[1])], k.w, 2.5, 2, r, out=[outs[0]], accumulate=True)
    ref = OracleBackend.interpolate_2d_render(x, y, ws[0] + ws[1], h, OracleKernel("quartic"), 2.5, 70, 45, 0.0, 1.0,
                                              0.1, 0.8, False)
    assert_parity(outs[0].cpu().numpy(), ref, TOL64, "accumulate")


def test_particle_set_is_split_when_the_bin_list_does_not_fit(B, monkeypatch):
    """A workspace budget too small for the whole bin list forces batches that accumulate."""
    from sarracen_b200 import ops
    from oracle import OracleBackend, OracleKernel
    rng = np.random.default_rng(12)
    n = 30000
    x, y, z = rng.uniform(0, 1, (3, n))
    h = np.exp(rng.uniform(np.log(0.01), np.log(0.15), n))
    w = rng.uniform(0.1, 1, n) * h ** 3
    k = _kernel("cubic")
    ops.release_workspaces()
    monkeypatch.setattr(ops, "max_workspace_fraction", 4e-6)   # ~ 0.7 MB of a 180 GB device
    out = B.interpolate_3d_grid(x, y, z, w, h, k.w, 2, 40, 36, 28, 0, 1, 0, 1, 0, 1)
    monkeypatch.undo()
    ops.release_workspaces()
    ref = OracleBackend.interpolate_3d_grid(x, y, z, w, h, OracleKernel("cubic"), 2, 40, 36, 28, 0, 1, 0, 1, 0, 1)
    assert_parity(out, ref, TOL64, "split batches")


def test_mixed_footprints_use_both_paths(B):
    """h from far below to far above the pixel size: direct path and tile path in one call."""
    from sarracen_b200 import ops
    from oracle import OracleBackend, OracleKernel
    rng = np.random.default_rng(13)
    n = 20000
    x, y = rng.uniform(-0.05, 1.05, n), rng.uniform(-0.05, 1.05, n)
    pw = 1.0 / 300
    h = np.exp(rng.uniform(np.log(pw / 10), np.log(40 * pw), n))
    w = rng.uniform(0.1, 1, n) * h ** 2
    k = _kernel("quintic")
    for dtype, tol in ((np.float64, TOL64), (np.float32, TOL32)):
        a = [v.astype(dtype) for v in (x, y, w, h)]
        out = B.interpolate_2d_render(*a, k.w, 3, 300, 217, 0, 1, 0, 1, False)
        assert ops.last_stats["n_direct"] > 0 and ops.last_stats["n_pairs"] > 0
        ref = OracleBackend.interpolate_2d_render(*a, OracleKernel("quintic"), 3, 300, 217, 0, 1, 0, 1, False)
        assert_parity(out, ref, tol, f"mixed footprints {np.dtype(dtype).name}")


@pytest.mark.parametrize("samples", [2, 3000, 10000])
def test_column_table_sizes(B, samples):
    """integral_samples is a user argument (base_kernel.py:68): 2 (the reference's
    test_column_samples), and tables too large for shared memory (read from global memory)."""
    from oracle import OracleBackend, OracleKernel
    rng = np.random.default_rng(21)
    n = 3000
    x, y = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    h = np.exp(rng.uniform(np.log(0.003), np.log(0.2), n))
    w = rng.uniform(0.1, 1, n) * h ** 2
    k = _kernel("cubic")
    out = B.interpolate_3d_projection(x, y, w, h, k.get_column_kernel_func(samples), 2, 90, 70, 0, 1, 0, 1, False)
    ref = OracleBackend.interpolate_3d_projection(x, y, w, h, OracleKernel.column("cubic", samples), 2, 90, 70, 0, 1,
                                                  0, 1, False)
    assert_parity(out, ref, TOL64, f"column table S={samples}")
    out = B.interpolate_3d_projection(*(a.astype(np.float32) for a in (x, y, w, h)),
                                      k.get_column_kernel_func(samples), 2, 90, 70, 0, 1, 0, 1, False)
    assert_parity(out, ref, TOL32, f"column table S={samples} f32")


@pytest.mark.parametrize("n", [25, 31, 63, 127, 129])
def test_particle_exactly_on_a_pixel_centre(B, n):
    """q = 0 at a pixel centre, reached through the row recurrences of q^2 (odd rasters put the
    centre pixel on the particle): the image must stay finite and match the oracle."""
    from oracle import OracleBackend, OracleKernel
    x, y, z = np.array([0.0, 0.25]), np.array([0.0, -0.5]), np.array([0.0, 0.0])
    h, w = np.array([0.9, 0.3]), np.array([0.35, 0.2])
    for dtype, tol in ((np.float64, TOL64), (np.float32, TOL32)):
        a = [v.astype(dtype) for v in (x, y, w, h)]
        for kind in KINDS:
            k = _kernel(kind)
            out = B.interpolate_2d_render(*a, k.w, k.get_radius(), n, n, -1.0, 1.0, -1.0, 1.0, False)
            ref = OracleBackend.interpolate_2d_render(*a, OracleKernel(kind), k.get_radius(), n, n, -1.0, 1.0, -1.0,
                                                      1.0, False)
            assert_parity(out, ref, tol, f"{kind} {n} {np.dtype(dtype).name}")
        k = _kernel("quintic")
        out = B.interpolate_3d_projection(*a, k.get_column_kernel_func(1000), 3, n, n, -1.0, 1.0, -1.0, 1.0, False)
        ref = OracleBackend.interpolate_3d_projection(*a, OracleKernel.column("quintic"), 3, n, n, -1.0, 1.0, -1.0, 1.0,
                                                      False)
        assert_parity(out, ref, tol, f"column {n} {np.dtype(dtype).name}")
    if n <= 31:
        k = _kernel("cubic")
        out = B.interpolate_3d_grid(x, y, z, w, h, k.w, 2, n, n, n, -1.0, 1.0, -1.0, 1.0, -1.0, 1.0)
        ref = OracleBackend.interpolate_3d_grid(x, y, z, w, h, OracleKernel("cubic"), 2, n, n, n, -1.0, 1.0, -1.0, 1.0,
                                                -1.0, 1.0)
        assert_parity(out, ref, TOL64, f"grid {n}")


def test_backend_accepts_what_the_front_end_passes(B):
    """pandas Series, column views of an N x 3 array (non-contiguous), integer weight
    columns and a float32 h next to float64 positions all reach the backend."""
    import pandas as pd
    from oracle import OracleBackend, OracleKernel
    rng = np.random.default_rng(31)
    n = 800
    pos = rng.uniform(0, 1, (n, 3))
    x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]           # strided views
    h32 = rng.uniform(0.02, 0.1, n).astype(np.float32)
    wi = rng.integers(1, 5, n)                           # integer weights
    k = _kernel("cubic")
    out = B.interpolate_3d_cross(pd.Series(x), pd.Series(y), z, 0.5, wi, pd.Series(h32), k.w, 2, 33, 21, 0, 1, 0, 1)
    ref = OracleBackend.interpolate_3d_cross(x, y, z, 0.5, wi.astype(float), h32.astype(float), OracleKernel("cubic"),
                                             2, 33, 21, 0, 1, 0, 1)
    assert out.dtype == np.float64 and out.flags["C_CONTIGUOUS"]
    assert_parity(out, ref, TOL64, "mixed inputs")
    assert x.base is pos and not x.flags["C_CONTIGUOUS"]  # inputs untouched


def test_streamed_uploads_match_single_upload(B, monkeypatch):
    """Large host inputs go up in double-buffered chunks that accumulate into one raster
    (backend.streamed_scatter); force the chunked path on a small set -- ragged last chunk, mixed
    dtypes, every sampled method -- and compare with the single-upload result and the oracle."""
    import torch
    from sarracen_b200 import backend
    from oracle import OracleBackend, OracleKernel
    rng = np.random.default_rng(77)
    n = 5003
    x, y, z = rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    h = rng.uniform(0.01, 0.06, n)
    w, w2 = rng.uniform(0.5, 1.5, n) / n, rng.normal(size=n) / n
    k = _kernel("cubic")
    col = k.get_column_kernel_func(1000)
    lim = (0.0, 1.0, 0.0, 1.0)

    def run_all():
        return [
            B.interpolate_2d_render(x, y, w, h, k.w, 2.0, 96, 80, *lim, False),
            *B.interpolate_2d_render_vec(x, y, w, w2, h, k.w, 2.0, 96, 80, *lim, False),
            B.interpolate_3d_projection(x, y, w, h, col, 2.0, 96, 80, *lim, False),
            *B.interpolate_3d_projection_vec(x, y, w, w2, h, col, 2.0, 96, 80, *lim, False),
            B.interpolate_3d_cross(x, y, z, 0.5, w, h, k.w, 2.0, 96, 80, *lim),
            *B.interpolate_3d_cross_vec(x, y, z, 0.5, w, w2, h, k.w, 2.0, 96, 80, *lim),
            B.interpolate_3d_grid(x, y, z, w, h, k.w, 2.0, 24, 20, 16, *lim, 0.0, 1.0),
        ]

    single = run_all()
    monkeypatch.setattr(backend, "_STREAM_MIN", 1000)
    monkeypatch.setattr(backend, "_CHUNK_FIRST_MIN", 300)
    monkeypatch.setattr(backend, "_CHUNK_MAX", 1500)
    assert backend._chunk_bounds(n) == [(0, 312), (312, 1560), (1560, 3060), (3060, 4560), (4560, 5003)]
    calls = []
    real = backend.ops.scatter2d
    monkeypatch.setattr(backend.ops, "scatter2d", lambda *a, **kw: (calls.append(kw.get("accumulate")), real(*a, **kw))[1])
    chunked = run_all()
    assert calls[:5] == [False] + [True] * 4   # the first call of a method clears, later chunks accumulate
    for a, b in zip(single, chunked):
        assert_parity(b, a, 1e-13, "streamed vs single upload")
    ref = OracleBackend.interpolate_3d_grid(x, y, z, w, h, OracleKernel("cubic"), 2.0, 24, 20, 16, *lim, 0.0, 1.0)
    assert_parity(chunked[-1], ref, TOL64, "streamed grid vs oracle")
    # mixed dtypes (float32 positions, int weights) and pinned torch columns take the same path
    xf, yf = x.astype(np.float32), y.astype(np.float32)
    yf.flags.writeable = False   # pandas hands out read-only columns under copy-on-write
    wi = np.ones(n, dtype=np.int64)
    a = B.interpolate_2d_render(torch.from_numpy(xf).pin_memory(), yf, wi, h, k.w, 2.0, 96, 80, *lim, False)
    monkeypatch.setattr(backend, "_STREAM_MIN", 1 << 30)
    b = B.interpolate_2d_render(xf, yf, wi, h, k.w, 2.0, 96, 80, *lim, False)
    assert_parity(a, b, 1e-13, "streamed mixed dtypes")


@pytest.mark.parametrize("samples", [2048, 2049, 6000])
@pytest.mark.parametrize("dtype,tol", [(np.float64, TOL64), (np.float32, TOL32)])
def test_large_column_tables(B, samples, dtype, tol):
    """Column tables above 2048 samples do not fit the render kernels' shared memory and are read
    from global memory (a separate kernel variant); 2048 is the largest shared-memory table.  Mixed
    footprints so that both the tile path and the direct path look the table up."""
    from oracle import OracleBackend, OracleKernel
    rng = np.random.default_rng(2048 + samples)
    n = 6000
    x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    h = np.exp(rng.uniform(np.log(0.004), np.log(0.3), n))
    w, w2 = rng.uniform(0, 1, n) / n, rng.normal(size=n) / n
    x, y, h, w, w2 = (a.astype(dtype) for a in (x, y, h, w, w2))
    lim = (-1.0, 1.0, -1.0, 1.0)
    k = _kernel("cubic")
    out = B.interpolate_3d_projection(x, y, w, h, k.get_column_kernel_func(samples), 2, 200, 160, *lim, False)
    okern = OracleKernel.column("cubic", samples)
    ref = OracleBackend.interpolate_3d_projection(x, y, w, h, okern, 2, 200, 160, *lim, False)
    assert_parity(out, ref, tol, f"projection, {samples}-sample table")
    a, b = B.interpolate_3d_projection_vec(x, y, w, w2, h, k.get_column_kernel_func(samples), 2, 200, 160, *lim, False)
    ra, rb = OracleBackend.interpolate_3d_projection_vec(x, y, w, w2, h, okern, 2, 200, 160, *lim, False)
    assert_parity(a, ra, tol, "vec x")
    assert_parity(b, rb, tol, "vec y")


@pytest.mark.parametrize("nw", [2, 3, 4])
@pytest.mark.parametrize("dtype,tol", [(np.float64, TOL64), (np.float32, TOL32)])
def test_multi_weight_tiles_across_a_tall_raster(B, nw, dtype, tol):
    """Weight counts pick different tile shapes and CTA residencies (float64: 32x128 tiles up to two
    weights, 32x64 for three and four); a raster several tiles tall and wide, wide and narrow
    footprints, analytic kernel and column table."""
    import torch
    from sarracen_b200 import ops
    from oracle import OracleBackend, OracleKernel
    rng = np.random.default_rng(300 + nw)
    n = 3000
    x, y = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    h = np.exp(rng.uniform(np.log(0.002), np.log(0.12), n))
    ws = [rng.normal(size=n) * h ** 2 for _ in range(nw)]
    x, y, h = (a.astype(dtype) for a in (x, y, h))
    ws = [w.astype(dtype) for w in ws]
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    nx, ny = 150, 330
    r = ops.Raster(nx, ny, 0.0, 1.0, 0.0, 1.0)
    k = _kernel("quintic")
    for wf, okern in ((k.w, OracleKernel("quintic")), (k.get_column_kernel_func(1000), OracleKernel.column("quintic", 1000))):
        outs = ops.scatter2d(t(x), t(y), t(h), [t(w) for w in ws], wf, 3.0, 2, r)
        for w, o in zip(ws, outs):
            ref = OracleBackend.interpolate_2d_render(x, y, w, h, okern, 3.0, nx, ny, 0.0, 1.0, 0.0, 1.0, False)
            assert_parity(o.cpu().numpy(), ref, tol, f"{nw} weights")


@pytest.mark.parametrize("dtype,tol", [(np.float64, TOL64), (np.float32, TOL32)])
def test_extreme_footprints_and_degenerate_rasters(B, dtype, tol):
    """Footprints far larger than the raster (every tile of every particle), far smaller than a
    pixel, rasters of one pixel, one column and one row, a one-plane voxel grid."""
    from oracle import OracleBackend, OracleKernel
    rng = np.random.default_rng(99)
    n = 40
    x, y, z = rng.uniform(-0.2, 1.2, n), rng.uniform(-0.2, 1.2, n), rng.uniform(0, 1, n)
    h = np.concatenate([rng.uniform(3.0, 12.0, n // 2), rng.uniform(1e-5, 1e-3, n - n // 2)])
    w = rng.uniform(0.5, 1.5, n)
    x, y, z, h, w = (a.astype(dtype) for a in (x, y, z, h, w))
    k = _kernel("cubic")
    ok = OracleKernel("cubic")
    for nx, ny in ((1, 1), (1, 257), (300, 1), (333, 141)):
        out = B.interpolate_2d_render(x, y, w, h, k.w, 2.0, nx, ny, 0.0, 1.0, 0.0, 1.0, False)
        ref = OracleBackend.interpolate_2d_render(x, y, w, h, ok, 2.0, nx, ny, 0.0, 1.0, 0.0, 1.0, False)
        assert out.shape == (ny, nx)
        assert_parity(out, ref, tol, f"{nx}x{ny} raster")
    out = B.interpolate_3d_cross(x, y, z, 0.5, w, h, k.w, 2.0, 65, 33, 0.0, 1.0, 0.0, 1.0)
    ref = OracleBackend.interpolate_3d_cross(x, y, z, 0.5, w, h, ok, 2.0, 65, 33, 0.0, 1.0, 0.0, 1.0)
    assert_parity(out, ref, tol, "cross-section")
    for nz in (1, 19):
        out = B.interpolate_3d_grid(x, y, z, w, h, k.w, 2.0, 40, 23, nz, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
        ref = OracleBackend.interpolate_3d_grid(x, y, z, w, h, ok, 2.0, 40, 23, nz, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
        assert out.shape == (nz, 23, 40)
        assert_parity(out, ref, tol, f"grid with {nz} planes")
