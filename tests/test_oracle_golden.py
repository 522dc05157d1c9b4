"""Pin the CPU oracle (oracle/sph_oracle.c) to the reference.

The golden vectors are outputs of the unmodified reference (sarracen 1.3.1),
written by tests/golden/make_golden.py.  The literal known-answer values are
the ones the reference's own tests assert
(sarracen/tests/kernels/test_kernels.py:31-193).
"""
import numpy as np
import pytest
from conftest import assert_parity, load_golden

from oracle import (OracleBackend, OracleKernel, column_kernel, kernel_w, line_int,
                    surface_int)

KINDS = ["cubic", "quartic", "quintic"]
RADIUS = {"cubic": 2, "quartic": 2.5, "quintic": 3}
TOL = 1e-12  # oracle vs reference: same algorithm, differences are rounding only


def test_kernel_values_match_reference():
    g = load_golden("kernels")
    for kind in KINDS:
        for d in (1, 2, 3):
            mine = np.array([kernel_w(kind, q, d) for q in g["q"]])
            np.testing.assert_allclose(mine, g[f"w_{kind}_{d}"], rtol=1e-14, atol=1e-17)


def test_kernel_closed_forms():
    # sarracen/tests/kernels/test_kernels.py:31-93
    pi = np.pi
    assert kernel_w("cubic", 0, 1) == pytest.approx(2 / 3)
    assert kernel_w("cubic", 1, 1) == pytest.approx(1 / 6)
    assert kernel_w("cubic", 0, 2) == pytest.approx(10 / (7 * pi))
    assert kernel_w("cubic", 1, 2) == pytest.approx(5 / (14 * pi))
    assert kernel_w("cubic", 0, 3) == pytest.approx(1 / pi)
    assert kernel_w("cubic", 1, 3) == pytest.approx(1 / (4 * pi))
    assert kernel_w("cubic", 2, 3) == 0 and kernel_w("cubic", 10, 3) == 0
    assert kernel_w("quartic", 0, 1) == pytest.approx(115 / 192)
    assert kernel_w("quartic", 0.5, 2) == pytest.approx(1056 / (1199 * pi))
    assert kernel_w("quartic", 1.5, 3) == pytest.approx(1 / (20 * pi))
    assert kernel_w("quartic", 2.5, 3) == 0
    assert kernel_w("quintic", 0, 1) == pytest.approx(11 / 20)
    assert kernel_w("quintic", 1, 2) == pytest.approx(182 / (478 * pi))
    assert kernel_w("quintic", 2, 3) == pytest.approx(1 / (120 * pi))
    assert kernel_w("quintic", 3, 3) == 0
    for kind in KINDS:  # out-of-bounds behaviour, test_kernels.py:182-193
        assert kernel_w(kind, -1, 3) == 0


def test_column_kernel_literals():
    # literals of sarracen/tests/kernels/test_kernels.py:123-168 (pytest.approx default rel 1e-6)
    pi = np.pi
    ck = OracleKernel.column("cubic", 10000)
    for q, v in [(0, 3 / (2 * pi)), (0.5, 0.33875339978), (1, 0.111036060968),
                 (1.5, 0.0114423169642)]:
        assert ck(q, 3) == pytest.approx(v)
    assert ck(2, 3) == 0 and ck(5, 3) == 0
    qk = OracleKernel.column("quartic", 10000)
    for q, v in [(0, 6 / (5 * pi)), (0.5, 0.288815941868), (1, 0.120120735858),
                 (1.5, 0.0233911861393), (2, 0.00116251851966)]:
        assert qk(q, 3) == pytest.approx(v)
    assert qk(2.5, 3) == 0 and qk(5, 3) == 0
    pk = OracleKernel.column("quintic", 10000)
    for q, v in [(0, 1 / pi), (0.5, 0.251567608959), (1, 0.121333261458),
                 (1.5, 0.0328632154395), (2, 0.00403036583315)]:
        assert pk(q, 3) == pytest.approx(v)
    assert pk(2.5, 3) == pytest.approx(0.0000979416858548, rel=1e-4)
    assert pk(3, 3) == 0 and pk(5, 3) == 0


def test_column_tables_match_reference():
    g = load_golden("kernels")
    for kind in KINDS:
        for s in (1000, 64, 2):
            np.testing.assert_allclose(column_kernel(kind, s), g[f"column_{kind}_{s}"],
                                       rtol=1e-13, atol=1e-15)
        f = OracleKernel("lut", g[f"column_{kind}_1000"], RADIUS[kind])
        mine = np.array([f(q, 3) for q in g["q"]])
        np.testing.assert_allclose(mine, g[f"columnfunc_{kind}_1000"], rtol=1e-13, atol=1e-16)
        # negative q reads entry 0; q >= R reads the (zero) last entry
        assert f(-1, 3) == f(0, 3) and f(RADIUS[kind] + 1, 3) == 0


@pytest.mark.parametrize("kind", KINDS)
def test_sampled_paths(kind):
    g = load_golden("sampled")
    x, y, z, h, w = (g[k] for k in "xyzhw")
    nx, ny, lim, zs = int(g["nx"]), int(g["ny"]), g["lim"], float(g["z_slice"])
    wf, R = OracleKernel(kind), RADIUS[kind]
    out = OracleBackend.interpolate_2d_render(x, y, w, h, wf, R, nx, ny, *lim, False)
    assert_parity(out, g[f"render2d_{kind}"], TOL, "render2d")
    out = OracleBackend.interpolate_3d_cross(x, y, z, zs, w, h, wf, R, nx, ny, *lim)
    assert_parity(out, g[f"cross_{kind}"], TOL, "cross")
    lut = OracleKernel.column(kind, 1000)
    out = OracleBackend.interpolate_3d_projection(x, y, w, h, lut, R, nx, ny, *lim, False)
    assert_parity(out, g[f"proj_{kind}"], TOL, "proj")


def test_vec_paths():
    g = load_golden("sampled")
    x, y, z, h, w, w2 = (g[k] for k in ("x", "y", "z", "h", "w", "w2"))
    nx, ny, lim, zs = int(g["nx"]), int(g["ny"]), g["lim"], float(g["z_slice"])
    wf, lut = OracleKernel("cubic"), OracleKernel.column("cubic", 1000)
    a, b = OracleBackend.interpolate_3d_projection_vec(x, y, w, w2, h, lut, 2, nx, ny, *lim, False)
    assert_parity(a, g["projvec_a"], TOL)
    assert_parity(b, g["projvec_b"], TOL)
    a, b = OracleBackend.interpolate_3d_cross_vec(x, y, z, zs, w, w2, h, wf, 2, nx, ny, *lim)
    assert_parity(a, g["crossvec_a"], TOL)
    assert_parity(b, g["crossvec_b"], TOL)
    a, b = OracleBackend.interpolate_2d_render_vec(x, y, w, w2, h, wf, 2, nx, ny, *lim, False)
    assert_parity(a, g["render2dvec_a"], TOL)
    assert_parity(b, g["render2dvec_b"], TOL)


def test_float32_inputs():
    """The reference handed float32 arrays: the oracle (float64 arithmetic on the
    exactly converted values) agrees far inside the 1e-5 float32 tolerance."""
    g = load_golden("sampled")
    x, y, z, h, w = (g[k].astype(np.float32) for k in "xyzhw")
    nx, ny, lim, zs = int(g["nx"]), int(g["ny"]), g["lim"], float(np.float32(g["z_slice"]))
    out = OracleBackend.interpolate_2d_render(x, y, w, h, OracleKernel("cubic"), 2, nx, ny, *lim, False)
    assert_parity(out, g["render2d_cubic_f32"], 1e-6)
    out = OracleBackend.interpolate_3d_cross(x, y, z, zs, w, h, OracleKernel("cubic"), 2, nx, ny, *lim)
    assert_parity(out, g["cross_cubic_f32"], 1e-6)
    out = OracleBackend.interpolate_3d_projection(x, y, w, h, OracleKernel.column("quintic"), 3,
                                                  nx, ny, *lim, False)
    assert_parity(out, g["proj_quintic_f32"], 1e-6)


def test_small_smoothing_lengths():
    g = load_golden("small_h")
    out = OracleBackend.interpolate_2d_render(g["x"], g["y"], g["w"], g["h"], OracleKernel("cubic"), 2,
                                              int(g["nx"]), int(g["ny"]), 0.0, 1.0, 0.0, 1.0, False)
    assert_parity(out, g["render2d_cubic"], TOL)


@pytest.mark.parametrize("kind", KINDS)
def test_grid(kind):
    g = load_golden("grid")
    out = OracleBackend.interpolate_3d_grid(g["x"], g["y"], g["z"], g["w"], g["h"], OracleKernel(kind),
                                            RADIUS[kind], int(g["nx"]), int(g["ny"]), int(g["nz"]),
                                            *g["lim"])
    assert_parity(out, g[f"grid_{kind}"], TOL)


@pytest.mark.parametrize("kind", KINDS)
def test_lines(kind):
    g = load_golden("lines")
    x, y, z, h, w = (g[k] for k in "xyzhw")
    px = int(g["pixels"])
    for i, c in enumerate(g["cuts2"]):
        out = OracleBackend.interpolate_2d_line(x, y, w, h, OracleKernel(kind), RADIUS[kind], px, *c)
        assert_parity(out, g[f"line2d_{kind}_{i}"], TOL, f"line2d {i}")
    for i, c in enumerate(g["cuts3"]):
        out = OracleBackend.interpolate_3d_line(x, y, z, w, h, OracleKernel(kind), RADIUS[kind], px, *c)
        assert_parity(out, g[f"line3d_{kind}_{i}"], TOL, f"line3d {i}")


def test_exact_scalar_integrals():
    g = load_golden("exact")
    mine = np.array([line_int(*row) for row in g["line_int_args"]])
    np.testing.assert_allclose(mine, g["line_int"], rtol=1e-11, atol=1e-14)
    mine = np.array([surface_int(*row) for row in g["surface_int_args"]])
    np.testing.assert_allclose(mine, g["surface_int"], rtol=1e-9, atol=1e-13)


def test_exact_paths():
    g = load_golden("exact")
    x, y, h, w = (g[k] for k in "xyhw")
    nx, ny, lim = int(g["nx"]), int(g["ny"]), g["lim"]
    out = OracleBackend.interpolate_2d_render(x, y, w, h, OracleKernel("cubic"), 2, nx, ny, *lim, True)
    assert_parity(out, g["exact2d"], 1e-11, "exact2d")
    out = OracleBackend.interpolate_3d_projection(x, y, w, h, OracleKernel.column("cubic"), 2, nx, ny,
                                                  *lim, True)
    assert_parity(out, g["exact3dproj"], 1e-11, "exact3dproj")
