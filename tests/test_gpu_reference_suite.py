"""The reference's own interpolation tests, restated against the B200 path.

sarracen/tests/interpolate/test_interpolate.py and test_rotation.py pin the hot path with
analytic known-answer tests on a few hand-placed particles.  They are parametrised over
backend in ['cpu', 'gpu']; the GPU box has no Sarracen, so the same checks are restated here
against `sarracen_b200.interpolate` on a `ParticleFrame` (expected values come from the
oracle's kernel functions, as the reference computes them from `kernel.w` in-test).
Each test names the reference test it follows.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import sarracen_b200 as s
    from oracle import OracleKernel, kernel_w
    from sarracen_b200 import interpolate as f

    class Env:
        pass
    e = Env()
    e.s, e.f, e.w = s, f, kernel_w
    e.col = OracleKernel.column("cubic", 1000)
    return e


def _frame(env, cols):
    return env.s.ParticleFrame({k: np.asarray(v, dtype=float) for k, v in cols.items()}, params={})


def test_3d_xsec_equivalency(env):
    """test_interpolate.py:309-364: a column integration equals the Riemann sum of cross-sections."""
    f = env.f
    sdf = _frame(env, dict(x=[0], y=[0], z=[0], A=[4], B=[6], C=[2], h=[0.9], rho=[0.4], m=[0.03]))
    kw = dict(x_pixels=50, xlim=(-1, 1), ylim=(-1, 1), normalize=False, hmin=False)
    column = f.interpolate_3d_proj(sdf, "A", dens_weight=False, **kw)
    column_vec = f.interpolate_3d_vec(sdf, "A", "B", "C", dens_weight=False, **kw)
    samples = 250
    acc, accv = np.zeros((50, 50)), [np.zeros((50, 50)), np.zeros((50, 50))]
    for z in np.linspace(0, 2 * 0.9, samples):
        acc += f.interpolate_3d_cross(sdf, "A", z_slice=z, **kw)
        v = f.interpolate_3d_cross_vec(sdf, "A", "B", "C", z_slice=z, **kw)
        accv[0] += v[0]
        accv[1] += v[1]
    scale = 2 * 0.9 * 2 / samples
    assert_allclose(acc * scale, column, rtol=1e-3, atol=1e-4)
    assert_allclose(accv[0] * scale, column_vec[0], rtol=1e-3, atol=1e-4)
    assert_allclose(accv[1] * scale, column_vec[1], rtol=1e-3, atol=1e-4)


def test_2d_xsec_equivalency(env):
    """test_interpolate.py:367-397: the rows of a 2-D image are horizontal line cuts."""
    f = env.f
    sdf = _frame(env, dict(x=[0], y=[0], A=[4], h=[0.9], rho=[0.4], m=[0.03]))
    img = f.interpolate_2d(sdf, "A", x_pixels=50, xlim=(-1, 1), ylim=(-1, 1), normalize=False, hmin=False)
    real = -1 + (np.arange(50) + 0.5) * (2 / 50)
    recon = np.zeros((50, 50))
    for j in range(50):
        recon[j, :] = f.interpolate_2d_line(sdf, "A", pixels=50, xlim=(-1, 1), ylim=(real[j], real[j]),
                                            normalize=False, hmin=False)
    assert_allclose(recon, img, rtol=1e-9, atol=1e-14)


def test_corner_particles(env):
    """test_interpolate.py:406-525: two particles superpose; default bounds are the data min / max."""
    f, w = env.f, env.w
    s2 = _frame(env, dict(x=[-1, 1], y=[-1, 1], A=[2, 1.5], B=[5, 2.3], h=[1.1, 1.3], rho=[0.55, 0.45], m=[0.04, 0.05]))
    s3 = _frame(env, dict(x=[-1, 1], y=[-1, 1], z=[-1, 1], A=[2, 1.5], B=[2, 1], C=[7, 8], h=[1.1, 1.3],
                          rho=[0.55, 0.45], m=[0.04, 0.05]))
    h, m, rho = np.array([1.1, 1.3]), np.array([0.04, 0.05]), np.array([0.55, 0.45])
    A, B = np.array([2, 1.5]), np.array([5, 2.3])
    w2 = m / (rho * h ** 2)
    real = (np.arange(25) + 0.5) * (2 / 25)
    r0 = np.sqrt(real[None, :] ** 2 + real[:, None] ** 2)
    r1 = np.sqrt(real[None, ::-1] ** 2 + real[::-1, None] ** 2)
    k2 = np.vectorize(lambda q: w("cubic", q, 2))
    img = f.interpolate_2d(s2, "A", x_pixels=25, y_pixels=25, normalize=False, hmin=False)
    exp = w2[0] * A[0] * k2(r0 / h[0]) + w2[1] * A[1] * k2(r1 / h[1])
    assert_allclose(img, exp, rtol=1e-6, atol=1e-12)
    ivx, ivy = f.interpolate_2d_vec(s2, "A", "B", x_pixels=25, y_pixels=25, normalize=False, hmin=False)
    assert_allclose(ivx, exp, rtol=1e-6, atol=1e-12)
    assert_allclose(ivy, w2[0] * B[0] * k2(r0 / h[0]) + w2[1] * B[1] * k2(r1 / h[1]), rtol=1e-6, atol=1e-12)
    line = f.interpolate_2d_line(s2, "A", pixels=25, normalize=False, hmin=False)
    d0, d1 = np.sqrt(2) * real, np.sqrt(2) * real[::-1]
    assert_allclose(line, w2[0] * A[0] * k2(d0 / h[0]) + w2[1] * A[1] * k2(d1 / h[1]), rtol=1e-6, atol=1e-12)
    colf = np.vectorize(lambda q: env.col(q, 3))
    img = f.interpolate_3d_proj(s3, "A", x_pixels=25, y_pixels=25, dens_weight=False, normalize=False, hmin=False)
    assert_allclose(img, w2[0] * A[0] * colf(r0 / h[0]) + w2[1] * A[1] * colf(r1 / h[1]), rtol=1e-6, atol=1e-12)
    w3 = m / (rho * h ** 3)
    k3 = np.vectorize(lambda q: w("cubic", q, 3))
    grid = f.interpolate_3d_grid(s3, "A", x_pixels=25, y_pixels=25, z_pixels=25, normalize=False, hmin=False)
    R0 = np.sqrt(r0[None, :, :] ** 2 + real[:, None, None] ** 2)
    R1 = np.sqrt(r1[None, :, :] ** 2 + real[::-1, None, None] ** 2)
    assert_allclose(grid, w3[0] * A[0] * k3(R0 / h[0]) + w3[1] * A[1] * k3(R1 / h[1]), rtol=1e-6, atol=1e-12)


def test_image_transpose(env):
    """test_interpolate.py:528-606: swapping the x / y columns transposes the image."""
    f = env.f
    s2 = _frame(env, dict(x=[-1, 1], y=[1, -1], A=[2, 1.5], B=[5, 4], h=[1.1, 1.3], rho=[0.55, 0.45], m=[0.04, 0.05]))
    kw = dict(x_pixels=20, y_pixels=20, normalize=False, hmin=False)
    assert_allclose(f.interpolate_2d(s2, "A", **kw), f.interpolate_2d(s2, "A", x="y", y="x", **kw).T, rtol=1e-12)
    a = f.interpolate_2d_vec(s2, "A", "B", **kw)
    b = f.interpolate_2d_vec(s2, "A", "B", x="y", y="x", **kw)
    assert_allclose(a[0], b[0].T, rtol=1e-12)
    assert_allclose(a[1], b[1].T, rtol=1e-12)
    s3 = _frame(env, dict(x=[-1, 1], y=[1, -1], z=[1, -1], A=[2, 1.5], B=[5, 4], C=[2.5, 3], h=[1.1, 1.3],
                          rho=[0.55, 0.45], m=[0.04, 0.05]))
    assert_allclose(f.interpolate_3d_proj(s3, "A", **kw), f.interpolate_3d_proj(s3, "A", x="y", y="x", **kw).T,
                    rtol=1e-12)
    assert_allclose(f.interpolate_3d_cross(s3, "A", **kw), f.interpolate_3d_cross(s3, "A", x="y", y="x", **kw).T,
                    rtol=1e-12)
    g1 = f.interpolate_3d_grid(s3, "A", x_pixels=12, y_pixels=12, z_pixels=12, normalize=False, hmin=False)
    g2 = f.interpolate_3d_grid(s3, "A", x="y", y="x", x_pixels=12, y_pixels=12, z_pixels=12, normalize=False,
                               hmin=False)
    assert_allclose(g1, g2.transpose(0, 2, 1), rtol=1e-12)


def test_default_kernel_and_column_samples(env):
    """test_interpolate.py:609-707: kernel kwarg vs frame kernel on a 1x1 image; integral_samples honoured."""
    s, f, w = env.s, env.f, env.w
    from oracle import column_kernel
    s2 = _frame(env, dict(x=[0], y=[0], A=[1], h=[1], rho=[1], m=[1]))
    s3 = _frame(env, dict(x=[0], y=[0], z=[0], A=[1], h=[1], rho=[1], m=[1]))
    kw = dict(x_pixels=1, y_pixels=1, xlim=(-1, 1), ylim=(-1, 1), normalize=False, hmin=False)
    s2.kernel = s.QuarticSplineKernel()
    s3.kernel = s.QuinticSplineKernel()
    assert f.interpolate_2d(s2, "A", **kw)[0, 0] == pytest.approx(w("quartic", 0, 2), rel=1e-12)
    assert f.interpolate_2d(s2, "A", kernel=s.QuinticSplineKernel(), **kw)[0, 0] == pytest.approx(
        w("quintic", 0, 2), rel=1e-12)
    assert f.interpolate_3d_cross(s3, "A", z_slice=0, **kw)[0, 0] == pytest.approx(w("quintic", 0, 3), rel=1e-12)
    assert f.interpolate_3d_grid(s3, "A", x_pixels=1, y_pixels=1, z_pixels=1, xlim=(-1, 1), ylim=(-1, 1),
                                 zlim=(-1, 1), normalize=False, hmin=False)[0, 0, 0] == pytest.approx(
        w("quintic", 0, 3), rel=1e-12)
    img = f.interpolate_3d_proj(s3, "A", integral_samples=2, dens_weight=False, **kw)
    assert img[0, 0] == pytest.approx(column_kernel("quintic", 2)[0], rel=1e-12)


def test_pixel_arguments_and_shapes(env):
    """test_interpolate.py:713-908: default pixel counts keep the aspect ratio; shapes are (ny, nx) / (nz, ny, nx)."""
    f = env.f
    s2 = _frame(env, dict(x=[-2, 4], y=[3, 7], A=[1, 1], B=[1, 1], h=[0.1, 0.1], rho=[1, 1], m=[1, 1]))
    s3 = _frame(env, dict(x=[-2, 4], y=[3, 7], z=[6, -2], A=[1, 1], B=[1, 1], C=[1, 1], h=[0.1, 0.1], rho=[1, 1],
                          m=[1, 1]))
    assert f.interpolate_2d(s2, "A").shape == (341, 512)                       # rint(512 * 4 / 6)
    assert f.interpolate_2d(s2, "A", x_pixels=60).shape == (40, 60)
    assert f.interpolate_2d(s2, "A", y_pixels=30).shape == (30, 45)
    assert f.interpolate_2d(s2, "A", x_pixels=11, y_pixels=7).shape == (7, 11)
    assert f.interpolate_2d_line(s2, "A").shape == (512,)
    assert f.interpolate_2d_line(s2, "A", pixels=17).shape == (17,)
    assert f.interpolate_3d_proj(s3, "A", x_pixels=60).shape == (40, 60)
    assert f.interpolate_3d_cross(s3, "A", y_pixels=30).shape == (30, 45)
    assert f.interpolate_3d_grid(s3, "A", x_pixels=30).shape == (40, 20, 30)    # nz = rint(30 * 8 / 6)
    assert f.interpolate_3d_line(s3, "A", pixels=9).shape == (9,)
    vx, vy = f.interpolate_3d_vec(s3, "A", "B", "C", x_pixels=24)
    assert vx.shape == vy.shape == (16, 24)


def test_irregular_bounds(env):
    """test_interpolate.py:911-1007: non-square pixels."""
    f, w = env.f, env.w
    s2 = _frame(env, dict(x=[0], y=[0], A=[4], B=[7], h=[0.9], rho=[0.4], m=[0.03]))
    img = f.interpolate_2d(s2, "A", x_pixels=50, y_pixels=25, xlim=(-1, 1), ylim=(-0.5, 0.5), normalize=False,
                           hmin=False)
    xs = -1 + (np.arange(50) + 0.5) * (2 / 50)
    ys = -0.5 + (np.arange(25) + 0.5) * (1 / 25)
    r = np.sqrt(xs[None, :] ** 2 + ys[:, None] ** 2)
    exp = 0.03 / (0.4 * 0.9 ** 2) * 4 * np.vectorize(lambda q: w("cubic", q, 2))(r / 0.9)
    assert_allclose(img, exp, rtol=1e-6, atol=1e-12)
    s3 = _frame(env, dict(x=[0], y=[0], z=[0], A=[4], h=[0.9], rho=[0.4], m=[0.03]))
    g = f.interpolate_3d_grid(s3, "A", x_pixels=20, y_pixels=10, z_pixels=15, xlim=(-1, 1), ylim=(-0.5, 0.5),
                              zlim=(-0.75, 0.75), normalize=False, hmin=False)
    xs = -1 + (np.arange(20) + 0.5) * (2 / 20)
    ys = -0.5 + (np.arange(10) + 0.5) * (1 / 10)
    zs = -0.75 + (np.arange(15) + 0.5) * (1.5 / 15)
    r = np.sqrt(xs[None, None, :] ** 2 + ys[None, :, None] ** 2 + zs[:, None, None] ** 2)
    exp = 0.03 / (0.4 * 0.9 ** 3) * 4 * np.vectorize(lambda q: w("cubic", q, 3))(r / 0.9)
    assert_allclose(g, exp, rtol=1e-6, atol=1e-12)


def test_oob_particles(env):
    """test_interpolate.py:1010-1114: a particle outside the view still contributes."""
    f, w = env.f, env.w
    s2 = _frame(env, dict(x=[0], y=[0], A=[4], B=[3], h=[1.9], rho=[0.4], m=[0.03]))
    s3 = _frame(env, dict(x=[0], y=[0], z=[0.5], A=[4], B=[3], C=[2], h=[1.9], rho=[0.4], m=[0.03]))
    real = 1 + (np.arange(25) + 0.5) / 25
    r = np.sqrt(real[None, :] ** 2 + real[:, None] ** 2)
    wt = 0.03 / (0.4 * 1.9 ** 2)
    kw = dict(x_pixels=25, y_pixels=25, xlim=(1, 2), ylim=(1, 2), normalize=False, hmin=False)
    img = f.interpolate_2d(s2, "A", **kw)
    assert_allclose(img, wt * 4 * np.vectorize(lambda q: w("cubic", q, 2))(r / 1.9), rtol=1e-6, atol=1e-12)
    line = f.interpolate_2d_line(s2, "A", pixels=25, xlim=(1, 2), ylim=(1, 2), normalize=False, hmin=False)
    assert_allclose(line, wt * 4 * np.vectorize(lambda q: w("cubic", q, 2))(np.sqrt(2) * real / 1.9), rtol=1e-6,
                    atol=1e-12)
    img = f.interpolate_3d_proj(s3, "A", dens_weight=False, **kw)
    assert_allclose(img, wt * 4 * np.vectorize(lambda q: env.col(q, 3))(r / 1.9), rtol=1e-6, atol=1e-12)
    img = f.interpolate_3d_cross(s3, "A", z_slice=0.5, **kw)
    assert_allclose(img, wt / 1.9 * 4 * np.vectorize(lambda q: w("cubic", q, 3))(r / 1.9), rtol=1e-6, atol=1e-12)


def test_exact_interpolation_and_culling(env):
    """test_interpolate.py:1220-1254, :1395-1417: the exact paths integrate to w h^2 / (4 bound^2) over the whole
    support, and do not over-cull."""
    f = env.f
    s2 = _frame(env, dict(x=[0], y=[0], A=[2], h=[1.1], rho=[0.55], m=[0.04]))
    s3 = _frame(env, dict(x=[0], y=[0], z=[1], A=[2], h=[1.1], rho=[0.55], m=[0.04]))
    wt = 0.04 * 2 / (0.55 * 1.1 ** 2)
    bound = 2 * 1.1
    kw = dict(x_pixels=1, xlim=(-bound, bound), ylim=(-bound, bound), exact=True, normalize=False, hmin=False)
    assert f.interpolate_2d(s2, "A", **kw).sum() == pytest.approx(wt * 1.1 ** 2 / (4 * bound ** 2), rel=1e-6)
    assert f.interpolate_3d_proj(s3, "A", dens_weight=False, **kw).sum() == pytest.approx(
        wt * 1.1 ** 2 / (4 * bound ** 2), rel=1e-6)
    t2 = _frame(env, dict(x=[0], y=[0], A=[2], h=[0.4], rho=[0.1], m=[1]))
    t3 = _frame(env, dict(x=[0], y=[0], z=[0], A=[2], h=[0.4], rho=[0.1], m=[1]))
    i2 = f.interpolate_2d(t2, "A", x_pixels=5, xlim=(-1, 1), ylim=(-1, 1), exact=True, normalize=False)
    i3 = f.interpolate_3d_proj(t3, "A", x_pixels=5, xlim=(-1, 1), ylim=(-1, 1), exact=True)
    assert i2[2, 4] != 0 and i3[2, 4] != 0


def test_density_weighted_and_normalize(env):
    """test_interpolate.py:1257-1392: dens_weight and normalize on a single-pixel image."""
    f, w = env.f, env.w
    s2 = _frame(env, dict(x=[0], y=[0], A=[2], B=[3], h=[0.5], rho=[0.25], m=[0.75]))
    kw = dict(x_pixels=1, y_pixels=1, xlim=(-1, 1), ylim=(-1, 1), hmin=False)
    k0 = w("cubic", 0, 2)
    assert f.interpolate_2d(s2, "A", dens_weight=False, normalize=False, **kw)[0, 0] == pytest.approx(
        0.75 * 2 / (0.25 * 0.5 ** 2) * k0, rel=1e-12)
    assert f.interpolate_2d(s2, "A", dens_weight=True, normalize=False, **kw)[0, 0] == pytest.approx(
        0.75 * 2 / 0.5 ** 2 * k0, rel=1e-12)
    assert f.interpolate_2d(s2, "A", dens_weight=False, normalize=True, **kw)[0, 0] == pytest.approx(2.0, rel=1e-12)
    vx, vy = f.interpolate_2d_vec(s2, "A", "B", dens_weight=True, normalize=True, **kw)
    assert vx[0, 0] == pytest.approx(2.0, rel=1e-12) and vy[0, 0] == pytest.approx(3.0, rel=1e-12)
    s3 = _frame(env, dict(x=[0], y=[0], z=[0], A=[2], h=[0.5], rho=[0.25], m=[0.75]))
    # projection: dens_weight defaults to True for any target but 'rho'
    assert f.interpolate_3d_proj(s3, "A", normalize=False, **kw)[0, 0] == pytest.approx(
        0.75 * 2 / 0.5 ** 2 * env.col(0, 3), rel=1e-10)
    assert f.interpolate_3d_proj(s3, "rho", normalize=False, **kw)[0, 0] == pytest.approx(
        0.75 * 0.25 / (0.25 * 0.5 ** 2) * env.col(0, 3), rel=1e-10)
    # an empty pixel normalises to 0, not NaN (np.nan_to_num)
    far = f.interpolate_2d(s2, "A", x_pixels=3, y_pixels=1, xlim=(5, 8), ylim=(-1, 1), normalize=True)
    assert (far == 0).all()


def test_minimum_smoothing_length_3d_and_lines(env):
    """test_interpolate.py:1462-1563: hmin equals clamping h to half a pixel by hand."""
    f, s = env.f, env.s
    rng = np.random.default_rng(9)
    n = 200
    cols = dict(x=rng.uniform(0, 1, n), y=rng.uniform(0, 1, n), z=rng.uniform(0, 1, n), h=rng.uniform(1e-3, 0.04, n),
                m=np.full(n, 1.0 / n), rho=np.ones(n), A=rng.uniform(0, 1, n))
    sdf = s.ParticleFrame(cols, params={})
    clamped = dict(cols)
    clamped["h"] = np.maximum(cols["h"], 0.5 / 16)
    sdc = s.ParticleFrame(clamped, params={})
    kw = dict(x_pixels=16, y_pixels=16, xlim=(0, 1), ylim=(0, 1), normalize=False)
    assert_allclose(f.interpolate_3d_proj(sdf, "A", hmin=True, **kw), f.interpolate_3d_proj(sdc, "A", hmin=False, **kw),
                    rtol=1e-13)
    assert_allclose(f.interpolate_3d_cross(sdf, "A", z_slice=0.5, hmin=True, **kw),
                    f.interpolate_3d_cross(sdc, "A", z_slice=0.5, hmin=False, **kw), rtol=1e-13)
    assert_allclose(f.interpolate_3d_grid(sdf, "A", z_pixels=16, zlim=(0, 1), hmin=True, **kw),
                    f.interpolate_3d_grid(sdc, "A", z_pixels=16, zlim=(0, 1), hmin=False, **kw), rtol=1e-13)
    lc = dict(cols)
    lc["h"] = np.maximum(cols["h"], 0.5 * np.sqrt(3) / 20)
    assert_allclose(f.interpolate_3d_line(sdf, "A", pixels=20, xlim=(0, 1), ylim=(0, 1), zlim=(0, 1), hmin=True,
                                          normalize=False),
                    f.interpolate_3d_line(s.ParticleFrame(lc, params={}), "A", pixels=20, xlim=(0, 1), ylim=(0, 1),
                                          zlim=(0, 1), hmin=False, normalize=False), rtol=1e-13)


def test_right_angle_rotations_are_axis_swaps(env):
    """test_rotation.py (axes_rotation_separation / 64 right-angle rotations): rotating the view by 90 degrees about
    z equals relabelling the axes."""
    f, s = env.f, env.s
    rng = np.random.default_rng(10)
    n = 50
    cols = dict(x=rng.uniform(-1, 1, n), y=rng.uniform(-1, 1, n), z=rng.uniform(-1, 1, n), h=rng.uniform(0.2, 0.5, n),
                m=np.full(n, 0.02), rho=np.ones(n), A=rng.uniform(0, 1, n))
    sdf = s.ParticleFrame(cols, params={})
    kw = dict(x_pixels=18, y_pixels=18, xlim=(-1.5, 1.5), ylim=(-1.5, 1.5), normalize=False)
    rot = f.interpolate_3d_proj(sdf, "A", rotation=[90, 0, 0], rot_origin=[0, 0, 0], **kw)
    # (x, y) -> (-y, x): the rotated image at (X, Y) samples the original at x = Y, y = -X
    swapped = dict(cols)
    swapped["x"], swapped["y"] = -cols["y"], cols["x"]
    ref = f.interpolate_3d_proj(s.ParticleFrame(swapped, params={}), "A", **kw)
    assert_allclose(rot, ref, rtol=1e-9, atol=1e-13)


def test_rotation_forms_agree(env):
    """test_rotation.py:143-260: Euler zyx list, scipy Rotation and quaternion describe the same view; the default
    origin is (0, 0, 0); 'com' and 'midpoint' move it."""
    from scipy.spatial.transform import Rotation
    f, s = env.f, env.s
    rng = np.random.default_rng(11)
    n = 60
    cols = dict(x=rng.uniform(0, 2, n), y=rng.uniform(0, 2, n), z=rng.uniform(0, 2, n), h=rng.uniform(0.2, 0.5, n),
                m=rng.uniform(0.01, 0.03, n), rho=np.ones(n), A=rng.uniform(0, 1, n), vx=rng.normal(size=n),
                vy=rng.normal(size=n), vz=rng.normal(size=n))
    sdf = s.ParticleFrame(cols, params={})
    kw = dict(x_pixels=16, y_pixels=16, xlim=(-1, 3), ylim=(-1, 3), normalize=False)
    a = f.interpolate_3d_proj(sdf, "A", rotation=[35, 65, -20], **kw)
    r = Rotation.from_euler("zyx", [35, 65, -20], degrees=True)
    assert_allclose(f.interpolate_3d_proj(sdf, "A", rotation=r, **kw), a, rtol=1e-11, atol=1e-14)
    assert_allclose(f.interpolate_3d_proj(sdf, "A", rotation=Rotation.from_quat(r.as_quat()), **kw), a, rtol=1e-11,
                    atol=1e-14)
    assert_allclose(f.interpolate_3d_proj(sdf, "A", rotation=[35, 65, -20], rot_origin=[0, 0, 0], **kw), a, rtol=1e-12)
    com = f.interpolate_3d_proj(sdf, "A", rotation=[35, 65, -20], rot_origin="com", **kw)
    assert_allclose(com, f.interpolate_3d_proj(sdf, "A", rotation=[35, 65, -20], rot_origin=sdf.centre_of_mass(), **kw),
                    rtol=1e-11, atol=1e-14)
    mid = [(cols[c].min() + cols[c].max()) / 2 for c in "xyz"]
    assert_allclose(f.interpolate_3d_proj(sdf, "A", rotation=[35, 65, -20], rot_origin="midpoint", **kw),
                    f.interpolate_3d_proj(sdf, "A", rotation=[35, 65, -20], rot_origin=mid, **kw), rtol=1e-11,
                    atol=1e-14)
    with pytest.raises(ValueError):
        f.interpolate_3d_proj(sdf, "A", rotation=[1, 2, 3], rot_origin="centre", **kw)
    # manual rotation of positions and vectors (test_rotation.py:54-141)
    pos = r.apply(np.stack([cols["x"], cols["y"], cols["z"]], 1))
    vel = r.apply(np.stack([cols["vx"], cols["vy"], cols["vz"]], 1))
    man = dict(cols)
    man["x"], man["y"], man["z"] = pos.T
    man["vx"], man["vy"], man["vz"] = vel.T
    gx, gy = f.interpolate_3d_vec(sdf, "vx", "vy", "vz", rotation=[35, 65, -20], **kw)
    mx, my = f.interpolate_3d_vec(s.ParticleFrame(man, params={}), "vx", "vy", "vz", **kw)
    assert_allclose(gx, mx, rtol=1e-9, atol=1e-12)
    assert_allclose(gy, my, rtol=1e-9, atol=1e-12)
