"""The device-side front-end (`sarracen_b200.interpolate`) against the reference's own
front-end: replays the calls recorded by tests/golden/make_golden_frontend.py
(`sarracen.interpolate.interpolate_*`, backend='cpu') and compares at 1e-10."""
import json

import numpy as np
import pytest
from conftest import assert_parity, load_golden

pytestmark = pytest.mark.gpu


def _frames(g):
    import sarracen_b200 as s
    f2 = {k[3:]: g[k] for k in g.files if k.startswith("f2_")}
    f3 = {k[3:]: g[k] for k in g.files if k.startswith("f3_")}
    strip = lambda d: {k: v for k, v in d.items() if k not in ("m", "rho")}
    return {"f2": s.ParticleFrame(f2, params={"hfact": 1.2}),
            "f3": s.ParticleFrame(f3, params={"hfact": 1.2}),
            "f2nomrho": s.ParticleFrame(strip(f2), params={"hfact": 1.2, "mass": 1.0 / len(f2["x"])}),
            "f3nomrho": s.ParticleFrame(strip(f3), params={"hfact": 1.2, "mass": 1.0 / len(f3["x"])})}


def test_frontend_matches_reference_frontend():
    import sarracen_b200 as s
    from sarracen_b200 import interpolate as front
    g = load_golden("frontend")
    calls = json.loads(str(g["calls"]))
    frames = _frames(g)
    kernels = {"cubic": s.CubicSplineKernel, "quartic": s.QuarticSplineKernel, "quintic": s.QuinticSplineKernel}
    for i, (fr, fn, args, kw) in enumerate(calls):
        kwargs = dict(kw)
        if "kernel" in kwargs:
            kwargs["kernel"] = kernels[kwargs["kernel"]]()
        for lim in ("xlim", "ylim", "zlim"):
            if lim in kwargs:
                kwargs[lim] = tuple(kwargs[lim])
        res = getattr(front, fn)(frames[fr], *args, backend="gpu", **kwargs)
        res = res if isinstance(res, tuple) else (res,)
        for j, r in enumerate(res):
            ref = g[f"res_{i}_{j}"]
            assert r.dtype == np.float64
            assert_parity(r, ref, 1e-10, f"call {i} {fn} {kw} [{j}]")


def test_hmin_equals_clamping_by_hand():
    """After sarracen/tests/interpolate/test_interpolate.py:1420-1563 (array-exact)."""
    import sarracen_b200 as s
    from sarracen_b200 import interpolate as front
    rng = np.random.default_rng(5)
    n = 400
    cols = dict(x=rng.uniform(0, 1, n), y=rng.uniform(0, 1, n), h=rng.uniform(0.001, 0.05, n), m=np.full(n, 1.0 / n),
                rho=np.ones(n), A=rng.uniform(0, 1, n))
    sdf = s.ParticleFrame(cols, params={"hfact": 1.2})
    a = front.interpolate_2d(sdf, "A", x_pixels=20, y_pixels=20, xlim=(0, 1), ylim=(0, 1), hmin=True)
    cols2 = dict(cols)
    cols2["h"] = np.maximum(cols["h"], 0.5 * (1.0 / 20))
    b = front.interpolate_2d(s.ParticleFrame(cols2, params={"hfact": 1.2}), "A", x_pixels=20, y_pixels=20,
                             xlim=(0, 1), ylim=(0, 1), hmin=False)
    np.testing.assert_allclose(a, b, rtol=1e-13, atol=0)


def test_normalized_constant_field_is_constant():
    """A constant target normalises to that constant wherever particles contribute, 0 elsewhere (nan_to_num)."""
    import sarracen_b200 as s
    from sarracen_b200 import interpolate as front
    rng = np.random.default_rng(6)
    n = 300
    sdf = s.ParticleFrame(dict(x=rng.uniform(0.3, 0.7, n), y=rng.uniform(0.3, 0.7, n), z=rng.uniform(0.3, 0.7, n),
                               h=np.full(n, 0.05), m=np.full(n, 1.0 / n), rho=np.ones(n), A=np.full(n, 2.5)),
                          params={"hfact": 1.2})
    img = front.interpolate_3d_proj(sdf, "A", x_pixels=30, y_pixels=30, xlim=(0, 1), ylim=(0, 1))
    assert np.isfinite(img).all()
    inside = img != 0
    assert inside.any() and not inside.all()
    np.testing.assert_allclose(img[inside], 2.5, rtol=1e-12)


def test_device_resident_frame_gives_the_same_results():
    """ParticleFrame.to_device(): columns stay on the GPU between calls; same result (to the rounding of the
    order in which red.global.add.f64 lands, which differs from run to run)."""
    import sarracen_b200 as s
    from sarracen_b200 import interpolate as front
    rng = np.random.default_rng(8)
    n = 5000
    cols = dict(x=rng.uniform(0, 1, n), y=rng.uniform(0, 1, n), z=rng.uniform(0, 1, n), h=rng.uniform(0.01, 0.08, n),
                m=np.full(n, 1.0 / n), rho=rng.uniform(0.5, 2, n), T=rng.uniform(0, 1, n))
    sdf = s.ParticleFrame(cols, params={"hfact": 1.2})
    kw = dict(x_pixels=64, y_pixels=48, xlim=(0, 1), ylim=(0, 1), rotation=[10, 20, 30], rot_origin="com")
    a = front.interpolate_3d_proj(sdf, "T", **kw)
    sdf.to_device()
    assert sdf._device_cache["x"].is_cuda
    b = front.interpolate_3d_proj(sdf, "T", **kw)
    np.testing.assert_allclose(a, b, rtol=1e-13, atol=0)
    sdf.release_device()
    assert sdf._device_cache is None


def test_several_targets_share_one_walk():
    """Extension: a list of targets is rendered in one bin / sort / walk and equals the
    single-target calls."""
    import sarracen_b200 as s
    from sarracen_b200 import interpolate as front, ops
    rng = np.random.default_rng(12)
    n = 4000
    cols = dict(x=rng.uniform(0, 1, n), y=rng.uniform(0, 1, n), z=rng.uniform(0, 1, n), h=rng.uniform(0.01, 0.08, n),
                m=np.full(n, 1.0 / n), rho=rng.uniform(0.5, 2, n), T=rng.uniform(0, 1, n), P=rng.uniform(1, 2, n),
                B=rng.normal(size=n))
    sdf = s.ParticleFrame(cols, params={"hfact": 1.2})
    kw = dict(x_pixels=48, y_pixels=40, xlim=(0, 1), ylim=(0, 1))
    for fn, extra in ((front.interpolate_3d_proj, {}), (front.interpolate_3d_cross, dict(z_slice=0.4)),
                      (front.interpolate_3d_grid, dict(z_pixels=8, zlim=(0, 1)))):
        imgs = fn(sdf, ["T", "P", "B"], **kw, **extra)
        assert isinstance(imgs, list) and len(imgs) == 3
        for name, img in zip(("T", "P", "B"), imgs):
            np.testing.assert_allclose(img, fn(sdf, name, **kw, **extra), rtol=1e-12, atol=1e-15)
    with pytest.raises(ValueError):
        front.interpolate_3d_proj(sdf, ["T", "P", "B", "rho"], dens_weight=True, **kw)   # 4 targets + norm > 4 weights
    assert len(front.interpolate_3d_proj(sdf, ["T", "P", "B", "rho"], dens_weight=True, normalize=False, **kw)) == 4
