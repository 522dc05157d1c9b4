"""Host-side semantics of the front-end that need no GPU: validation order and messages
(sarracen/interpolate/interpolate.py:164-255; the reference's tests test_dimension_check,
test_invalid_region, test_required_columns, test_interpolate.py:273-306, :1117-1217)."""
import numpy as np
import pytest

import sarracen_b200 as s
from sarracen_b200 import interpolate as front


def _f2(n=5, **extra):
    cols = dict(x=np.linspace(0, 1, n), y=np.linspace(0, 1, n), h=np.full(n, 0.1), m=np.full(n, 0.1),
                rho=np.ones(n), A=np.ones(n))
    cols.update(extra)
    return s.ParticleFrame(cols, params={"hfact": 1.2})


def _f3(n=5):
    return _f2(n, z=np.linspace(0, 1, n), vx=np.ones(n), vy=np.ones(n), vz=np.ones(n))


def test_special_columns_and_dim():
    f = s.ParticleFrame({"rx": [0.0], "ry": [1.0], "rz": [2.0], "h": [1.0], "mass": [1.0], "density": [1.0]})
    assert (f.xcol, f.ycol, f.zcol, f.hcol, f.mcol, f.rhocol) == ("rx", "ry", "rz", "h", "mass", "density")
    assert f.get_dim() == 3 and _f2().get_dim() == 2
    g = s.ParticleFrame({"a": [0.0], "b": [1.0]})
    assert (g.xcol, g.ycol, g.zcol, g.hcol) == ("a", "b", None, None)
    assert isinstance(f.kernel, s.CubicSplineKernel)
    with pytest.raises(TypeError):
        f.kernel = "cubic"


def test_dimension_check():
    for fn, args in ((front.interpolate_2d, ("A",)), (front.interpolate_2d_vec, ("A", "A")),
                     (front.interpolate_2d_line, ("A",))):
        with pytest.raises(ValueError, match="not 2-dimensional"):
            fn(_f3(), *args)
    for fn, args in ((front.interpolate_3d_proj, ("A",)), (front.interpolate_3d_cross, ("A",)),
                     (front.interpolate_3d_grid, ("A",)), (front.interpolate_3d_line, ("A",)),
                     (front.interpolate_3d_vec, ("vx", "vy", "vz")), (front.interpolate_3d_cross_vec, ("vx", "vy", "vz"))):
        with pytest.raises(ValueError, match="not 3-dimensional"):
            fn(_f2(), *args)


def test_required_columns():
    f = _f2()
    with pytest.raises(KeyError, match="x-directional"):
        front.interpolate_2d(f, "A", x="nope")
    no_h = s.ParticleFrame({"x": [0.0], "y": [0.0], "m": [1.0], "rho": [1.0], "A": [1.0]})
    with pytest.raises(KeyError, match="Smoothing length"):
        front.interpolate_2d(no_h, "A")
    no_m = s.ParticleFrame({"x": [0.0], "y": [0.0], "h": [1.0], "rho": [1.0], "A": [1.0]})
    with pytest.raises(KeyError, match="mass"):
        front.interpolate_2d(no_m, "A")
    no_rho = s.ParticleFrame({"x": [0.0], "y": [0.0], "h": [1.0], "m": [1.0], "A": [1.0]})
    with pytest.raises(KeyError, match="Density"):
        front.interpolate_2d(no_rho, "A")


def test_backend_codes():
    with pytest.raises(ValueError, match="Invalid backend"):
        front.interpolate_2d(_f2(), "A", backend="tpu")
    with pytest.raises(ValueError, match="no CPU path"):
        front.interpolate_2d(_f2(), "A", backend="cpu")


def test_default_axes_and_pixels():
    f = _f3()
    assert front._default_xyz(f, None, None, None) == ("x", "y", "z")
    assert front._default_xyz(f, "y", None, None) == ("y", "x", "z")
    assert front._default_xyz(f, "z", "x", None) == ("z", "x", "y")
    assert front._default_xy(_f2(), None, "x") == ("y", "x")
    assert front._set_pixels(None, None, (0, 2), (0, 1)) == (512, 256)
    assert front._set_pixels(None, 100, (0, 2), (0, 1)) == (200, 100)
    assert front._set_pixels(30, None, (0, 1), (0, 3)) == (30, 90)


def test_invalid_region_messages():
    with pytest.raises(ValueError, match="xlim"):
        front._check_boundaries(5, 5, (1, 1), (0, 1))
    with pytest.raises(ValueError, match="ylim"):
        front._check_boundaries(5, 5, (0, 1), (2, 1))
    with pytest.raises(ValueError, match="x_pixels"):
        front._check_boundaries(0, 5, (0, 1), (0, 1))
    with pytest.raises(ValueError, match="y_pixels"):
        front._check_boundaries(5, -1, (0, 1), (0, 1))


def test_corotation_matches_reference_formula():
    rot, origin = front._corotate([[0.1, 0.2, 0.0], [0.6, -0.3, 0.0]], None)
    assert origin == [0.1, 0.2, 0.0]
    np.testing.assert_allclose(rot, [-np.degrees(np.arctan2(-0.5, 0.5)), 0, 0])


def test_upload_chunk_schedule_and_host_column_views():
    """Host logic of backend.streamed_scatter: the chunk schedule covers the particle range exactly
    once, grows geometrically and respects its caps; host columns are wrapped without copies
    (read-only and pandas columns included) and anything non-numeric falls back to the plain path."""
    import warnings

    import pandas as pd
    import torch

    from sarracen_b200 import backend

    for n in (1 << 22, (1 << 24) + 12345, 1 << 27, 5_000_001):
        b = backend._chunk_bounds(n)
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
        sizes = [hi - lo for lo, hi in b]
        assert backend._CHUNK_FIRST_MIN <= sizes[0] <= backend._CHUNK_FIRST_MAX
        assert max(sizes) <= backend._CHUNK_MAX
        assert all(sizes[i + 1] <= backend._CHUNK_GROWTH * sizes[i] for i in range(len(sizes) - 1))
    assert len(backend._chunk_bounds(1 << 24)) == 3 and len(backend._chunk_bounds(1 << 27)) == 9

    a = np.arange(6.0)
    ro = np.arange(6.0)
    ro.flags.writeable = False
    with warnings.catch_warnings():
        warnings.simplefilter("error")   # no "non-writable array" warning, no copy
        cols = backend._host_columns([a, ro, pd.Series(a), torch.arange(6.0), a[::-1].astype(">f4")])
    assert cols is not None and all(c.shape == (6,) for c in cols)
    assert cols[0].data_ptr() == a.ctypes.data and cols[1].data_ptr() == ro.ctypes.data
    assert cols[4].dtype == torch.float32 and cols[4][0].item() == 5.0
    assert backend._common_dtype(cols) == torch.float64
    assert backend._common_dtype([cols[4], cols[4]]) == torch.float32
    assert backend._host_columns([a, np.array(["x"] * 6)]) is None     # not numeric
    assert backend._host_columns([a, np.arange(5.0)]) is None          # ragged


def test_bench_workloads_are_rank_count_independent():
    """bench.py's fixed particle sets: any rank count generates the same set (independently seeded blocks),
    and the exact workloads keep BASELINE config 5's shape (h log-uniform in [pw/20, 20 pw], centres inside)."""
    import bench
    for name, n in (("proj", 1 << 16), ("grid", 1 << 16), ("exact2d", 1 << 15), ("exact3d", 1 << 15)):
        wl = bench.describe(name, n, "f64")
        whole = bench.generate(wl, 0, 1, "f64")
        parts = [bench.generate(wl, r, 4, "f64") for r in range(4)]
        for key, col in whole.items():
            np.testing.assert_array_equal(np.concatenate([p[key] for p in parts]), col)
        assert whole["x"].shape[0] == n
    wl = bench.describe("exact2d", 1 << 15, "f64")
    c = bench.generate(wl, 0, 1, "f64")
    pw = 1.0 / 1024
    assert c["h"].min() >= pw / 20 and c["h"].max() <= 20 * pw
    assert c["x"].min() >= 0.0 and c["x"].max() <= 1.0 and set(c) == {"x", "y", "h", "w"}
    assert bench.default_particles("exact3d") == 1 << 23 and bench.default_particles("grid") == 1 << 27
