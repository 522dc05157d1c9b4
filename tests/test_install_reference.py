"""Drop-in boundary against the REAL reference objects (the unmodified Sarracen from baseline/_ref
or /root/reference, see oracle/ref_shim.py; skipped when neither is present).  The `gpu` tests
run Sarracen's own front-end with backend='gpu' on the B200 and compare with backend='cpu'."""
import numpy as np
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference not present")


def test_install_routes_backend_gpu():
    sar = ref_shim.load_reference()
    import sarracen_b200
    from sarracen.interpolate import interpolate as front
    installed = sarracen_b200.install()
    from sarracen.interpolate.base_backend import BaseBackend
    assert front.get_backend("gpu") is installed
    assert issubclass(installed, sarracen_b200.B200Backend) and issubclass(installed, BaseBackend)
    for name in (n for n in vars(BaseBackend) if n.startswith("interpolate_")):     # our method, not the stub
        assert getattr(installed, name) is getattr(sarracen_b200.B200Backend, name)
    assert front.get_backend("cpu") is sar.interpolate.CPUBackend
    with pytest.raises(ValueError):
        front.get_backend("tpu")


def test_reference_weight_functions_resolve():
    ref_shim.load_reference()
    from sarracen.kernels import CubicSplineKernel, QuarticSplineKernel, QuinticSplineKernel
    from sarracen_b200 import resolve_weight_function
    for K, kind in ((CubicSplineKernel, "cubic"), (QuarticSplineKernel, "quartic"), (QuinticSplineKernel, "quintic")):
        k = K()
        wf = resolve_weight_function(k.w, k.get_radius())          # class-level numba dispatcher
        assert wf.kind == kind and wf.radius == k.get_radius()
        col = resolve_weight_function(k.get_column_kernel_func(64), k.get_radius())   # closure over the table
        assert col.kind == "lut" and col.radius == k.get_radius()
        np.testing.assert_array_equal(col.lut, k.get_column_kernel(64))


def test_frontend_reaches_the_backend_with_reference_signatures(monkeypatch):
    """sdf.sph_interpolate / interpolate_3d_proj with backend='gpu' call B200Backend with the
    reference's positional arguments (checked with a recording subclass; no GPU needed)."""
    sar = ref_shim.load_reference()
    import sarracen_b200
    from sarracen.interpolate import interpolate as front
    calls = []

    class Spy(sarracen_b200.B200Backend):
        @staticmethod
        def interpolate_3d_projection(*args):
            calls.append(("proj", args))
            return np.ones((args[7], args[6]))

        @staticmethod
        def interpolate_3d_grid(*args):
            calls.append(("grid", args))
            return np.ones((args[9], args[8], args[7]))

    monkeypatch.setattr(front, "GPUBackend", Spy)
    rng = np.random.default_rng(0)
    sdf = sar.SarracenDataFrame({"x": rng.random(50), "y": rng.random(50), "z": rng.random(50),
                                 "h": np.full(50, 0.1), "m": np.full(50, 0.02), "rho": np.ones(50)},
                                params={"hfact": 1.2})
    img = front.interpolate_3d_proj(sdf, "rho", x_pixels=8, y_pixels=6, backend="gpu", normalize=False)
    assert img.shape == (6, 8) and calls[0][0] == "proj"
    wf = sarracen_b200.resolve_weight_function(calls[0][1][4], calls[0][1][5])
    assert wf.kind == "lut" and calls[0][1][12] is False
    cube = sdf.sph_interpolate("rho", x_pixels=4, y_pixels=5, z_pixels=3, backend="gpu")
    assert cube.shape == (3, 5, 4) and calls[-1][0] == "grid"
    assert sarracen_b200.resolve_weight_function(calls[-1][1][5], calls[-1][1][6]).kind == "cubic"


# ---- the real front-end on the real GPU (E1) --------------------------------------------------
def _reference_frame(sar, n=20000, seed=5):
    rng = np.random.default_rng(seed)
    h = np.exp(rng.uniform(np.log(0.01), np.log(0.08), n))
    return sar.SarracenDataFrame({"x": rng.random(n), "y": rng.random(n), "z": rng.random(n),
                                  "vx": rng.normal(size=n), "vy": rng.normal(size=n), "vz": rng.normal(size=n),
                                  "A": rng.random(n), "h": h, "m": np.full(n, 1.0 / n),
                                  "rho": rng.uniform(0.5, 2.0, n)}, params={"hfact": 1.2})


@pytest.mark.gpu
def test_real_frontend_on_gpu_matches_reference_cpu_backend():
    """Sarracen's own `interpolate_*` / `sph_interpolate` with backend='gpu' (-> B200Backend on the
    B200) against the same call with backend='cpu' (Sarracen's numba CPUBackend): 1e-10."""
    from conftest import assert_parity
    sar = ref_shim.load_reference()
    import sarracen_b200
    from sarracen.interpolate import interpolate as front
    sarracen_b200.install()
    sdf = _reference_frame(sar)
    kw = dict(x_pixels=96, y_pixels=80, xlim=(0.1, 0.9), ylim=(0.05, 0.95))
    for name, call in {
        "interpolate_3d_proj": lambda b: front.interpolate_3d_proj(sdf, "A", backend=b, **kw),
        "interpolate_3d_proj rotated": lambda b: front.interpolate_3d_proj(
            sdf, "A", rotation=[30, 20, 10], backend=b, normalize=False, **kw),
        "interpolate_3d_cross": lambda b: front.interpolate_3d_cross(sdf, "A", z_slice=0.4, backend=b, **kw),
        "sph_interpolate (3-D grid)": lambda b: sdf.sph_interpolate(
            "A", x_pixels=40, y_pixels=36, z_pixels=32, backend=b),
        "interpolate_3d_line": lambda b: front.interpolate_3d_line(sdf, "A", pixels=200, backend=b),
    }.items():
        assert_parity(call("gpu"), call("cpu"), 1e-10, name)
    for name, call in {
        "interpolate_3d_vec": lambda b: front.interpolate_3d_vec(sdf, "vx", "vy", "vz", backend=b, **kw),
        "interpolate_3d_cross_vec": lambda b: front.interpolate_3d_cross_vec(
            sdf, "vx", "vy", "vz", z_slice=0.6, backend=b, **kw),
    }.items():
        for g, c in zip(call("gpu"), call("cpu")):
            assert_parity(g, c, 1e-10, name)


@pytest.mark.gpu
def test_real_frontend_2d_and_exact_on_gpu():
    from conftest import assert_parity
    sar = ref_shim.load_reference()
    import sarracen_b200
    from sarracen.interpolate import interpolate as front
    sarracen_b200.install()
    rng = np.random.default_rng(9)
    n = 3000
    sdf = sar.SarracenDataFrame({"x": rng.uniform(0.2, 0.8, n), "y": rng.uniform(0.2, 0.8, n), "A": rng.random(n),
                                 "h": np.exp(rng.uniform(np.log(0.004), np.log(0.06), n)),
                                 "m": np.full(n, 1.0 / n), "rho": rng.uniform(0.5, 2.0, n)}, params={"hfact": 1.2})
    kw = dict(x_pixels=64, y_pixels=48, xlim=(0, 1), ylim=(0, 1))
    assert_parity(front.interpolate_2d(sdf, "A", backend="gpu", **kw),
                  front.interpolate_2d(sdf, "A", backend="cpu", **kw), 1e-10, "interpolate_2d")
    # exact: without the normalisation pass (outside the particles' supports both backends hold only
    # the rounding residue of cancelling edge integrals, and the ratio of two residues is noise)
    assert_parity(front.interpolate_2d(sdf, "A", exact=True, normalize=False, backend="gpu", **kw),
                  front.interpolate_2d(sdf, "A", exact=True, normalize=False, backend="cpu", **kw), 1e-10,
                  "interpolate_2d exact")
    assert_parity(front.interpolate_2d_line(sdf, "A", pixels=128, backend="gpu"),
                  front.interpolate_2d_line(sdf, "A", pixels=128, backend="cpu"), 1e-10, "interpolate_2d_line")
    # our own front-end fed with the reference's data frame and kernel objects (numba dispatchers)
    from sarracen_b200 import interpolate as ours
    assert_parity(ours.interpolate_2d(sdf, "A", kernel=sar.kernels.QuinticSplineKernel(), backend="gpu", **kw),
                  front.interpolate_2d(sdf, "A", kernel=sar.kernels.QuinticSplineKernel(), backend="cpu", **kw),
                  1e-10, "sarracen_b200.interpolate_2d on a SarracenDataFrame")
