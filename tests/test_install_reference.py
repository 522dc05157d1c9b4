"""Drop-in boundary against the REAL reference objects (authoring container only:
skipped when /root/reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference not present")


def test_install_routes_backend_gpu():
    sar = ref_shim.load_reference()
    import sarracen_b200
    from sarracen.interpolate import interpolate as front
    sarracen_b200.install()
    assert front.get_backend("gpu") is sarracen_b200.B200Backend
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
