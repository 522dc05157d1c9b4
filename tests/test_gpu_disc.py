"""radial-binning disc profiles on the GPU against the reference's pandas implementation
(`sarracen.disc`, when the reference is installed) and against an independent NumPy restatement."""
import numpy as np
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.gpu


def _disc(n=200_000, seed=21):
    rng = np.random.default_rng(seed)
    r = rng.uniform(0.5, 3.0, n) ** 1.3
    phi = rng.uniform(0, 2 * np.pi, n)
    z = rng.normal(0, 0.05 * r)
    vk = r ** -0.5
    cols = dict(x=r * np.cos(phi), y=r * np.sin(phi), z=z, vx=-vk * np.sin(phi), vy=vk * np.cos(phi),
                vz=rng.normal(0, 0.01, n), h=rng.uniform(0.01, 0.05, n).astype(np.float32),
                rho=rng.uniform(0.5, 2, n), m=rng.uniform(0.5, 1.5, n) / n, T=rng.random(n))
    return cols


def _numpy_profiles(c, bins, log, origin, geometry):
    x, y, z = c["x"] - origin[0], c["y"] - origin[1], c["z"] - origin[2]
    r = np.sqrt(x * x + y * y + (z * z if geometry == "spherical" else 0.0))
    eps = np.finfo(float).eps
    lo, hi = r.min() - eps, r.max() + eps
    edges = np.logspace(np.log10(lo), np.log10(hi), bins + 1) if log else np.linspace(lo, hi, bins + 1)
    idx = np.searchsorted(edges, r, "left") - 1
    ok = (idx >= 0) & (idx < bins)
    def tot(v):
        return np.bincount(idx[ok], weights=v[ok], minlength=bins)
    cnt = np.bincount(idx[ok], minlength=bins)
    sigma = tot(c["m"]) / (np.pi * (edges[1:] ** 2 - edges[:-1] ** 2))
    with np.errstate(invalid="ignore", divide="ignore"):
        mean_t = tot(c["T"]) / cnt
    L = [tot(c["m"] * (y * c["vz"] - z * c["vy"])), tot(c["m"] * (z * c["vx"] - x * c["vz"])),
         tot(c["m"] * (x * c["vy"] - y * c["vx"]))]
    return edges, cnt, sigma, mean_t, L


@pytest.mark.parametrize("log,geometry", [(False, "cylindrical"), (True, "cylindrical"), (False, "spherical")])
def test_profiles_against_numpy(log, geometry):
    import sarracen_b200 as s
    from sarracen_b200 import disc
    c = _disc()
    origin = [0.01, -0.02, 0.005]
    sdf = s.ParticleFrame(c, params={"hfact": 1.2}).to_device()
    edges, cnt, sigma, mean_t, L = _numpy_profiles(c, 120, log, origin, geometry)
    out, mid = disc.surface_density(sdf, bins=120, log=log, geometry=geometry, origin=origin, retbins=True)
    np.testing.assert_allclose(out, sigma, rtol=1e-12)
    np.testing.assert_allclose(mid, np.sqrt(edges[:-1] * edges[1:]) if log else 0.5 * (edges[1:] + edges[:-1]),
                               rtol=1e-14)
    np.testing.assert_allclose(disc.azimuthal_average(sdf, "T", bins=120, log=log, geometry=geometry, origin=origin),
                               mean_t, rtol=1e-12)
    lx, ly, lz = disc.angular_momentum(sdf, bins=120, log=log, geometry=geometry, origin=origin, unit_vector=False)
    for got, ref in zip((lx, ly, lz), L):
        np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-12 * np.abs(L[2]).max())


@pytest.mark.skipif(not ref_shim.available(), reason="reference not installed")
def test_profiles_against_the_reference():
    """Every function of sarracen.disc against the GPU version, same arguments."""
    sar = ref_shim.load_reference()
    import sarracen_b200 as s
    from sarracen_b200 import disc
    c = _disc(60_000)
    ref_sdf = sar.SarracenDataFrame(c, params={"hfact": 1.2})
    sdf = s.ParticleFrame(c, params={"hfact": 1.2})
    kw = dict(bins=50, r_in=0.6, r_out=3.5)
    np.testing.assert_allclose(disc.surface_density(sdf, **kw), sar.disc.surface_density(ref_sdf, **kw), rtol=1e-12)
    np.testing.assert_allclose(disc.surface_density(sdf, log=True, **kw),
                               sar.disc.surface_density(ref_sdf, log=True, **kw), rtol=1e-12)
    np.testing.assert_allclose(disc.azimuthal_average(sdf, "T", **kw), sar.disc.azimuthal_average(ref_sdf, "T", **kw),
                               rtol=1e-12)
    for a, b in zip(disc.angular_momentum(sdf, **kw), sar.disc.angular_momentum(ref_sdf, **kw)):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)
    h, mid = disc.scale_height(sdf, retbins=True, **kw)
    h_ref, mid_ref = sar.disc.scale_height(ref_sdf, retbins=True, **kw)
    np.testing.assert_allclose(h, h_ref, rtol=1e-9)
    np.testing.assert_allclose(mid, mid_ref, rtol=1e-14)
    # the h column is float32: pandas averages it to a float32 mean, here the mean is taken in double
    np.testing.assert_allclose(disc.honH(sdf, **kw), sar.disc.honH(ref_sdf, **kw), rtol=3e-7)
    c64 = dict(c, h=c["h"].astype(np.float64))
    np.testing.assert_allclose(disc.honH(s.ParticleFrame(c64, params={}), **kw),
                               sar.disc.honH(sar.SarracenDataFrame(c64, params={}), **kw), rtol=1e-9)
    # a global particle mass instead of a mass column; default (data) range
    c2 = {k: v for k, v in c.items() if k != "m"}
    np.testing.assert_allclose(disc.surface_density(s.ParticleFrame(c2, params={"mass": 1e-5}), bins=40),
                               sar.disc.surface_density(sar.SarracenDataFrame(c2, params={"mass": 1e-5}), bins=40),
                               rtol=1e-12)
    with pytest.raises(ValueError):
        disc.surface_density(sdf, geometry="cartesian")
