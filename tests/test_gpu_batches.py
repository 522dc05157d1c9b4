"""several products of one geometry in one call -- K cross-sections (one pass when the depths are
evenly spaced), K rotated views sharing one upload, several target columns sharing one walk -- each checked
against the CPU oracle (1e-10), not against the path itself."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from conftest import assert_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import sarracen_b200 as s
    from oracle import OracleBackend, OracleKernel
    from sarracen_b200 import interpolate as f

    class Env:
        pass
    e = Env()
    e.s, e.f, e.ob, e.ok = s, f, OracleBackend, OracleKernel
    rng = np.random.default_rng(77)
    n = 30000
    e.cols = dict(x=rng.random(n), y=rng.random(n), z=rng.random(n), A=rng.random(n), B=rng.normal(size=n),
                  h=np.exp(rng.uniform(np.log(0.008), np.log(0.07), n)), m=np.full(n, 1.0 / n),
                  rho=rng.uniform(0.5, 2.0, n))
    e.sdf = s.ParticleFrame(e.cols, params={"hfact": 1.2})
    e.kw = dict(x_pixels=72, y_pixels=60, xlim=(0.05, 0.95), ylim=(0.1, 0.9))
    return e


def _oracle_cross(e, z_val, target="A", normalize=False):
    c = e.cols
    w = c[target] * c["m"] / c["rho"]
    args = (e.ok("cubic"), 2.0, 72, 60, 0.05, 0.95, 0.1, 0.9)
    img = e.ob.interpolate_3d_cross(c["x"], c["y"], c["z"], z_val, w, c["h"], *args)
    if normalize:
        norm = e.ob.interpolate_3d_cross(c["x"], c["y"], c["z"], z_val, c["m"] / c["rho"], c["h"], *args)
        img = np.nan_to_num(img / norm)
    return img


@pytest.mark.parametrize("normalize", [False, True])
def test_evenly_spaced_depths_in_one_pass(env, normalize):
    from sarracen_b200 import ops
    depths = np.linspace(0.2, 0.8, 7)
    stack = env.f.interpolate_3d_cross(env.sdf, "A", z_slice=depths, normalize=normalize, **env.kw)
    assert stack.shape == (7, 60, 72)
    assert ops.last_stats["tile"][2] > 1          # rendered by the voxel-grid pipeline: one pass for all depths
    for i, d in enumerate(depths):
        assert_parity(stack[i], _oracle_cross(env, float(d), normalize=normalize), 1e-10, f"depth {d:.2f}")


def test_irregular_depths_and_several_targets(env):
    depths = [0.13, 0.5, 0.55, 0.91]
    a, b = env.f.interpolate_3d_cross(env.sdf, ["A", "B"], z_slice=depths, normalize=False, **env.kw)
    assert a.shape == b.shape == (4, 60, 72)
    for i, d in enumerate(depths):
        assert_parity(a[i], _oracle_cross(env, d, "A"), 1e-10, f"A at {d}")
        assert_parity(b[i], _oracle_cross(env, d, "B"), 1e-10, f"B at {d}")
    single = env.f.interpolate_3d_cross(env.sdf, "A", z_slice=0.5, normalize=False, **env.kw)
    assert_parity(a[1], single, 1e-13, "stack entry vs single call")


def test_rotated_views_share_one_upload(env):
    rots = [[0, 0, 0], [30, 20, 10], [90, 0, 45]]
    kw = dict(x_pixels=64, y_pixels=64, xlim=(-0.3, 1.3), ylim=(-0.3, 1.3), normalize=False, dens_weight=False)
    views = env.f.interpolate_3d_proj_views(env.sdf, "A", rots, rot_origin=[0.5, 0.5, 0.5], **kw)
    assert views.shape == (3, 64, 64)
    c = env.cols
    w = c["A"] * c["m"] / c["rho"]
    col = env.ok.column("cubic", 1000)
    for i, r in enumerate(rots):
        pos = np.stack([c["x"], c["y"], c["z"]], axis=1) - 0.5
        pos = Rotation.from_euler("zyx", r, degrees=True).apply(pos) + 0.5
        ref = env.ob.interpolate_3d_projection(pos[:, 0].copy(), pos[:, 1].copy(), w, c["h"], col, 2.0, 64, 64,
                                               -0.3, 1.3, -0.3, 1.3, False)
        assert_parity(views[i], ref, 1e-10, f"view {r}")
