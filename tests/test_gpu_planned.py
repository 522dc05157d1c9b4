"""the plan / execute form of the scatter -- no hidden synchronisation.  `sphb_scatter_planned`
must be capturable into a CUDA graph (a capture fails on any synchronising call), reproduce `sphb_scatter2d / 3d`
with the plan of the same inputs, and survive a stale plan (excess dropped, flag raised)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    import sarracen_b200 as s
    from sarracen_b200 import ops

    class Env:
        pass
    e = Env()
    e.torch, e.s, e.ops = torch, s, ops
    rng = np.random.default_rng(5)
    n = 300_000
    dev = torch.device("cuda", 0)
    e.cols = {k: torch.from_numpy(rng.random(n)).to(dev) for k in ("x", "y", "z")}
    e.cols["h"] = torch.from_numpy(np.exp(rng.uniform(np.log(0.002), np.log(0.06), n))).to(dev)   # cell + tile path
    e.cols["w"] = torch.from_numpy(rng.random(n)).to(dev)
    e.dev = dev
    return e


def test_planned_image_matches_and_is_graph_capturable(env):
    torch, ops, c = env.torch, env.ops, env.cols
    wf = env.s.QuinticSplineKernel().get_column_kernel_func(1000)
    r = ops.Raster(512, 384, 0.0, 1.0, 0.0, 1.0)
    args = (c["x"], c["y"], c["h"], [c["w"]], wf, 3.0, 2, r)
    ref = ops.scatter2d(*args)[0].clone()
    stats = dict(ops.last_stats)
    plan = ops.scatter_plan(*args)
    torch.cuda.synchronize()
    assert plan.c.n_pairs == stats["n_pairs"] and plan.c.n_small == stats["n_direct"] and plan.c.n_large > 0
    ops.scatter_planned_workspace(plan, *args)
    out = [torch.empty((384, 512), dtype=torch.float64, device=env.dev)]
    ops.scatter_planned(plan, *args, out=out)                       # warm-up outside the capture
    assert float((out[0] - ref).abs().max() / ref.abs().max()) < 1e-13
    assert not ops.scatter_overflowed(env.dev)
    # a capture raises on any synchronising (or otherwise illegal) call inside the library
    graph = torch.cuda.CUDAGraph()
    out[0].zero_()
    with torch.cuda.graph(graph):
        ops.scatter_planned(plan, *args, out=out)
    graph.replay()
    torch.cuda.synchronize()
    assert float((out[0] - ref).abs().max() / ref.abs().max()) < 1e-13
    c["w"].mul_(3.0)                                               # new data in the same buffers, same graph
    graph.replay()
    torch.cuda.synchronize()
    assert float((out[0] - 3.0 * ref).abs().max() / ref.abs().max()) < 1e-12
    c["w"].div_(3.0)


def test_planned_grid_and_cross_section(env):
    torch, ops, c = env.torch, env.ops, env.cols
    wf = env.s.CubicSplineKernel().w
    r3 = ops.Raster(96, 80, 0.0, 1.0, 0.0, 1.0, 64, 0.0, 1.0)
    args = (c["x"], c["y"], c["h"], [c["w"], c["h"]], wf, 2.0, 3, r3)
    ref = ops.scatter3d(c["x"], c["y"], c["z"], c["h"], [c["w"], c["h"]], wf, 2.0, r3)
    plan = ops.scatter_plan(*args, z=c["z"])
    torch.cuda.synchronize()
    ops.scatter_planned_workspace(plan, *args, z=c["z"])
    out = [torch.empty((64, 80, 96), dtype=torch.float64, device=env.dev) for _ in range(2)]
    ops.scatter_planned(plan, *args, out=out, z=c["z"])
    for o, g in zip(out, ref):
        assert float((o - g).abs().max() / g.abs().max()) < 1e-13
    r2 = ops.Raster(200, 160, 0.0, 1.0, 0.0, 1.0)
    a2 = (c["x"], c["y"], c["h"], [c["w"]], wf, 2.0, 3, r2)
    ref2 = ops.scatter2d(*a2, z=c["z"], z_slice=0.37)[0]
    plan2 = ops.scatter_plan(*a2, z=c["z"], z_slice=0.37)
    torch.cuda.synchronize()
    ops.scatter_planned_workspace(plan2, *a2, z=c["z"], z_slice=0.37)
    out2 = [torch.empty((160, 200), dtype=torch.float64, device=env.dev)]
    ops.scatter_planned(plan2, *a2, out=out2, z=c["z"], z_slice=0.37)
    assert float((out2[0] - ref2).abs().max() / ref2.abs().max()) < 1e-13


def test_stale_plan_is_flagged_not_fatal(env):
    torch, ops, c = env.torch, env.ops, env.cols
    wf = env.s.CubicSplineKernel().w
    r = ops.Raster(256, 256, 0.0, 1.0, 0.0, 1.0)
    half = c["x"].numel() // 2
    small = (c["x"][:half], c["y"][:half], c["h"][:half] * 0.5, [c["w"][:half]], wf, 2.0, 2, r)
    full = (c["x"], c["y"], c["h"], [c["w"]], wf, 2.0, 2, r)
    plan = ops.scatter_plan(*small)                    # fewer particles, smaller boxes than the call below
    torch.cuda.synchronize()
    ops.scatter_planned_workspace(plan, *full)
    out = [torch.empty((256, 256), dtype=torch.float64, device=env.dev)]
    ops.scatter_planned(plan, *full, out=out)
    assert ops.scatter_overflowed(env.dev)
    assert bool(torch.isfinite(out[0]).all())
    ref = ops.scatter2d(*full)[0]                      # the library still works afterwards
    assert float(ref.sum()) > 0


def test_sum_slots_receive_side():
    """sphb_sum_slots (receive side of the multi-GPU exchange): dst = own + sum of the slots except `skip`,
    odd and even lengths, against torch."""
    import torch
    from sarracen_b200 import ops
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(5)
    for n in (1, 2, 999, 4098, (1 << 20) + 1):
        for n_slots, skip in ((1, 0), (2, 1), (8, 3)):
            own = torch.rand(n, generator=g, dtype=torch.float64).to(dev)
            slots = torch.rand(n_slots, n, generator=g, dtype=torch.float64).to(dev)
            dst = torch.full((n,), -1.0, dtype=torch.float64, device=dev)
            ops.sum_slots_(dst, own, slots, skip)
            keep = [s for s in range(n_slots) if s != skip]
            ref = own + (slots[keep].sum(dim=0) if keep else 0.0)
            assert torch.allclose(dst, ref, rtol=1e-14, atol=0.0)
    with pytest.raises(ValueError):
        ops.sum_slots_(torch.zeros(4, dtype=torch.float64, device=dev), torch.zeros(4, dtype=torch.float64, device=dev),
                       torch.zeros(2, 5, dtype=torch.float64, device=dev), 0)
