"""World-size-2 gloo test of the particle sharding and combine logic (CPU, no GPU).

The per-rank partial rasters come from the CPU oracle here; on the GPU box the same
`sarracen_b200.parallel` functions run on CUDA tensors over NCCL (bench.py --gpus N).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import OracleBackend, OracleKernel, set_num_threads
    from sarracen_b200.parallel import combine_grid_along_z, combine_image, shard_range, slab_range
    set_num_threads(2)
    rng = np.random.default_rng(5)
    n = 4001
    x, y, z = rng.uniform(0, 1, (3, n))
    h = rng.uniform(0.02, 0.1, n)
    w = rng.uniform(0.1, 1, n)
    lo, hi = shard_range(n, rank, world)
    k = OracleKernel("cubic")
    part = OracleBackend.interpolate_2d_render(x[lo:hi], y[lo:hi], w[lo:hi], h[lo:hi], k, 2, 24, 18, 0, 1, 0, 1, False)
    img = combine_image(torch.from_numpy(part.copy()), dst=0)
    gpart = OracleBackend.interpolate_3d_grid(x[lo:hi], y[lo:hi], z[lo:hi], w[lo:hi], h[lo:hi], k, 2, 10, 9, 8,
                                              0, 1, 0, 1, 0, 1)
    zl, zh = slab_range(8, rank, world)
    slab = torch.empty((zh - zl, 9, 10), dtype=torch.float64)
    combine_grid_along_z(torch.from_numpy(gpart.copy()), slab)
    if rank == 0:
        full = OracleBackend.interpolate_2d_render(x, y, w, h, k, 2, 24, 18, 0, 1, 0, 1, False)
        q.put(("img", float(np.abs(img.numpy() - full).max() / np.abs(full).max())))
    gfull = OracleBackend.interpolate_3d_grid(x, y, z, w, h, k, 2, 10, 9, 8, 0, 1, 0, 1, 0, 1)
    q.put((f"slab{rank}", float(np.abs(slab.numpy() - gfull[zl:zh]).max() / np.abs(gfull).max())))
    # the pieces of the backend's multi-GPU mode that do not need a device: in-place reduction along z with
    # UNEQUAL slabs (nz = 9 over 2 ranks), gathering the slabs, packed images on every rank
    from sarracen_b200 import parallel
    g9 = OracleBackend.interpolate_3d_grid(x[lo:hi], y[lo:hi], z[lo:hi], w[lo:hi], h[lo:hi], k, 2, 6, 5, 9,
                                           0, 1, 0, 1, 0, 1)
    f9 = OracleBackend.interpolate_3d_grid(x, y, z, w, h, k, 2, 6, 5, 9, 0, 1, 0, 1, 0, 1)
    parallel.enable(images="all", grids="all")
    mode = parallel.active()
    mine = parallel.reduce_along_z(torch.from_numpy(g9.copy()))
    l9, h9 = slab_range(9, rank, world)
    whole = parallel.gather_slabs(mine.clone(), 9, mode)
    imgs = parallel.finish_images([torch.from_numpy(part.copy()), torch.from_numpy(2.0 * part)], mode)
    full = OracleBackend.interpolate_2d_render(x, y, w, h, k, 2, 24, 18, 0, 1, 0, 1, False)
    q.put((f"mode{rank}", max(float(np.abs(mine.numpy() - f9[l9:h9]).max() / np.abs(f9).max()),
                              float(np.abs(whole.numpy() - f9).max() / np.abs(f9).max()),
                              float(np.abs(imgs[0].numpy() - full).max() / np.abs(full).max()),
                              float(np.abs(imgs[1].numpy() - 2 * full).max() / np.abs(full).max()))))
    assert [len(c) for c in parallel.shard_columns([x, y], mode)] == [hi - lo, hi - lo]
    parallel.disable()
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    from sarracen_b200.parallel import shard_range, slab_range
    for n in (0, 1, 7, 1000, (1 << 27) + 3):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
    assert slab_range(512, 3, 8) == (192, 256)
    for nz in (1, 9, 512, 515):
        for world in (1, 2, 8):
            edges = [slab_range(nz, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == nz
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


def test_block_cyclic_plane_ownership(monkeypatch):
    """owned_planes / z_chunks: every plane has exactly one owner for any chunking that divides evenly,
    the default is DEFAULT_CHUNKS reduced until it does, SPHB_Z_CHUNKS overrides it."""
    from sarracen_b200 import parallel
    from sarracen_b200.parallel import owned_planes, slab_range, z_chunks
    monkeypatch.delenv("SPHB_Z_CHUNKS", raising=False)
    assert parallel.DEFAULT_CHUNKS == 8 and parallel.TRANSPORTS == ("peer", "nccl")
    assert z_chunks(512, 8) == 8 and z_chunks(512, 8, 2) == 2 and z_chunks(64, 8) == 8 and z_chunks(48, 8) == 6
    assert z_chunks(24, 8) == 3 and z_chunks(8, 8) == 1 and z_chunks(35, 8) == 0 and z_chunks(7, 2) == 0
    for nz, world, chunks in ((512, 8, None), (512, 2, 4), (64, 4, 8), (48, 8, None), (35, 8, None), (7, 2, 3), (6, 2, 1)):
        planes = [owned_planes(nz, r, world, chunks) for r in range(world)]
        assert sorted(p for pl in planes for p in pl) == list(range(nz))
        c = z_chunks(nz, world, chunks)
        if c == 0:
            assert planes == [list(range(*slab_range(nz, r, world))) for r in range(world)]
        else:
            b = nz // (c * world)
            assert all(pl[:b] == list(range(r * b, (r + 1) * b)) for r, pl in enumerate(planes))
    monkeypatch.setenv("SPHB_Z_CHUNKS", "2")
    assert z_chunks(512, 8) == 2 and z_chunks(512, 8, 4) == 4


def test_two_rank_combine_matches_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(5))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert set(results) == {"img", "slab0", "slab1", "mode0", "mode1"}
    for name, err in results.items():
        assert err < 1e-13, (name, err)
