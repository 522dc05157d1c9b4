#!/usr/bin/env python
"""Time the BASELINE.json configs that are not bench.py's headline lines (configs 0, 2, 4 of the
list: 2-D render, cross-section + vector render, exact paths) on one GPU, device-resident, CUDA
events, 2 warm-ups + 3 timed runs; writes gpurun_out/other_configs.json (copied to profiles/ by hand).

    python profiles/measure_configs.py [--quick]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sarracen_b200 as s  # noqa: E402
from sarracen_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
quick = "--quick" in sys.argv


def timed(fn, warm=2, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def dv(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a.astype(dtype))).to(dev)


results = []


def record(name, n, ms, **extra):
    row = dict(config=name, particles=n, ms=ms, particles_per_s=n / (ms * 1e-3), **extra)
    results.append(row)
    print(json.dumps(row), flush=True)


# ---- config 1 (BASELINE configs[0]): interpolate_2d, 2^20 uniform 2-D particles, cubic, 512^2, f64
n = 1 << 20
rng = np.random.default_rng(1001)
x, y = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
A = rng.uniform(0, 1, n)
h = np.full(n, 1.2 * n ** -0.5)
w = A * (1.0 / n)
r = ops.Raster(512, 512, 0, 1, 0, 1)
k = s.CubicSplineKernel()
for dt in (np.float64, np.float32):
    xd, yd, hd, wd = (dv(a, dt) for a in (x, y, h, w))
    ms = timed(lambda: ops.scatter2d(xd, yd, hd, [wd], k.w, 2, 2, r))
    c = ops.count_contributions(xd, yd, hd, 2.0, r)
    record(f"cfg1 interpolate_2d 512^2 cubic {np.dtype(dt).name}", n, ms, contributions=c,
           stats=dict(ops.last_stats))

# ---- config 3 (BASELINE configs[2]): cross-section + vector render, 2^25 uniform 3-D, 4096^2
n = 1 << (22 if quick else 25)
rng = np.random.default_rng(1003)
x, y, z = rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n)
A = rng.uniform(0, 1, n)
vx, vy = rng.normal(size=n), rng.normal(size=n)
h = np.full(n, 1.2 * n ** (-1 / 3))
m = 1.0 / n
r = ops.Raster(4096, 4096, 0, 1, 0, 1)
for dt in (np.float32, np.float64):
    xd, yd, zd, hd, wd = (dv(a, dt) for a in (x, y, z, h, A * m))
    ms = timed(lambda: ops.scatter2d(xd, yd, hd, [wd], k.w, 2, 3, r, z=zd, z_slice=0.5))
    c = ops.count_contributions(xd, yd, hd, 2.0, r, z=zd, z_slice=0.5)
    record(f"cfg3a interpolate_3d_cross 4096^2 cubic {np.dtype(dt).name}", n, ms, contributions=c,
           stats=dict(ops.last_stats))
    wx, wy = dv(vx * m, dt), dv(vy * m, dt)
    col = k.get_column_kernel_func(1000)
    ms = timed(lambda: ops.scatter2d(xd, yd, hd, [wx, wy], col, 2, 2, r), warm=1, reps=2)
    c = ops.count_contributions(xd, yd, hd, 2.0, r)
    record(f"cfg3b interpolate_3d_vec (2 weights, one walk) 4096^2 cubic column table {np.dtype(dt).name}", n, ms,
           contributions=c, stats=dict(ops.last_stats))
    ms = timed(lambda: ops.scatter2d(xd, yd, hd, [wx, wy, wd], col, 2, 2, r), warm=1, reps=2)
    record(f"cfg3c vector render + normalisation weight (3 weights, one walk) 4096^2 {np.dtype(dt).name}", n, ms,
           contributions=c, stats=dict(ops.last_stats))
    del xd, yd, zd, hd, wd, wx, wy
    torch.cuda.empty_cache()

# ---- huge footprints on the sampled path (missing #7): h >> tile.  2^14 particles whose support
# spans 200 .. 2000 pixels of a 2048^2 image (up to the whole image: ~1000 tile pairs per particle)
n = 1 << 14
rng = np.random.default_rng(1007)
x, y = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
h = np.exp(rng.uniform(np.log(50 / 2048), np.log(500 / 2048), n))
w = rng.uniform(0, 1, n) / n
r = ops.Raster(2048, 2048, 0, 1, 0, 1)
for dt in (np.float64, np.float32):
    xd, yd, hd, wd = (dv(a, dt) for a in (x, y, h, w))
    for kern, wf, nd in (("cubic", k.w, 2), ("cubic column table", k.get_column_kernel_func(1000), 2)):
        ms = timed(lambda: ops.scatter2d(xd, yd, hd, [wd], wf, 2, nd, r))
        c = ops.count_contributions(xd, yd, hd, 2.0, r)
        record(f"huge footprints (box 200-2000 px) 2048^2 {kern} {np.dtype(dt).name}", n, ms, contributions=c,
               contributions_per_s=c / (ms * 1e-3), stats=dict(ops.last_stats))

# ---- config 5 (BASELINE configs[4]): exact paths, 2^23 particles, h log-uniform [pw/20, 20 pw], 1024^2
pw = 1.0 / 1024
for name, n_full, fn in (("exact2d", 1 << 23, ops.exact2d), ("exact3d_proj", 1 << 23, ops.exact3d_proj)):
    n = n_full >> (6 if quick else (0 if name == "exact2d" else 3))
    rng = np.random.default_rng(1005)
    x, y = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    A = rng.uniform(0, 1, n)
    h = np.exp(rng.uniform(np.log(pw / 20), np.log(20 * pw), n))
    w = A * (1.0 / n)
    r = ops.Raster(1024, 1024, 0, 1, 0, 1)
    xd, yd, hd, wd = (dv(a, np.float64) for a in (x, y, h, w))
    t0 = time.perf_counter()
    ms = timed(lambda: fn(xd, yd, hd, [wd], r), warm=1, reps=1)
    bbox = float(np.mean((np.ceil(4 * h / pw) + 1) ** 2))
    record(f"cfg5 {name} 1024^2 cubic f64", n, ms, mean_bbox_pixels=bbox, wall_s=time.perf_counter() - t0)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(results, open(os.path.join(ROOT, "gpurun_out", "other_configs.json"), "w"), indent=1)
