#!/usr/bin/env python
"""Where the end-to-end time of a host-array call goes (one GPU): raw pinned H2D / D2H bandwidth,
device pipeline per chunk size, and the streamed backend call.  Development aid for
backend._streamed; prints JSON lines."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sarracen_b200 as s  # noqa: E402
from sarracen_b200 import backend, ops  # noqa: E402

dev = torch.device("cuda", 0)


def wall(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


n = 1 << 27
h_pin = torch.empty(n, dtype=torch.float64).pin_memory()
h_pin.fill_(1.0)
d = torch.empty(n, dtype=torch.float64, device=dev)
ms = wall(lambda: d.copy_(h_pin, non_blocking=True))
print(json.dumps({"what": "H2D pinned 1 GiB", "ms": ms, "GB/s": n * 8 / ms / 1e6}), flush=True)
ms = wall(lambda: h_pin.copy_(d, non_blocking=True))
print(json.dumps({"what": "D2H pinned 1 GiB", "ms": ms, "GB/s": n * 8 / ms / 1e6}), flush=True)
cs = torch.cuda.Stream(dev)


def both():
    with torch.cuda.stream(cs):
        d.copy_(h_pin, non_blocking=True)
    d2 = torch.empty_like(d)
    d2.fill_(2.0)


ms = wall(both)
print(json.dumps({"what": "H2D 1 GiB on a copy stream + fill kernel on main", "ms": ms}), flush=True)
t0 = time.perf_counter()
p2 = torch.empty(n, dtype=torch.float64, pin_memory=True)
print(json.dumps({"what": "pinned alloc 1 GiB (first)", "ms": (time.perf_counter() - t0) * 1e3}), flush=True)
del p2
t0 = time.perf_counter()
p2 = torch.empty(n, dtype=torch.float64, pin_memory=True)
print(json.dumps({"what": "pinned alloc 1 GiB (cached)", "ms": (time.perf_counter() - t0) * 1e3}), flush=True)
del p2, d, h_pin

# grid workload (config 4)
which = sys.argv[1] if len(sys.argv) > 1 else "grid"
rng = np.random.default_rng(5)
if which == "grid":
    n = 1 << 27
    cols = {k: rng.uniform(0, 1, n) for k in "xyz"}
    cols["h"] = np.full(n, 1.2 * n ** (-1 / 3))
    cols["w"] = rng.uniform(0, 1, n) / n
    k = s.CubicSplineKernel()
    call = lambda hc: s.B200Backend.interpolate_3d_grid(hc["x"], hc["y"], hc["z"], hc["w"], hc["h"], k.w, 2.0, 512, 512,
                                                        512, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
elif which in ("proj", "proj32"):
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.make_workload("proj", None, 0, "f32" if which == "proj32" else "f64")
    cols, n = wl["cols"], wl["n"]
    col = s.QuinticSplineKernel().get_column_kernel_func(1000)
    call = lambda hc: s.B200Backend.interpolate_3d_projection(hc["x"], hc["y"], hc["w"], hc["h"], col, wl["radius"],
                                                              wl["nx"], wl["ny"], *wl["lim"], False)
if "--pageable" in sys.argv:
    ms = wall(lambda: call(cols), reps=2)
    print(json.dumps({"what": f"{which} e2e from pageable NumPy arrays, default schedule", "ms": ms}), flush=True)
host = {kk: torch.from_numpy(v).pin_memory() for kk, v in cols.items()}
for first, growth, cap in ((1 << 23, 3, 1 << 24), (1 << 23, 16, 1 << 24), (1 << 23, 4, 1 << 25), (1 << 23, 1, 1 << 23), (1 << 22, 2, 1 << 25), (0, 0, 0)):
    if first == 0:
        backend._STREAM_MIN = 1 << 40   # single upload
    else:
        backend._CHUNK_FIRST_MAX, backend._CHUNK_GROWTH, backend._CHUNK_MAX = first, growth, cap
    ms = wall(lambda: call(host), reps=2)
    print(json.dumps({"what": f"{which} e2e, first<= {first}, growth {growth}, cap {cap}",
                      "chunks": len(backend._chunk_bounds(n)) if first else 1, "ms": ms}), flush=True)
