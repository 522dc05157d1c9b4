#!/usr/bin/env python
"""BASELINE config 5 at 2^17 particles (1024^2, h log-uniform in [pw/20, 20 pw]): two calls of each exact
path -- the command the exact-kernel ncu captures of round 2 profile (profiles/run_r2_ncu_b.sh)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sarracen_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
n, pw = 1 << 17, 1.0 / 1024
rng = np.random.default_rng(1005)
cols = [rng.uniform(0, 1, n), rng.uniform(0, 1, n),
        np.exp(rng.uniform(np.log(pw / 20), np.log(20 * pw), n)), rng.uniform(0, 1, n) / n]
x, y, h, w = (torch.from_numpy(c).to(dev) for c in cols)
r = ops.Raster(1024, 1024, 0, 1, 0, 1)
for fn in (ops.exact2d, ops.exact2d, ops.exact3d_proj, ops.exact3d_proj):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn(x, y, h, [w], r)[0]
    e1.record()
    torch.cuda.synchronize()
    print(fn.__name__, e0.elapsed_time(e1), "ms", float(out.sum()))
