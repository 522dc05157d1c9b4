#!/usr/bin/env python
"""Benchmark of the SPH particle -> grid scatter path (BASELINE.json metric:
particles/sec for interpolate_3d_proj & interpolate_3d_grid at 1/2/4/8 B200, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload proj|grid|grid_hmin]
                    [--sub grid,grid_hmin,exact2d,exact3d|none] [--dtype f64|f32] [--impl reference]

A "step" is one pass of the hot path over one FIXED synthetic particle set:

  proj       BASELINE configs[1]: interpolate_3d_proj column render, 2^24 Plummer-sphere
             particles, quintic kernel + integrated column table (1000 samples), 2048 x 2048.
  grid       BASELINE configs[3]: interpolate_3d_grid, 2^27 uniform particles -> 512^3 voxels,
             cubic kernel, h = 1.2 N^-1/3 (2h / voxel = 2.44).
  grid_hmin  the HBM-bound variant of configs[3]: same particles, h = 0.3 voxel
             with hmin=True, i.e. clamped to 0.5 voxel.
  exact2d, exact3d   BASELINE configs[4] (sub-records only): interpolate_2d / interpolate_3d_proj with
             exact=True, 2^23 particles, h log-uniform in [pw/20, 20 pw], 1024 x 1024.

The JSON line's top level is the headline workload (default `proj`); `workloads` holds one full
sub-record (ms_per_step, value, phases_ms, roofline incl. traffic, e2e, cpu_baseline) for every
other workload of `--sub` (default: grid, grid_hmin, exact2d, exact3d), so the whole metric -- both functions,
and the HBM-bound variant -- is in the one line the driver reads.

`value` times the device pipeline (bin + sort + render [+ collective]) with inputs resident in
HBM; `e2e` times the reference-facing call `B200Backend.interpolate_3d_projection / _grid` with
PAGEABLE host NumPy arrays (what Sarracen's front-end passes), host<->device copies inside the
timed region.  With N GPUs the SAME particle set is sharded by contiguous index range (strong
scaling: total work fixed), every rank renders its shard into a full-size partial raster and the
partial rasters are combined inside the timed region: NCCL reduce to rank 0 (images) or a
reduce-scatter along z, issued per z-chunk over peer memory (copy-engine pushes + sphb_sum_slots; NCCL
reduce_scatter where symmetric memory is unavailable) and overlapped with the render of the following
chunks (voxel grids, sarracen_b200/parallel.py).  `ms_collective` is the exchange timed on its own, `ms_compute` the
step without it; `collective_hidden` = (ms_compute + ms_collective - ms_per_step) / ms_collective.

`--impl reference` times the reference's own CPU implementation of the headline workload on the
host cores: Sarracen's numba `CPUBackend` from baseline/_ref (kind "reference") when it imports,
else the C/OpenMP port in oracle/ (kind "port"); each step renders a bounded sample of the set.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


# ------------------------------------------------------------------ workloads
# Every workload is a FIXED particle set made of independently seeded blocks, so that any rank
# count reproduces the same set: rank r of G generates (and renders) blocks [r B / G, (r+1) B / G).
def _blocks(n_total):
    nb = max(8, n_total >> 20)
    return [(b * n_total // nb, (b + 1) * n_total // nb) for b in range(nb)]


def _plummer_block(n, seed, n_total, a=1.0, rmax=10.0):
    """draw order positions (u, cos theta, phi), then A."""
    rng = np.random.default_rng(seed)
    u = rng.uniform(0.0, (1 + a * a / rmax ** 2) ** -1.5, n)
    ct = rng.uniform(-1.0, 1.0, n)
    ph = rng.uniform(0.0, 2 * np.pi, n)
    A = rng.uniform(0.0, 1.0, n)
    r = a / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    st = np.sqrt(1.0 - ct * ct)
    rho = 3.0 / (4 * np.pi * a ** 3) * (1 + r * r / (a * a)) ** -2.5
    m = 1.0 / n_total
    h = 1.2 * (m / rho) ** (1.0 / 3.0)
    w = A * m / rho  # weight A m / rho  (interpolate.py:468-472)
    return dict(x=r * st * np.cos(ph), y=r * st * np.sin(ph), h=h, w=w)


def _exact_block(n, seed, n_total):
    """centres strictly inside the image, h log-uniform in [pw/20, 20 pw]."""
    rng = np.random.default_rng(seed)
    pw = 1.0 / 1024
    x = rng.uniform(0.0, 1.0, n)
    y = rng.uniform(0.0, 1.0, n)
    A = rng.uniform(0.0, 1.0, n)
    h = np.exp(rng.uniform(np.log(pw / 20), np.log(20 * pw), n))
    return dict(x=x, y=y, h=h, w=A * (1.0 / n_total))


def _cube_block(n, seed, n_total):
    """SURVEY 8(d) config 4 (h is filled in by the workload: constant)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.0, 1.0, n)
    y = rng.uniform(0.0, 1.0, n)
    z = rng.uniform(0.0, 1.0, n)
    A = rng.uniform(0.0, 1.0, n)
    return dict(x=x, y=y, z=z, w=A * (1.0 / n_total) / 1.0)


def grid_side(n_total):
    return 512 if n_total >= (1 << 26) else max(32, int(round((n_total / (1 << 27)) ** (1 / 3) * 512 / 32)) * 32)


def describe(name, n_total, dtype):
    """Static description of a workload (no arrays)."""
    if name in ("exact2d", "exact3d"):
        fn = "interpolate_2d(exact=True)" if name == "exact2d" else "interpolate_3d_proj(exact=True)"
        return dict(key=name, name=fn, kind="cubic", lut=name == "exact3d", radius=2.0, ndim=2, fields=4,
                    nx=1024, ny=1024, nz=1, lim=(0.0, 1.0, 0.0, 1.0), n=n_total, seed=1005,
                    desc=f"{fn}: {n_total} uniform particles, h log-uniform in [pw/20, 20 pw], cubic, 1024x1024, f64")
    if name == "proj":
        return dict(key="proj", name="interpolate_3d_proj", kind="quintic", lut=True, radius=3.0, ndim=2, fields=4,
                    nx=2048, ny=2048, nz=1, lim=(-2.0, 2.0, -2.0, 2.0), n=n_total, seed=1002,
                    desc=f"interpolate_3d_proj: {n_total} Plummer-sphere particles, quintic column table (1000 "
                         f"samples), 2048x2048, {dtype}")
    m = grid_side(n_total)
    if name == "grid":
        hval, tag = 1.2 * n_total ** (-1.0 / 3.0), "h = 1.2 N^-1/3"
    else:
        hval, tag = max(0.3 / m, 0.5 / m), "h = 0.3 voxel clamped by hmin to 0.5 voxel"
    return dict(key=name, name="interpolate_3d_grid", kind="cubic", lut=False, radius=2.0, ndim=3, fields=5,
                nx=m, ny=m, nz=m, lim=(0.0, 1.0, 0.0, 1.0, 0.0, 1.0), n=n_total, seed=1004, hval=hval,
                desc=f"interpolate_3d_grid: {n_total} uniform particles -> {m}^3 voxels, cubic, {tag}, {dtype}")


def generate(wl, rank, world, dtype, share=None):
    """This rank's shard of the workload's particle set as contiguous NumPy columns."""
    np_dtype = np.float64 if dtype == "f64" else np.float32
    blocks = _blocks(wl["n"])
    mine = blocks[len(blocks) * rank // world: len(blocks) * (rank + 1) // world]
    if share is not None:  # grid_hmin reuses the positions of grid
        cols = {k: v for k, v in share.items() if k != "h"}
    else:
        parts = []
        for b, (lo, hi) in enumerate(mine, start=len(blocks) * rank // world):
            gen = {"proj": _plummer_block, "exact2d": _exact_block, "exact3d": _exact_block}.get(wl["key"], _cube_block)
            parts.append(gen(hi - lo, [wl["seed"], b], wl["n"]))
        cols = {k: np.ascontiguousarray(np.concatenate([p[k] for p in parts]).astype(np_dtype)) for k in parts[0]}
    if wl["key"] in ("grid", "grid_hmin"):
        cols["h"] = np.full(cols["x"].shape[0], wl["hval"], dtype=np_dtype)
    return cols


def alg_bytes(wl, itemsize):
    # every particle field read once, every output element written once
    return wl["n"] * wl["fields"] * itemsize + wl["nx"] * wl["ny"] * wl["nz"] * 8


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons during the timed region, sampled through NVML
    (a light in-process query; spawning nvidia-smi every 200 ms perturbs the run)."""

    def __init__(self, gpu_index, period=0.5):
        self.period = period
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() \
                else gpu_index
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.handle = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.reasons |= {k for k, bit in names.items() if mask & bit}
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def start(self):
        if self.handle is not None:
            self.thread.start()
            self.started = True

    def stop(self):
        self.stop_flag.set()
        if getattr(self, "started", False):
            self.thread.join(timeout=3)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ------------------------------------------------------------------ CPU legs
def host_threads():
    """All the host threads this process may use, whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


class CpuReference:
    """The reference's CPU implementation of the path on the host cores.

    kind "reference": Sarracen's own numba `CPUBackend` (sarracen/interpolate/cpu_backend.py),
    imported unmodified from baseline/_ref through oracle/ref_shim.py, called at the backend
    boundary with ONE reused weight function per kernel (excludes the
    re-JIT artefact of base_kernel.py:84-99) after an untimed warm-up call (JIT).
    kind "port": oracle/sph_oracle.c (C + OpenMP restatement), when Sarracen / numba do not import."""

    def __init__(self, prefer_port=False):
        self.threads = host_threads()
        self.kind, self.backend, self.note = "port", None, ""
        self._wf = {}
        if not prefer_port:
            try:
                os.environ["NUMBA_NUM_THREADS"] = str(self.threads)
                os.environ.pop("OMP_NUM_THREADS", None)
                from oracle import ref_shim
                sar = ref_shim.load_reference()
                import numba
                numba.set_num_threads(min(self.threads, numba.config.NUMBA_NUM_THREADS))
                self.threads = numba.get_num_threads()
                self.kernels = sar.kernels
                self.backend = sar.interpolate.CPUBackend
                self.kind = "reference"
                self.note = (f"Sarracen {sar.__version__} CPUBackend, numba {numba.__version__}, "
                             f"{self.threads} threads")
            except Exception as e:  # no reference install, or numba cannot come up
                self.note = f"reference CPUBackend unavailable ({type(e).__name__}: {e}); "
        if self.backend is None:
            import oracle
            oracle.set_num_threads(self.threads)
            self.threads = oracle.num_threads()
            self.oracle = oracle
            self.note += f"oracle/sph_oracle.c with {self.threads} OpenMP threads"

    def weight_function(self, wl):
        key = (wl["kind"], wl["lut"])
        if key not in self._wf:
            if self.kind == "reference":
                k = {"cubic": self.kernels.CubicSplineKernel, "quartic": self.kernels.QuarticSplineKernel,
                     "quintic": self.kernels.QuinticSplineKernel}[wl["kind"]]()
                self._wf[key] = k.get_column_kernel_func(1000) if wl["lut"] else k.w
            else:
                self._wf[key] = (self.oracle.OracleKernel.column(wl["kind"], 1000) if wl["lut"]
                                 else self.oracle.OracleKernel(wl["kind"]))
        return self._wf[key]

    def run(self, wl, cols, ns, m=None, h=None):
        """One call over the first `ns` particles; returns seconds."""
        be = self.backend if self.kind == "reference" else self.oracle.OracleBackend
        wf = self.weight_function(wl)
        sl = {k: v[:ns] for k, v in cols.items()}
        hh = sl["h"] if h is None else h[:ns]
        t0 = time.perf_counter()
        if wl["key"] == "exact2d":
            be.interpolate_2d_render(sl["x"], sl["y"], sl["w"], hh, wf, wl["radius"], wl["nx"], wl["ny"], *wl["lim"], True)
        elif wl["nz"] == 1:
            be.interpolate_3d_projection(sl["x"], sl["y"], sl["w"], hh, wf, wl["radius"], wl["nx"], wl["ny"],
                                         *wl["lim"], wl["key"] == "exact3d")
        else:
            be.interpolate_3d_grid(sl["x"], sl["y"], sl["z"], sl["w"], hh, wf, wl["radius"], m, m, m, *wl["lim"])
        return time.perf_counter() - t0


def cpu_sample(cpu, wl, cols, budget_s=15.0, min_particles=0, max_s=None):
    """Time the CPU implementation on a bounded sample of the workload (same arrays, same raster)."""
    n_have = cols["x"].shape[0]
    if wl["nz"] == 1:
        ns = min(n_have, 1 << 16)
        cpu.run(wl, cols, min(ns, 4096))          # JIT / thread start-up, untimed
        t = cpu.run(wl, cols, ns)
        t = cpu.run(wl, cols, ns)
        # per-call fixed cost (the per-thread images: zero + serial reduce, cpu_backend.py:253,
        # :310-311) is part of the reference; the sample must be large enough that it does not
        # dominate: at least `min_particles`, and up to what the budget allows
        rate = ns / max(t, 1e-3)
        if max_s is not None:   # ... but a slow host must still finish: the floor gives way to the time cap
            min_particles = min(min_particles, int(rate * max_s))
        ns2 = int(min(n_have, max(min_particles, rate * budget_s)))
        t2 = cpu.run(wl, cols, ns2) if ns2 > ns else t
        if ns2 <= ns:
            ns2 = ns
        elif t2 < 0.6 * budget_s and ns2 < n_have:
            ns3 = int(min(n_have, ns2 * budget_s / max(t2, 1e-3)))
            if ns3 > ns2 * 1.2:
                ns2, t2 = ns3, cpu.run(wl, cols, ns3)
        return {"value": ns2 / t2, "unit": "particles/s", "cores": cpu.threads, "kind": cpu.kind, "seconds": t2,
                "sample_particles": ns2, "sample_fraction": ns2 / wl["n"],
                "sample": f"first {ns2} of {wl['n']} particles of the same set, same {wl['nx']}x{wl['ny']} raster, "
                          f"{t2:.2f} s; {cpu.note}"}
    # voxel grid: the reference makes nz full passes over the particles (cpu_backend.py:213-218), far too
    # slow at full size; time a geometrically similar problem (same 2h / voxel), SURVEY 8(d).
    m = 64
    res = None
    while True:
        ratio = (wl["nx"] / m) ** 3
        ns = min(n_have, max(1024, int(wl["n"] / ratio)))
        h = np.full(ns, cols["h"][0] * (wl["nx"] / m), dtype=cols["h"].dtype)
        if res is None:
            cpu.run(wl, cols, min(ns, 2048), m, h)   # JIT, untimed
        t = cpu.run(wl, cols, ns, m, h)
        res = (ns, m, t)
        if t < budget_s / 8 and m < 256 and m * 2 <= wl["nx"]:
            m *= 2
            continue
        break
    ns, m, t = res
    scale = m / wl["nx"]
    return {"value": ns / t * scale, "unit": "particles/s", "cores": cpu.threads, "kind": cpu.kind, "seconds": t,
            "measured_particles_per_s": ns / t, "sample_particles": ns, "sample_fraction": ns / wl["n"],
            "sample": f"{ns} particles -> {m}^3 voxels with the same 2h/voxel: {ns / t:.4g} particles/s in {t:.2f} s; "
                      f"the reference makes nz full passes over the particles (cpu_backend.py:213-218), so "
                      f"particles/s falls as 1/nz: value = measured x {scale:.4g} for {wl['nx']}^3 (extrapolated); "
                      f"{cpu.note}"}


# ------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = describe(args.workload, args.particles or default_particles(args.workload), args.dtype)
    cols = generate(wl, 0, 1, args.dtype)
    cpu = CpuReference(prefer_port=args.cpu_port)
    calls = args.warmup + args.steps
    # every step renders the same bounded sample; sized so that the whole run stays within a few
    # minutes, but never below 2^22 particles for the projection (fixed cost < 5 %)
    budget = max(2.0, 150.0 / calls)
    first = cpu_sample(cpu, wl, cols, budget_s=budget, min_particles=1 << 22, max_s=330.0 / calls)
    per_step, secs = [], []
    for i in range(calls):
        if wl["nz"] == 1:
            t = cpu.run(wl, cols, first["sample_particles"])
            v = first["sample_particles"] / t
        else:
            again = cpu_sample(cpu, wl, cols, budget_s=budget)
            t, v = again["seconds"], again["value"]
        if i >= args.warmup:
            per_step.append(v)
            secs.append(t)
    value = float(np.mean(per_step))
    base = dict(first, value=value, seconds=float(np.mean(secs)))
    line = {"impl": "reference", "metric": "particles/sec", "value": value, "unit": "particles/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(secs)) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic", "config": {"workload": wl["desc"]}, "cpu_baseline": base,
            "e2e": {"value": value, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def default_particles(name):
    return {"proj": 1 << 24, "exact2d": 1 << 23, "exact3d": 1 << 23}.get(name, 1 << 27)


# ------------------------------------------------------------------ our arm
def measure(wl, cols, args, ctx, want_cpu, clock_sampler=None):
    """All figures of one workload on this rank set; returns the record (rank 0) or None."""
    import torch
    import torch.distributed as dist
    import sarracen_b200 as s
    from sarracen_b200 import ops, parallel

    dev, world, rank = ctx["dev"], ctx["world"], ctx["rank"]
    kernel = {"cubic": s.CubicSplineKernel, "quartic": s.QuarticSplineKernel, "quintic": s.QuinticSplineKernel}[
        wl["kind"]]()
    wf = kernel.get_column_kernel_func(1000) if wl["lut"] else kernel.w
    dim3 = wl["nz"] > 1
    raster = ops.Raster(wl["nx"], wl["ny"], *wl["lim"][:4], wl["nz"], *(wl["lim"][4:] if dim3 else (0.0, 1.0)))
    n_local = cols["x"].shape[0]
    itemsize = cols["x"].dtype.itemsize

    devc = {k: torch.from_numpy(v).to(dev) for k, v in cols.items()}
    out_shape = (wl["nz"], wl["ny"], wl["nx"]) if dim3 else (wl["ny"], wl["nx"])
    out = [torch.empty(out_shape, dtype=torch.float64, device=dev)]
    slab_shape = (wl["nz"] // world, wl["ny"], wl["nx"]) if (dim3 and world > 1) else out_shape
    slab = torch.empty(slab_shape, dtype=torch.float64, device=dev) if (dim3 and world > 1) else None
    torch.cuda.synchronize()

    def render_only():
        if dim3:
            ops.scatter3d(devc["x"], devc["y"], devc["z"], devc["h"], [devc["w"]], wf, wl["radius"], raster, out=out)
        else:
            ops.scatter2d(devc["x"], devc["y"], devc["h"], [devc["w"]], wf, wl["radius"], wl["ndim"], raster,
                          out=out)

    def collective_only():
        if dim3:
            parallel.reduce_along_z(out[0])      # the exchange alone, same transport, nothing to hide behind
        else:
            parallel.combine_image(out[0], dst=0)

    def step_device():
        if world == 1:
            render_only()
        elif dim3:
            parallel.scatter3d_sharded(devc["x"], devc["y"], devc["z"], devc["h"], [devc["w"]], wf, wl["radius"],
                                       raster, out=out)
        else:
            render_only()
            collective_only()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, collect=False):
        phases = {"ms_bin": 0.0, "ms_count": 0.0, "ms_sort": 0.0, "ms_render": 0.0, "ms_direct": 0.0}
        launches = 0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
            if collect:
                for k in phases:
                    phases[k] += ops.last_stats.get(k, 0.0)
                launches += ops.last_stats.get("launches", 0)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, {k: v / steps for k, v in phases.items()}, launches

    # ---- device-resident timing -------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    if clock_sampler is not None:
        clock_sampler.start()
    ms_step, _, _ = timed(step_device, args.steps)
    clocks = clock_sampler.stop() if clock_sampler is not None else None
    # the phases (bin / sort / render / direct) are timed in a second pass: bracketing them with events
    # adds a stream synchronisation per call, which the headline number must not carry
    ops.time_phases = True
    ms_phased, phases, launches = timed(render_only, max(1, min(args.steps, 5)), collect=True)
    ops.time_phases = False
    launches = launches // max(1, min(args.steps, 5)) * args.steps
    n_pairs, n_direct = ops.last_stats.get("n_pairs", 0), ops.last_stats.get("n_direct", 0)
    tile = ops.last_stats.get("tile")
    ms_compute = ms_collective = hidden = None
    if world > 1:
        ms_compute, _, _ = timed(render_only, max(1, min(args.steps, 5)))
        ms_collective, _, _ = timed(collective_only, max(1, min(args.steps, 5)))
        hidden = max(0.0, min(1.0, (ms_compute + ms_collective - ms_step) / ms_collective)) if ms_collective else None
    value = wl["n"] / (ms_step * 1e-3)

    # ---- end-to-end timing through the backend API -----------------------------
    e2e = None
    if not args.no_e2e:
        order = "xyzwh" if dim3 else "xywh"
        host_out = torch.empty(slab_shape, dtype=torch.float64).pin_memory() if world > 1 else None

        def step_e2e():
            if world > 1:
                # N ranks: pageable host shard -> device (streamed), partial raster, NCCL combine, result back
                # to the host (rank 0's image, or every rank's z-slab of the grid)
                if dim3:
                    part = s.backend.streamed_scatter(
                        [cols[k] for k in order], dev, lambda c, o, acc: ops.scatter3d(
                            c[0], c[1], c[2], c[4], [c[3]], wf, wl["radius"], raster, out=o, accumulate=acc))[0]
                    host_out.copy_(parallel.reduce_along_z(part))
                else:
                    part = s.backend.streamed_scatter(
                        [cols[k] for k in order], dev, lambda c, o, acc: ops.scatter2d(
                            c[0], c[1], c[3], [c[2]], wf, wl["radius"], wl["ndim"], raster, out=o, accumulate=acc))[0]
                    parallel.combine_image(part, dst=0)
                    if rank == 0:
                        host_out.copy_(part)
                torch.cuda.synchronize()
                return host_out
            # one GPU: the call a Sarracen user reaches through backend='gpu': NumPy arrays in, NumPy array out
            if dim3:
                return s.B200Backend.interpolate_3d_grid(cols["x"], cols["y"], cols["z"], cols["w"], cols["h"], wf,
                                                         wl["radius"], wl["nx"], wl["ny"], wl["nz"], *wl["lim"])
            return s.B200Backend.interpolate_3d_projection(cols["x"], cols["y"], cols["w"], cols["h"], wf,
                                                           wl["radius"], wl["nx"], wl["ny"], *wl["lim"], False)

        e2e_steps = max(10, min(args.steps, 20))
        # two warm-ups that keep the previous result alive, like the timed loop does (the backend returns
        # the raster in page-locked memory from torch's caching host allocator; a first-time 1 GiB
        # page-locked allocation costs ~0.4 s)
        img = step_e2e()
        img = step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            img = step_e2e()  # noqa: F841
        torch.cuda.synchronize()
        t_e2e = (time.perf_counter() - t0) / e2e_steps
        te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        t_e2e = float(te.item())
        h2d = sum(v.nbytes for v in cols.values())
        d2h = int(np.prod(slab_shape)) * 8
        e2e = {"value": wl["n"] / t_e2e, "unit": "particles/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e * 1e3, "steps": e2e_steps,
               "host_memory": "pageable NumPy arrays in; result in a new NumPy array"
                              + (" (per rank: its shard up, its slab / rank 0's image down)" if world > 1 else "")}

    # ---- contributions (outside the timed region) ------------------------------
    contrib = ops.count_contributions(devc["x"], devc["y"], devc["h"], wl["radius"], raster,
                                      z=devc.get("z") if dim3 else None)
    ct = torch.tensor([contrib], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ct)
    contrib = int(ct.item())
    del devc, out, slab
    ops.release_workspaces()
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    # the dominant kernel: the one that took longest -- tile kernel (large footprints), cell kernel (small
    # footprints) or the emit kernel of the binning (count = ms_count, emit = ms_bin - ms_count); its share
    # of the per-rank algorithmic bytes is the whole shard (every record read once, every output element
    # written once)
    phases["ms_emit"] = max(0.0, phases["ms_bin"] - phases["ms_count"])
    dom = max(("ms_render", "ms_direct", "ms_emit"), key=lambda k: phases[k])
    ms_dom = phases[dom]
    dom_kernel = {"ms_render": "render_kernel", "ms_direct": "cell_kernel", "ms_emit": "emit_kernel"}[dom]
    bytes_rank = n_local * wl["fields"] * itemsize + wl["nx"] * wl["ny"] * wl["nz"] * 8
    achieved = bytes_rank / (ms_dom * 1e-3) / 1e9 if ms_dom > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp) and world == 1:
        traffic = (json.load(open(tp)).get(f"{wl['key']}_{args.dtype}") or {}).get(dom_kernel)
    whole = alg_bytes(wl, itemsize) / (ms_step * 1e-3) / 1e9 / peak / world

    sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
    ms_eval = max(phases["ms_render"], phases["ms_direct"])   # the kernel that evaluates the contributions
    rate = contrib / world / (ms_eval * 1e-3) if ms_eval > 0 else 0.0   # per GPU
    if args.dtype == "f64":
        compute = {"bound": "fp64 pipe + issue slots", "contributions_per_s_per_gpu_in_kernel": rate,
                   "peak_fp64_lane_inst_per_s": 148 * 64 * sm_clock * 1e6,
                   "fp64_lane_inst_per_contribution_at_peak": 148 * 64 * sm_clock * 1e6 / rate if rate else None}
    else:
        compute = {"bound": "issue slots", "contributions_per_s_per_gpu_in_kernel": rate,
                   "peak_fp32_lane_inst_per_s": 148 * 128 * sm_clock * 1e6,
                   "lane_inst_per_contribution_at_peak": 148 * 128 * sm_clock * 1e6 / rate if rate else None}

    rec = {
        "metric": "particles/sec", "value": value, "unit": "particles/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": wl["desc"], "particles_total": wl["n"], "particles_per_gpu": n_local,
                   "l2": f"inputs larger than L2 ({n_local * wl['fields'] * itemsize / 1e6:.0f} MB of particle "
                         "columns per GPU per step)", "tile": tile, "pairs": n_pairs, "direct_particles": n_direct,
                   "parallelism": f"particle-sharded x{world} (fixed set, strong scaling)" + (
                       "" if world == 1 else (", NCCL reduce-scatter along z in 2 z-chunks overlapped with the render" if dim3
                                              else ", NCCL reduce to rank 0"))},
        "pixel_contributions": contrib, "contributions_per_s": contrib / (ms_step * 1e-3),
        "phases_ms": phases, "ms_step_with_phase_events": ms_phased,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "kernel": dom_kernel,
                     "alg_bytes": bytes_rank, "peak_source": peak_src, "whole_step_frac": whole},
        "compute_bound": compute,
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    if world > 1:
        rec.update({"ms_compute": ms_compute, "ms_collective": ms_collective, "collective_hidden": hidden,
                    "collective": ("reduce-scatter along z in %d z-chunks overlapped with the staged render; transport %s; "
                                   "ms_collective times the same exchange on its own after a finished render"
                                   % (parallel.z_chunks(wl["nz"], world), {
                                       "peer": "peer memory (copy-engine pushes into the owners' symmetric-memory "
                                               "slots + sphb_sum_slots)",
                                       "nccl": "ncclReduceScatter on NCCL's high-priority stream"}.get(
                                       parallel.last_transport, str(parallel.last_transport)))) if dim3
                                  else "ncclReduce to rank 0"})
    if want_cpu:
        rec["cpu_baseline"] = cpu_sample(ctx["cpu"], wl, cols, budget_s=args.cpu_budget,
                                         min_particles=(1 << 22) if not dim3 else 0)
    return rec


def measure_exact(wl, cols, args, ctx, want_cpu):
    """BASELINE configs[4]: the exact=True paths at full size -- device-resident time, end to end through
    the backend, reference CPU on a sample.  Fewer steps than the scatter workloads (a step is seconds)."""
    import torch
    import torch.distributed as dist
    import sarracen_b200 as s
    from sarracen_b200 import ops, parallel

    dev, world, rank = ctx["dev"], ctx["world"], ctx["rank"]
    fn = ops.exact2d if wl["key"] == "exact2d" else ops.exact3d_proj
    raster = ops.Raster(wl["nx"], wl["ny"], *wl["lim"])
    devc = {k: torch.from_numpy(v).to(dev) for k, v in cols.items()}
    steps, warm = max(1, min(args.steps, 3)), 1

    def step_device():
        out = fn(devc["x"], devc["y"], devc["h"], [devc["w"]], raster)[0]
        if world > 1:
            parallel.combine_image(out, dst=0)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        step_device()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_device()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())
    e2e = None
    if not args.no_e2e and world == 1:
        kern = s.CubicSplineKernel()
        wf = kern.get_column_kernel_func(1000) if wl["lut"] else kern.w
        call = s.B200Backend.interpolate_2d_render if wl["key"] == "exact2d" else s.B200Backend.interpolate_3d_projection
        call(cols["x"], cols["y"], cols["w"], cols["h"], wf, 2.0, wl["nx"], wl["ny"], *wl["lim"], True)
        t0 = time.perf_counter()
        for _ in range(steps):
            call(cols["x"], cols["y"], cols["w"], cols["h"], wf, 2.0, wl["nx"], wl["ny"], *wl["lim"], True)
        t_e2e = (time.perf_counter() - t0) / steps
        e2e = {"value": wl["n"] / t_e2e, "unit": "particles/s", "h2d_bytes_per_step": sum(v.nbytes for v in cols.values()),
               "d2h_bytes_per_step": wl["nx"] * wl["ny"] * 8, "ms_per_step": t_e2e * 1e3, "steps": steps,
               "host_memory": "pageable NumPy arrays in; result in a new NumPy array"}
    # pixel integrals: every pixel of each particle's box (2-D); pixels passing the reference's cull (3-D) are fewer
    h = devc["h"]
    pw = 1.0 / wl["nx"]
    boxes = float((((4.0 * h / pw).ceil() + 1.0) ** 2).sum().item())
    del devc
    ops.release_workspaces()
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else FALLBACK_HBM_GBS
    ab = alg_bytes(wl, 8)
    rec = {"metric": "particles/sec", "value": wl["n"] / (ms_step * 1e-3), "unit": "particles/s", "n_gpus": world,
           "steps": steps, "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": wl["desc"], "particles_total": wl["n"],
                      "bbox_pixels_upper_bound": boxes * world if world > 1 else boxes},
           "roofline": {"bound": "hbm", "achieved": ab / (ms_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": ab / (ms_step * 1e-3) / 1e9 / peak, "traffic": None, "alg_bytes": ab,
                        "kernel": "exact_lattice_kernel<PROJ3=%s>" % ("false" if wl["key"] == "exact2d" else "true"),
                        "note": "compute-bound by construction: 10^3-10^4 double-precision operations (closed forms with "
                                "log / atan / acos) per pixel integral; the HBM fraction is reported for uniformity"},
           "e2e": e2e, "gpu_launches": steps * (3 if wl["key"] == "exact2d" else 2)}  # count (+ lines) + lattice kernel
    if want_cpu:
        cpu = ctx["cpu"]
        ns = 1 << (14 if wl["key"] == "exact2d" else 12)
        cpu.run(wl, cols, 256)                       # JIT, untimed
        tt = cpu.run(wl, cols, ns)
        ns2 = int(min(wl["n"], max(ns, ns * args.cpu_budget / max(tt, 1e-3))))
        if ns2 > 1.5 * ns:
            ns, tt = ns2, cpu.run(wl, cols, ns2)
        rec["cpu_baseline"] = {"value": ns / tt, "unit": "particles/s", "cores": cpu.threads, "kind": cpu.kind,
                               "seconds": tt, "sample_particles": ns, "sample_fraction": ns / wl["n"],
                               "sample": f"first {ns} of {wl['n']} particles of the same set, same raster, {tt:.2f} s; {cpu.note}"}
    return rec


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL on a high-priority stream: the reduction of a finished z-slab must get SMs while the
        # render kernels of the next slabs keep the GPU full (sarracen_b200/parallel.py)
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    ctx = {"dev": dev, "world": world, "rank": rank}
    want_cpu = (not args.no_cpu) and world == 1   # rank 0 at N = 1 only
    if want_cpu:
        ctx["cpu"] = CpuReference(prefer_port=args.cpu_port)

    names = [args.workload] + [w for w in (args.sub.split(",") if args.sub not in ("", "none") else [])
                               if w and w != args.workload]
    records, shared = {}, None
    for i, name in enumerate(names):
        wl = describe(name, args.particles or default_particles(name), args.dtype)
        if args.emulate_shard > 1 and world == 1:
            # experiments: one GPU renders the shard rank 0 of an N-rank run would get (value / N is meaningless)
            cols = generate(wl, 0, args.emulate_shard, args.dtype)
        else:
            cols = generate(wl, rank, world, args.dtype, share=shared if name in ("grid", "grid_hmin") else None)
        if name in ("grid", "grid_hmin"):
            shared = cols
        sampler = ClockSampler(local) if (i == 0 and rank == 0 and not os.environ.get("BENCH_NO_CLOCKS")) else None
        if name in ("exact2d", "exact3d"):
            records[name] = measure_exact(wl, cols, args, ctx, want_cpu)
        else:
            records[name] = measure(wl, cols, args, ctx, want_cpu, sampler)
        del cols
    if rank == 0:
        line = records[names[0]]
        if len(names) > 1:
            line["workloads"] = {n: records[n] for n in names[1:]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="proj", choices=["proj", "grid", "grid_hmin"])
    ap.add_argument("--sub", default="grid,grid_hmin,exact2d,exact3d",
                    help="comma list of further workloads reported as sub-records under `workloads` ('none')")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--particles", type=int, default=0, help="override the TOTAL particle count of every workload")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end legs (profiling runs)")
    ap.add_argument("--emulate-shard", type=int, default=1,
                    help="experiments only: render rank 0's shard of an N-rank run on one GPU")
    ap.add_argument("--cpu-port", action="store_true", help="CPU legs: use the C port even if Sarracen imports")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work per cpu_baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
