#!/usr/bin/env python
"""Benchmark of the SPH particle -> grid scatter path (BASELINE.json metric:
particles/sec for interpolate_3d_proj / interpolate_3d_grid, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload proj|grid|grid_hmin]
                    [--dtype f64|f32] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic particles:

  proj       BASELINE configs[1]: interpolate_3d_proj column render, 2^24 Plummer-sphere
             particles, quintic kernel + integrated column table (1000 samples), 2048 x 2048.
  grid       BASELINE configs[3]: interpolate_3d_grid, 2^27 uniform particles -> 512^3 voxels,
             cubic kernel, h = 1.2 N^-1/3 (2h / voxel = 2.44).
  grid_hmin  the HBM-bound variant of configs[3]: h = 0.3 voxel with hmin=True,
             i.e. clamped to 0.5 voxel.

`value` times the device pipeline (bin + sort + render) with inputs resident in HBM; `e2e`
times the reference-facing call `B200Backend.interpolate_3d_projection / _grid` with pinned
HOST buffers, host<->device copies inside the timed region.  With N GPUs every rank renders
its own particle shard (weak scaling) and the partial grids are combined with an NCCL reduce
(images) or reduce-scatter along z (voxel grids) inside the timed region.

`--impl reference` times the CPU oracle port of the reference algorithm (oracle/) on the
host cores over a bounded sample of the same workload.
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
def plummer(n, seed, a=1.0, rmax=10.0):
    """draw order positions (u, cos theta, phi), then A."""
    rng = np.random.default_rng(seed)
    u = rng.uniform(0.0, (1 + a * a / rmax ** 2) ** -1.5, n)
    ct = rng.uniform(-1.0, 1.0, n)
    ph = rng.uniform(0.0, 2 * np.pi, n)
    A = rng.uniform(0.0, 1.0, n)
    r = a / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    st = np.sqrt(1.0 - ct * ct)
    rho = 3.0 / (4 * np.pi * a ** 3) * (1 + r * r / (a * a)) ** -2.5
    m = 1.0 / n
    h = 1.2 * (m / rho) ** (1.0 / 3.0)
    w = A * m / rho  # weight A m / rho  (interpolate.py:468-472)
    return dict(x=r * st * np.cos(ph), y=r * st * np.sin(ph), h=h, w=w)


def uniform_cube(n, seed, h_value):
    """SURVEY 8(d) config 4."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.0, 1.0, n)
    y = rng.uniform(0.0, 1.0, n)
    z = rng.uniform(0.0, 1.0, n)
    A = rng.uniform(0.0, 1.0, n)
    return dict(x=x, y=y, z=z, h=np.full(n, h_value), w=A * (1.0 / n) / 1.0)


def make_workload(name, n_override, rank, dtype):
    np_dtype = np.float64 if dtype == "f64" else np.float32
    if name == "proj":
        n = n_override or (1 << 24)
        cols = plummer(n, 1002 + rank)
        wl = dict(name="interpolate_3d_proj", kind="quintic", lut=True, radius=3.0, ndim=2, fields=4,
                  nx=2048, ny=2048, nz=1, lim=(-2.0, 2.0, -2.0, 2.0),
                  desc=f"interpolate_3d_proj: {n} Plummer-sphere particles, quintic column table (1000 "
                       f"samples), 2048x2048, {dtype}")
    else:
        n = n_override or (1 << 27)
        m = 512 if n >= (1 << 26) else max(32, int(round((n / (1 << 27)) ** (1 / 3) * 512 / 32)) * 32)
        if name == "grid":
            hval = 1.2 * n ** (-1.0 / 3.0)
            tag = "h = 1.2 N^-1/3"
        else:
            hval = max(0.3 / m, 0.5 / m)  # hmin=True clamps 0.3 voxel to 0.5 voxel
            tag = "h = 0.3 voxel clamped by hmin to 0.5 voxel"
        cols = uniform_cube(n, 1004 + rank, hval)
        wl = dict(name="interpolate_3d_grid", kind="cubic", lut=False, radius=2.0, ndim=3, fields=5,
                  nx=m, ny=m, nz=m, lim=(0.0, 1.0, 0.0, 1.0, 0.0, 1.0),
                  desc=f"interpolate_3d_grid: {n} uniform particles -> {m}^3 voxels, cubic, {tag}, {dtype}")
    wl["n"] = n
    wl["cols"] = {k: np.ascontiguousarray(v.astype(np_dtype)) for k, v in cols.items()}
    wl["itemsize"] = np.dtype(np_dtype).itemsize
    # algorithmic bytes: every particle field read once, every output element written once
    wl["alg_bytes"] = n * wl["fields"] * wl["itemsize"] + wl["nx"] * wl["ny"] * wl["nz"] * 8
    return wl


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


# ------------------------------------------------------------------ CPU baseline (oracle port)
def cpu_baseline(wl, budget_s=15.0):
    """Time the CPU oracle (port of the reference algorithm, all host threads) on a bounded
    sample of the workload: the first n_s particles of the same arrays, same raster."""
    import oracle
    from oracle import OracleBackend, OracleKernel
    cols = wl["cols"]
    # all the host threads this process may use, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    oracle.set_num_threads(avail)
    threads = oracle.num_threads()

    def run(ns, m=None):
        sl = {k: v[:ns] for k, v in cols.items()}
        t0 = time.perf_counter()
        if wl["nz"] == 1:
            kern = OracleKernel.column(wl["kind"], 1000)
            OracleBackend.interpolate_3d_projection(sl["x"], sl["y"], sl["w"], sl["h"], kern, wl["radius"],
                                                    wl["nx"], wl["ny"], *wl["lim"], False)
        else:
            OracleBackend.interpolate_3d_grid(sl["x"], sl["y"], sl["z"], sl["w"], sl["h"], OracleKernel(wl["kind"]),
                                              wl["radius"], m, m, m, *wl["lim"])
        return time.perf_counter() - t0

    if wl["nz"] == 1:
        ns = min(wl["n"], 1 << 14)
        t = run(ns)
        t = run(ns)  # second call: threads and pages warm
        scale = max(1.0, min(budget_s / max(t, 1e-3), wl["n"] / ns))
        ns2 = int(ns * scale)
        t2 = run(ns2) if ns2 > ns else t
        if t2 < 0.6 * budget_s and ns2 < wl["n"]:
            # the small calibration run overestimates the cost per particle: scale once more
            ns3 = int(min(wl["n"], ns2 * budget_s / max(t2, 1e-3)))
            if ns3 > ns2:
                ns2, t2 = ns3, run(ns3)
        return {"value": ns2 / t2, "unit": "particles/s", "cores": threads, "kind": "port", "seconds": t2,
                "sample": f"first {ns2} of {wl['n']} particles of the same set, same {wl['nx']}x{wl['ny']} raster, "
                          f"{t2:.2f} s, oracle/sph_oracle.c with {threads} OpenMP threads"}
    # voxel grid: the reference makes nz full passes over the particles (cpu_backend.py:213-218), far too
    # slow at full size; time a geometrically similar problem (same 2h / voxel), SURVEY 8(d).
    m = 64
    ratio = (wl["nx"] / m) ** 3
    ns = max(1024, int(wl["n"] / ratio))
    sl_h = cols["h"][0] * (wl["nx"] / m)
    saved = cols["h"]
    cols["h"] = np.full_like(saved, sl_h)
    try:
        t = run(ns, m)
        if t < budget_s / 6 and m < 128:
            m = 128
            ratio = (wl["nx"] / m) ** 3
            ns = max(1024, int(wl["n"] / ratio))
            cols["h"] = np.full_like(saved, saved[0] * (wl["nx"] / m))
            t = run(ns, m)
    finally:
        cols["h"] = saved
    return {"value": ns / t, "unit": "particles/s", "cores": threads, "kind": "port", "seconds": t,
            "sample": f"{ns} particles -> {m}^3 voxels with the same 2h/voxel ({t:.2f} s); the reference does nz "
                      f"full passes so particles/s falls as 1/nz: at {wl['nx']}^3 expect x{m / wl['nx']:.3f}, "
                      f"oracle/sph_oracle.c with {threads} OpenMP threads"}


# ------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args.workload, args.particles, 0, args.dtype)
    per_step, secs = [], []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(wl, budget_s=max(2.0, 60.0 / (args.warmup + args.steps)))
        if i >= args.warmup:
            per_step.append(base["value"])
            secs.append(base["seconds"])
    value = float(np.mean(per_step))
    base["value"] = value
    line = {"impl": "reference", "metric": "particles/sec", "value": value, "unit": "particles/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(secs)) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic", "config": {"workload": wl["desc"]}, "cpu_baseline": base,
            "e2e": {"value": value, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------ our arm
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
        dist.init_process_group("nccl", device_id=dev)

    import sarracen_b200 as s
    from sarracen_b200 import ops
    from sarracen_b200.parallel import combine_image, combine_grid_along_z

    wl = make_workload(args.workload, args.particles, rank, args.dtype)
    kernel = {"cubic": s.CubicSplineKernel, "quartic": s.QuarticSplineKernel, "quintic": s.QuinticSplineKernel}[
        wl["kind"]]()
    wf = kernel.get_column_kernel_func(1000) if wl["lut"] else kernel.w
    dim3 = wl["nz"] > 1
    raster = ops.Raster(wl["nx"], wl["ny"], *wl["lim"][:4], wl["nz"], *(wl["lim"][4:] if dim3 else (0.0, 1.0)))

    host = {k: torch.from_numpy(v).pin_memory() for k, v in wl["cols"].items()}
    devc = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    out_shape = (wl["nz"], wl["ny"], wl["nx"]) if dim3 else (wl["ny"], wl["nx"])
    out = [torch.empty(out_shape, dtype=torch.float64, device=dev)]
    if dim3 and world > 1:
        slab = torch.empty((wl["nz"] // world, wl["ny"], wl["nx"]), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def step_device():
        if dim3:
            ops.scatter3d(devc["x"], devc["y"], devc["z"], devc["h"], [devc["w"]], wf, wl["radius"], raster, out=out)
            if world > 1:
                combine_grid_along_z(out[0], slab)
        else:
            ops.scatter2d(devc["x"], devc["y"], devc["h"], [devc["w"]], wf, wl["radius"], wl["ndim"], raster,
                          out=out)
            if world > 1:
                combine_image(out[0], dst=0)

    slab_shape = (wl["nz"] // world, wl["ny"], wl["nx"]) if (dim3 and world > 1) else out_shape
    host_out = torch.empty(slab_shape, dtype=torch.float64).pin_memory()

    def step_e2e():
        if world > 1:
            # N ranks: pinned host columns -> device, partial raster, NCCL combine, result back to the
            # host (rank 0's image, or every rank's z-slab of the grid) -- the multi-GPU recipe of
            # INTEGRATION.md section 4 with the host copies inside the timed region
            if dim3:
                part = s.backend.streamed_scatter(
                    [host[k] for k in "xyzwh"], dev, lambda c, out, acc: ops.scatter3d(
                        c[0], c[1], c[2], c[4], [c[3]], wf, wl["radius"], raster, out=out, accumulate=acc))[0]
                combine_grid_along_z(part, slab)
                host_out.copy_(slab)
            else:
                part = s.backend.streamed_scatter(
                    [host[k] for k in "xywh"], dev, lambda c, out, acc: ops.scatter2d(
                        c[0], c[1], c[3], [c[2]], wf, wl["radius"], wl["ndim"], raster, out=out, accumulate=acc))[0]
                combine_image(part, dst=0)
                if rank == 0:
                    host_out.copy_(part)
            torch.cuda.synchronize()
            return host_out
        # one GPU: the call a Sarracen user reaches through backend='gpu': host arrays in, host array out
        if dim3:
            img = s.B200Backend.interpolate_3d_grid(host["x"], host["y"], host["z"], host["w"], host["h"], wf,
                                                    wl["radius"], wl["nx"], wl["ny"], wl["nz"], *wl["lim"])
        else:
            img = s.B200Backend.interpolate_3d_projection(host["x"], host["y"], host["w"], host["h"], wf,
                                                          wl["radius"], wl["nx"], wl["ny"], *wl["lim"], False)
        return img

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------
    ops.time_phases = True
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("BENCH_NO_CLOCKS"):
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phases = {"ms_bin": 0.0, "ms_sort": 0.0, "ms_render": 0.0, "ms_direct": 0.0}
    launches = 0
    e0.record()
    dbg = []
    for _ in range(args.steps):
        t_dbg = time.perf_counter()
        step_device()
        dbg.append(round((time.perf_counter() - t_dbg) * 1e3, 2))
        for k in phases:
            phases[k] += ops.last_stats.get(k, 0.0)
        launches += ops.last_stats.get("launches", 0)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    if os.environ.get("BENCH_DEBUG"):
        print("per-step host ms:", dbg, file=sys.stderr)
    clocks = sampler.stop() if rank == 0 else None
    n_pairs = ops.last_stats.get("n_pairs", 0)
    n_direct = ops.last_stats.get("n_direct", 0)
    tile = ops.last_stats.get("tile")
    ops.time_phases = False

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = wl["n"] * world / (ms_step * 1e-3)

    # ---- end-to-end timing through the backend API -----------------------------
    e2e_steps = max(1, min(args.steps, 3))
    # two warm-ups that keep the previous result alive, like the timed loop does: the backend returns
    # the raster in page-locked memory from torch's caching host allocator, and the steady state
    # alternates between two such blocks (a first-time 1 GiB page-locked allocation costs ~0.4 s)
    img = step_e2e()
    img = step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        img = step_e2e()
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    h2d = sum(v.numel() * v.element_size() for k, v in host.items())
    d2h = int(np.prod(slab_shape)) * 8

    # ---- contributions (outside the timed region) ------------------------------
    contrib = ops.count_contributions(devc["x"], devc["y"], devc["h"], wl["radius"], raster,
                                      z=devc.get("z") if dim3 else None)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    # the dominant kernel: tile render or direct, whichever took longer
    dom = "ms_render" if phases["ms_render"] >= phases["ms_direct"] else "ms_direct"
    ms_render = phases[dom] / args.steps
    achieved = wl["alg_bytes"] / (ms_render * 1e-3) / 1e9 if ms_render > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(f"{args.workload}_{args.dtype}")

    base = cpu_baseline(wl) if (not args.no_cpu and world == 1) else None  # rank 0 at N = 1 only

    # What bounds the dominant kernel in practice: at SPH-consistent smoothing
    # lengths the per-contribution arithmetic, not HBM.  Instruction counts per contribution are read
    # off the SASS of the hot loops (profiles/); peak = 148 SMs x 64 FP64 (128 FP32) lanes x SM clock.
    sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
    rate = contrib / (ms_render * 1e-3) if ms_render > 0 else 0.0
    if args.dtype == "f64":
        # render: 8 per evaluation in the hot loop, ~12 per useful contribution with idle lanes and the
        # per-hit set-up (ncu: fp64 pipe 49.5 % at 84 ms, profiles/r1_proj_f64_full_v9.txt)
        dp_per = 12.0 if dom == "ms_render" else 16.0
        compute = {"bound": "fp64 pipe + issue slots", "fp64_inst_per_contribution_est": dp_per,
                   "achieved_fp64_inst_per_s": rate * dp_per, "peak_fp64_inst_per_s": 148 * 64 * sm_clock * 1e6,
                   "frac_est": rate * dp_per / (148 * 64 * sm_clock * 1e6)}
        if dom == "ms_direct":
            # the direct kernel issues one red.global.add.f64 per contribution: spread-address REDG
            # costs 1.29 cycles per lane per SM (B300_MICROARCH.md, atomics table)
            red_peak = 148 * sm_clock * 1e6 / 1.29
            compute.update({"bound": "REDG rate + issue slots", "achieved_red_per_s": rate,
                            "peak_red_per_s": red_peak, "red_frac": rate / red_peak})
    else:
        # thread-instruction slots per contribution: the double kernel executes 5.78e10 warp instructions
        # for 6.2e10 contributions (29.8 slots each, ncu v9); the float loop is 11.5 instead of 14.5
        # instructions per evaluation -- not captured by ncu itself, hence an estimate
        per = 26.0
        compute = {"bound": "issue slots", "inst_per_contribution_est": per,
                   "achieved_inst_per_s": rate * per, "peak_inst_per_s": 148 * 128 * sm_clock * 1e6,
                   "frac_est": rate * per / (148 * 128 * sm_clock * 1e6)}

    line = {
        "metric": "particles/sec", "value": value, "unit": "particles/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": wl["desc"], "particles_per_gpu": wl["n"], "l2": "inputs larger than L2 "
                   f"({h2d / 1e6:.0f} MB of particle columns per step)", "tile": tile, "pairs": n_pairs,
                   "direct_particles": n_direct,
                   "parallelism": f"particle-sharded x{world}" + ("" if world == 1 else
                                                                  (", NCCL reduce-scatter along z" if dim3
                                                                   else ", NCCL reduce to rank 0"))},
        "pixel_contributions": contrib, "contributions_per_s": contrib * world / (ms_step * 1e-3),
        "phases_ms": {k: v / args.steps for k, v in phases.items()},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "kernel": "render_kernel" if dom == "ms_render" else "direct_kernel",
                     "alg_bytes": wl["alg_bytes"], "peak_source": peak_src,
                     "whole_step_frac": wl["alg_bytes"] / (ms_step * 1e-3) / 1e9 / peak},
        "compute_bound": compute,
        "cpu_baseline": base,
        "e2e": {"value": wl["n"] * world / t_e2e, "unit": "particles/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e * 1e3},
        "gpu_launches": launches, "clocks": clocks,
    }
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
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--particles", type=int, default=0, help="override the particle count per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
