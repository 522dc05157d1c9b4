"""Build libsphb200.so (CUDA, sm_100a) in-tree with nvcc.

    python -m sarracen_b200.build [--force]

The library is a plain C-ABI shared object (include/sphb200.h); it does not
link against torch.  Objects are compiled in parallel, one nvcc per .cu file.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
BUILD_DIR = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(OUT_DIR, "libsphb200.so")
SOURCES = ["sphb_scatter.cu", "sphb_render.cu", "sphb_cell.cu", "sphb_misc.cu", "sphb_exact.cu", "sphb_bins.cu"]
HEADERS = [os.path.join(CSRC, "sphb_common.cuh"), os.path.join(CSRC, "sphb_scatter.cuh"),
           os.path.join(os.path.dirname(HERE), "include", "sphb200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
              "-Xptxas", "-v"]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; libsphb200.so cannot be built")
    return cand


def _stale(target, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, debug_bounds: bool = False, defines=(), suffix: str = "") -> str:
    """debug_bounds=True builds lib/libsphb200_dbg.so with -DSPHB_DEBUG_BOUNDS (device asserts on
    every index the kernels form); point SPHB_LIB_OVERRIDE at it to run the GPU tests against it."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    nvcc = _nvcc()
    # `defines` / `suffix`: experimental variants (python -m sarracen_b200.build --suffix _x -DNAME=1)
    suffix = suffix or ("_dbg" if debug_bounds else "")
    flags = NVCC_FLAGS + (["-DSPHB_DEBUG_BOUNDS"] if debug_bounds else []) + ["-D" + d for d in defines]
    lib = LIB.replace(".so", suffix + ".so")
    jobs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(BUILD_DIR, src.replace(".cu", suffix + ".o"))
        if force or _stale(obj, [path] + HEADERS):
            jobs.append((path, obj))

    def compile_one(job):
        path, obj = job
        cmd = [nvcc] + ARCH + flags + ["-c", path, "-o", obj]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log = obj.replace(".o", ".ptxas.log")
        with open(log, "w") as f:
            f.write(res.stdout)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {path}:\n{res.stdout[-4000:]}")
        return res.stdout

    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            outs = list(ex.map(compile_one, jobs))
        if verbose:
            for o in outs:
                print(o)
    objs = [os.path.join(BUILD_DIR, s.replace(".cu", suffix + ".o")) for s in SOURCES]
    if force or jobs or _stale(lib, objs):
        cmd = [nvcc] + ARCH + ["-shared", "--cudart", "shared", "-o", lib] + objs
        subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    _suffix = sys.argv[sys.argv.index("--suffix") + 1] if "--suffix" in sys.argv else ""
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug_bounds="--debug-bounds" in sys.argv,
                defines=[a[2:] for a in sys.argv if a.startswith("-D")], suffix=_suffix))
