"""ctypes binding of libsphb200.so (the C ABI declared in include/sphb200.h).

The library is the product: if it is missing or fails to load, every op raises
-- there is no CPU or PyTorch fallback on this path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPHB_LIB_OVERRIDE") or os.path.join(_HERE, "lib", "libsphb200.so")

SPHB_MAX_WEIGHTS = 4
SPHB_F32, SPHB_F64 = 0, 1
SPHB_CUBIC, SPHB_QUARTIC, SPHB_QUINTIC, SPHB_COLUMN_LUT = 0, 1, 2, 3
SPHB_OK, SPHB_ERR_ARGUMENT, SPHB_ERR_WORKSPACE, SPHB_ERR_CUDA, SPHB_ERR_TOO_MANY = 0, -1, -2, -3, -4

EXPORTS = ["sphb_version", "sphb_last_error", "sphb_column_kernel", "sphb_scatter2d", "sphb_scatter3d",
           "sphb_scatter3d_staged", "sphb_scatter_plan", "sphb_scatter_planned", "sphb_scatter_overflowed",
           "sphb_line2d", "sphb_line3d", "sphb_exact2d", "sphb_exact3d_proj", "sphb_normalize", "sphb_sum_slots",
           "sphb_count_contributions", "sphb_radial_index", "sphb_bin_sums"]


class SphbParticles(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int64), ("dtype", ctypes.c_int),
                ("x", ctypes.c_void_p), ("y", ctypes.c_void_p), ("z", ctypes.c_void_p),
                ("h", ctypes.c_void_p), ("n_weights", ctypes.c_int),
                ("w", ctypes.c_void_p * SPHB_MAX_WEIGHTS)]


class SphbKernel(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("ndim", ctypes.c_int), ("radius", ctypes.c_double),
                ("lut", ctypes.c_void_p), ("lut_samples", ctypes.c_int)]


class SphbRaster(ctypes.Structure):
    _fields_ = [("nx", ctypes.c_int), ("ny", ctypes.c_int), ("nz", ctypes.c_int),
                ("x_min", ctypes.c_double), ("x_max", ctypes.c_double),
                ("y_min", ctypes.c_double), ("y_max", ctypes.c_double),
                ("z_min", ctypes.c_double), ("z_max", ctypes.c_double)]


class SphbStats(ctypes.Structure):
    _fields_ = [("n_pairs", ctypes.c_int64), ("n_direct", ctypes.c_int64), ("n_tiles", ctypes.c_int64),
                ("tile_x", ctypes.c_int32), ("tile_y", ctypes.c_int32), ("tile_z", ctypes.c_int32),
                ("time_phases", ctypes.c_int32), ("ms_bin", ctypes.c_float), ("ms_sort", ctypes.c_float),
                ("ms_render", ctypes.c_float), ("ms_direct", ctypes.c_float), ("launches", ctypes.c_int32),
                ("ms_count", ctypes.c_float)]


class SphbPlan(ctypes.Structure):
    _fields_ = [("n_small", ctypes.c_uint64), ("n_pairs", ctypes.c_uint64), ("max_small_edge", ctypes.c_uint64),
                ("reserved0", ctypes.c_uint64), ("n_large", ctypes.c_uint64), ("reserved", ctypes.c_uint64 * 3)]


class SphbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libsphb200 error {code}: {message}")
        self.code = code


STAGE_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int)
SPHB_MAX_STAGES = 64
SPHB_VERSION = 200

_lib = None


def load():
    """Load libsphb200.so; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m sarracen_b200.build` "
            "(nvcc, sm_100a).  sarracen_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    P, K, R, S = (ctypes.POINTER(SphbParticles), ctypes.POINTER(SphbKernel), ctypes.POINTER(SphbRaster),
                  ctypes.POINTER(SphbStats))
    outs = ctypes.POINTER(ctypes.c_void_p)
    c_int, c_dbl, c_vp, c_sz = ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t
    lib.sphb_version.restype = c_int
    lib.sphb_version.argtypes = []
    lib.sphb_last_error.restype = ctypes.c_char_p
    lib.sphb_last_error.argtypes = []
    lib.sphb_column_kernel.argtypes = [c_int, c_int, c_vp, c_vp]
    lib.sphb_scatter2d.argtypes = [P, K, R, c_int, c_dbl, c_int, outs, c_vp, c_sz,
                                   ctypes.POINTER(c_sz), S, c_vp]
    lib.sphb_scatter3d.argtypes = [P, K, R, c_int, outs, c_vp, c_sz, ctypes.POINTER(c_sz), S, c_vp]
    lib.sphb_scatter3d_staged.argtypes = [P, K, R, c_int, outs, c_vp, c_sz, ctypes.POINTER(c_sz), S, c_int,
                                          ctypes.POINTER(c_int), STAGE_FN, c_vp, c_vp]
    PL = ctypes.POINTER(SphbPlan)
    lib.sphb_scatter_plan.argtypes = [P, K, R, c_int, c_int, c_dbl, c_vp, c_sz, ctypes.POINTER(c_sz), c_vp, c_vp]
    lib.sphb_scatter_planned.argtypes = [P, K, R, c_int, c_int, c_dbl, c_int, outs, c_vp, c_sz, ctypes.POINTER(c_sz),
                                         PL, c_vp]
    lib.sphb_scatter_overflowed.argtypes = [c_vp, ctypes.POINTER(c_int), c_vp]
    lib.sphb_line2d.argtypes = [P, K, c_int, c_dbl, c_dbl, c_dbl, c_dbl, c_int, outs, c_vp]
    lib.sphb_line3d.argtypes = [P, K, c_int, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_int, outs, c_vp]
    lib.sphb_exact2d.argtypes = [P, R, c_int, outs, c_vp, c_sz, ctypes.POINTER(c_sz), c_vp]
    lib.sphb_exact3d_proj.argtypes = [P, R, c_int, outs, c_vp, c_sz, ctypes.POINTER(c_sz), c_vp]
    lib.sphb_radial_index.argtypes = [c_int, c_vp, c_vp, c_vp, ctypes.c_int64, ctypes.POINTER(c_dbl), c_int, c_vp, c_int,
                                      c_vp, c_vp, c_vp]
    lib.sphb_bin_sums.argtypes = [c_int, c_vp, c_int, outs, ctypes.c_int64, c_int, outs, c_vp]
    lib.sphb_normalize.argtypes = [c_vp, c_vp, ctypes.c_int64, c_vp]
    lib.sphb_sum_slots.argtypes = [c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, c_vp]
    lib.sphb_count_contributions.argtypes = [P, c_dbl, R, c_int, c_dbl, c_vp, c_vp]
    for name in EXPORTS:
        if name != "sphb_last_error":
            getattr(lib, name).restype = c_int
    _lib = lib
    return lib


def check(code):
    if code != SPHB_OK:
        raise SphbError(code, load().sphb_last_error().decode())
