"""The C-ABI library loads and exports every symbol include/sphb200.h declares
(no compute calls: this runs without a GPU), and argument validation works."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sphb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sphb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from sarracen_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 11
    for name in names:
        assert hasattr(lib, name), f"{name} declared in sphb200.h but not exported"
    assert sorted(_lib.EXPORTS) == names
    assert lib.sphb_version() == _lib.SPHB_VERSION == 200


def test_struct_layouts_match_the_header(tmp_path):
    """Compile the header with gcc and compare struct sizes with the ctypes mirrors."""
    import subprocess
    from sarracen_b200 import _lib
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "sphb200.h"\nint main(void){printf("%zu %zu %zu %zu\\n",'
                   'sizeof(sphb_particles),sizeof(sphb_kernel),sizeof(sphb_raster),sizeof(sphb_stats));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(t) for t in subprocess.check_output([str(exe)]).split()]
    mirrors = [_lib.SphbParticles, _lib.SphbKernel, _lib.SphbRaster, _lib.SphbStats]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]


def test_invalid_arguments_are_rejected_without_a_device():
    from sarracen_b200 import _lib
    lib = _lib.load()
    rc = lib.sphb_column_kernel(_lib.SPHB_CUBIC, 1, None, None)
    assert rc == _lib.SPHB_ERR_ARGUMENT
    assert b"samples" in lib.sphb_last_error()
    p, k, r = _lib.SphbParticles(), _lib.SphbKernel(), _lib.SphbRaster()
    p.n, p.dtype, p.n_weights = 0, 7, 1
    rc = lib.sphb_scatter2d(ctypes.byref(p), ctypes.byref(k), ctypes.byref(r), 0, 0.0, 0,
                            (ctypes.c_void_p * 1)(), None, 0, None, None, None)
    assert rc == _lib.SPHB_ERR_ARGUMENT


def test_ops_refuse_cpu_tensors():
    import torch
    from sarracen_b200 import CubicSplineKernel, ops
    t = torch.zeros(4, dtype=torch.float64)
    with pytest.raises(TypeError):
        ops.scatter2d(t, t, t, [t], CubicSplineKernel.w, 2.0, 2, ops.Raster(4, 4, 0, 1, 0, 1))


def test_backend_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from sarracen_b200 import B200Backend, CubicSplineKernel
    a = np.zeros(3)
    with pytest.raises(RuntimeError):
        B200Backend.interpolate_2d_render(a, a, a, a + 1, CubicSplineKernel.w, 2, 4, 4, 0, 1, 0, 1, False)


def test_weight_function_resolution():
    from sarracen_b200 import CubicSplineKernel, WeightFunction, resolve_weight_function
    wf = resolve_weight_function(CubicSplineKernel.w, 2)
    assert wf.kind == "cubic"

    class Duck:
        kind, radius, lut = "lut", 3.0, np.linspace(1, 0, 8)
    wf = resolve_weight_function(Duck(), 3)
    assert isinstance(wf, WeightFunction) and wf.kind == "lut" and wf.lut.size == 8
    with pytest.raises(TypeError):
        resolve_weight_function(lambda q, d: 1.0, 2)


def test_host_kernel_values():
    """kernels.WeightFunction.__call__ against the oracle (closed forms)."""
    from oracle import kernel_w
    from sarracen_b200 import CubicSplineKernel, QuarticSplineKernel, QuinticSplineKernel
    for K, name in ((CubicSplineKernel, "cubic"), (QuarticSplineKernel, "quartic"), (QuinticSplineKernel, "quintic")):
        for q in (0.0, 0.3, 0.5, 1.0, 1.4, 1.5, 2.0, 2.2, 2.5, 2.9, 3.0, 3.5, -1.0):
            for d in (1, 2, 3):
                assert K.w(q, d) == pytest.approx(kernel_w(name, q, d), rel=1e-14, abs=1e-17)
