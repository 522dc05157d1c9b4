"""E1: the reference's OWN interpolation tests (unmodified files) with backend='gpu' routed to the
B200 backend -- SURVEY 4 "the hook the new backend inherits for free", BASELINE.md section 3.

Runs `sarracen/tests/interpolate/test_interpolate.py` and `test_rotation.py` from the installed
reference (baseline/_ref, see baseline/install_ref.py) in a child pytest with
oracle/refsuite_plugin.py loaded; skipped when no reference install is present.
"""
import os
import re
import subprocess
import sys

import pytest

from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_shim.available(), reason="reference not installed")]


def test_reference_interpolate_suite_through_backend_gpu(tmp_path):
    ref = ref_shim.reference_root()
    tests = os.path.join(ref, "sarracen", "tests", "interpolate")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""),
               PYTHONDONTWRITEBYTECODE="1")
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "oracle.refsuite_plugin", "-p", "no:cacheprovider",
           "--rootdir", str(tmp_path), "-o", "python_files=test_*.py", "-rxX", tests]
    res = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         timeout=1500)
    tail = res.stdout[-6000:]
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "reference_suite_gpu.log"), "w") as f:
            f.write(res.stdout)
    summary = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else ""
    passed = int((re.search(r"(\d+) passed", summary) or [0, 0])[1])
    failed = int((re.search(r"(\d+) failed", summary) or [0, 0])[1])
    errors = int((re.search(r"(\d+) error", summary) or [0, 0])[1])
    assert res.returncode == 0 and failed == 0 and errors == 0, tail
    # 64 backend=gpu instances in test_interpolate.py (test_pixel_arguments is cpu-only) and 6 in
    # test_rotation.py; only the three array-exact hmin tests may xfail
    xpassed = int((re.search(r"(\d+) xpassed", summary) or [0, 0])[1])
    assert passed + xpassed >= 67 and passed >= 64, tail
