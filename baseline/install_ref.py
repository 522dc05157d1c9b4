#!/usr/bin/env python
"""Install the UNMODIFIED reference (sarracen) into baseline/_ref (git-ignored, travels to the
GPU box with the gpurun snapshot).

    python baseline/install_ref.py [--force]

First choice is pip, exactly as the bench contract words it:

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref /root/reference

In this image that fails: the reference's build backend is `flit_core` (pyproject.toml:1-3) and
neither the venv nor /opt/wheelhouse holds it.  The reference is a pure-Python package, for
which a flit wheel is nothing but the package directory plus a `.dist-info` folder, so the
fall-back lays down that same layout by hand: `sarracen/` copied verbatim (no file is edited)
and `sarracen-<version>.dist-info/{METADATA,INSTALLER,RECORD}` written beside it.  The outcome
(which of the two ways worked) is recorded in baseline/_ref/INSTALL_LOG.txt and in DESIGN.md.

baseline/_ref is test/bench infrastructure only: `oracle/ref_shim.py` imports it for the
`--impl reference` arm, the `cpu_baseline` leg and the reference-suite tests.  Nothing under
sarracen_b200/ imports it.
"""
import hashlib
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("SARRACEN_REFERENCE_SRC", "/root/reference")


def _version(pkg_dir):
    with open(os.path.join(pkg_dir, "__init__.py")) as f:
        m = re.search(r'__version__\s*=\s*"([^"]+)"', f.read())
    return m.group(1) if m else "0"


def _pip() -> bool:
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
           "--find-links", "/opt/wheelhouse", "--target", TARGET, SOURCE]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode == 0 and os.path.isdir(os.path.join(TARGET, "sarracen")):
        return True
    sys.stderr.write("pip install of the reference failed (%s); laying the pure-Python package down by hand\n"
                     % (res.stdout.strip().splitlines()[-1] if res.stdout.strip() else res.returncode))
    return False


def _by_hand() -> None:
    src = os.path.join(SOURCE, "sarracen")
    dst = os.path.join(TARGET, "sarracen")
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    ver = _version(dst)
    info = os.path.join(TARGET, f"sarracen-{ver}.dist-info")
    os.makedirs(info, exist_ok=True)
    with open(os.path.join(info, "METADATA"), "w") as f:
        f.write(f"Metadata-Version: 2.1\nName: sarracen\nVersion: {ver}\n"
                "Summary: SPH analysis and visualization\n")
    with open(os.path.join(info, "INSTALLER"), "w") as f:
        f.write("baseline/install_ref.py\n")
    rows = []
    for root, _dirs, files in os.walk(dst):
        for name in sorted(files):
            p = os.path.join(root, name)
            with open(p, "rb") as fh:
                data = fh.read()
            rows.append(f"{os.path.relpath(p, TARGET)},sha256={hashlib.sha256(data).hexdigest()},{len(data)}")
    with open(os.path.join(info, "RECORD"), "w") as f:
        f.write("\n".join(rows) + "\n")


def install(force: bool = False) -> str:
    """Returns 'present', 'pip' or 'copy'.  Raises if the reference source is absent."""
    if os.path.isdir(os.path.join(TARGET, "sarracen")) and not force:
        return "present"
    if not os.path.isdir(os.path.join(SOURCE, "sarracen")):
        raise RuntimeError(f"reference source not found under {SOURCE}")
    if os.path.isdir(TARGET):
        shutil.rmtree(TARGET)
    os.makedirs(TARGET)
    how = "pip" if _pip() else "copy"
    if how == "copy":
        for leftover in os.listdir(TARGET):
            shutil.rmtree(os.path.join(TARGET, leftover), ignore_errors=True)
        _by_hand()
    with open(os.path.join(TARGET, "INSTALL_LOG.txt"), "w") as f:
        f.write(f"sarracen {_version(os.path.join(TARGET, 'sarracen'))} installed from {SOURCE} by "
                f"{'pip --target' if how == 'pip' else 'verbatim package copy (flit_core unavailable)'}\n")
    return how


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
