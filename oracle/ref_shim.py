"""Import the unmodified reference (sarracen).

TEST / BENCH INFRASTRUCTURE ONLY -- nothing under sarracen_b200/ imports this.

Search order: $SARRACEN_REFERENCE_ROOT, then `baseline/_ref/` (the
reference installed by baseline/install_ref.py; git-ignored, travels to the GPU box with the
gpurun snapshot), then `/root/reference` (authoring container only).  Users:
tests/golden/make_golden*.py (golden vectors), the oracle-vs-reference tests, the reference
suite routed through backend='gpu' (tests/test_reference_suite_gpu.py), and bench.py's CPU legs
(`cpu_baseline.kind = "reference"`: Sarracen's own numba CPUBackend timed on the host cores).

The reference imports matplotlib and seaborn at package import time
(sarracen/__init__.py:22,26); neither is installed and neither is on the scatter path, so
empty stand-in modules are registered first.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _candidates():
    env = os.environ.get("SARRACEN_REFERENCE_ROOT")
    if env:
        yield env
    yield os.path.join(_REPO, "baseline", "_ref")
    yield "/root/reference"


def reference_root():
    """Directory that holds the `sarracen` package, or None."""
    for root in _candidates():
        if os.path.isfile(os.path.join(root, "sarracen", "__init__.py")):
            return root
    return None


def available() -> bool:
    return reference_root() is not None


def _stub(name, **attrs):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    for k, v in attrs.items():
        if not hasattr(mod, k):
            setattr(mod, k, v)
    return mod


def stub_plotting() -> None:
    """Register empty matplotlib / seaborn modules when the real ones are absent."""
    try:
        import matplotlib  # noqa: F401
        import seaborn  # noqa: F401
    except ImportError:
        dummy = type("_Dummy", (), {})
        mpl = _stub("matplotlib")
        mpl.axes = _stub("matplotlib.axes", Axes=dummy)
        mpl.colors = _stub("matplotlib.colors", Colormap=dummy, LogNorm=dummy, SymLogNorm=dummy)
        mpl.pyplot = _stub("matplotlib.pyplot")
        _stub("seaborn")


def load_reference():
    """Return the reference ``sarracen`` module (raises if it is absent)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not found (looked in $SARRACEN_REFERENCE_ROOT, baseline/_ref, /root/reference)")
    if "sarracen" in sys.modules:
        return sys.modules["sarracen"]
    stub_plotting()
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
    import sarracen
    return sarracen
