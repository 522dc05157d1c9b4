"""Import the unmodified reference (sarracen) from /root/reference.

TEST INFRASTRUCTURE ONLY, and only usable in the authoring container:
/root/reference does not exist on the GPU box, so nothing that runs there
(-m gpu tests, smoke(), bench.py) may call this.  It is used by
tests/golden/make_golden.py to generate the committed golden vectors and by the
``-m "not gpu"`` oracle-vs-reference tests, which skip when the reference is
absent.

The reference imports matplotlib and seaborn at package import time
(sarracen/__init__.py:22,26); neither is installed and neither is on the
scatter path, so empty stand-in modules are registered first.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SARRACEN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sarracen"))


def _stub(name, **attrs):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    for k, v in attrs.items():
        if not hasattr(mod, k):
            setattr(mod, k, v)
    return mod


def load_reference():
    """Return the reference ``sarracen`` module (raises if it is absent)."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if "sarracen" in sys.modules:
        return sys.modules["sarracen"]
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
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import sarracen
    return sarracen
