"""
TEST / BENCH INFRASTRUCTURE — loads the *unmodified* reference modules: from /root/reference
in the build container, else from ``oracle/_ref/`` (the same two modules byte-compiled by
``oracle/make_ref.py`` — the GPU box has no /root/reference).  Used by
``tests/golden/make_golden.py`` to generate the committed fixtures, by the tests to pin
``oracle/qnmfits_oracle.py`` against the real thing, and by ``bench.py``'s CPU arms to time
the reference's own functions.  Nothing in ``qnmfits_b200/`` may import this module.

Recipe: the reference's ``qnmfits/qnmfits.py`` and
``qnmfits/qnm.py`` import matplotlib, mpl_toolkits, h5py and the ``qnm`` PyPI
package at module scope (reference qnmfits/qnmfits.py:2,8; qnmfits/qnm.py:2,8).
None is installed here, and none is on the hot path, so they are replaced by empty
stubs; ``qnm.modes_cache`` is served by the synthetic table provider.  The package
``__init__`` (which imports sxs/spherical) is bypassed by registering a bare
package object whose ``__path__`` points at the reference directory.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("QNMFITS_REFERENCE_ROOT", "/root/reference")
COMPILED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_package_dir():
    """Directory to import ``qnmfits.qnm`` / ``qnmfits.qnmfits`` from, or None."""
    src = os.path.join(REFERENCE_ROOT, "qnmfits")
    if os.path.isfile(os.path.join(src, "qnmfits.py")):
        return src
    built = os.path.join(COMPILED_ROOT, "qnmfits")
    if os.path.isfile(os.path.join(built, "qnmfits.code")) and os.path.isfile(os.path.join(built, "qnm.code")):
        return built
    return None


def _import_reference(package_dir):
    """``qnmfits.qnmfits`` from the sources, or from the marshalled code objects of oracle/_ref
    (executed as the modules ``qnmfits.qnm`` and ``qnmfits.qnmfits`` of the bare package)."""
    if os.path.isfile(os.path.join(package_dir, "qnmfits.py")):
        return importlib.import_module("qnmfits.qnmfits")
    import marshal
    mods = {}
    for name in ("qnm", "qnmfits"):
        full = "qnmfits." + name
        mod = types.ModuleType(full)
        mod.__package__ = "qnmfits"
        mod.__file__ = os.path.join(package_dir, name + ".code")
        sys.modules[full] = mod
        setattr(sys.modules["qnmfits"], name, mod)
        with open(mod.__file__, "rb") as fh:
            exec(marshal.load(fh), mod.__dict__)
        mods[name] = mod
    return mods["qnmfits"]


def reference_available():
    return reference_package_dir() is not None


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


_loaded = None


def load_reference(modes_cache=None):
    """Return the reference's ``qnmfits.qnmfits`` module (functions + ``qnm`` instance)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    package_dir = reference_package_dir()
    if package_dir is None:
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT} nor built under {COMPILED_ROOT}")
    if modes_cache is None:
        here = os.path.dirname(os.path.abspath(__file__))
        sys.path.insert(0, os.path.dirname(here))
        from qnmfits_b200.synthetic import modes_cache as _mc
        modes_cache = _mc

    saved = {k: sys.modules.get(k) for k in
             ("matplotlib", "matplotlib.pyplot", "mpl_toolkits",
              "mpl_toolkits.axes_grid1", "h5py", "qnm", "qnmfits",
              "qnmfits.qnm", "qnmfits.qnmfits")}
    try:
        if "matplotlib" not in sys.modules:
            mpl = _stub("matplotlib")
            mpl.pyplot = _stub("matplotlib.pyplot")
        if "mpl_toolkits.axes_grid1" not in sys.modules:
            _stub("mpl_toolkits")
            _stub("mpl_toolkits.axes_grid1", make_axes_locatable=lambda ax: None)
        if "h5py" not in sys.modules:
            _stub("h5py", File=None)
        _stub("qnm", modes_cache=modes_cache)
        pkg = types.ModuleType("qnmfits")
        pkg.__path__ = [package_dir]
        sys.modules["qnmfits"] = pkg
        ref = _import_reference(package_dir)
    finally:
        # Leave no stubs behind: the reference module keeps its own references.
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _loaded = ref
    return ref
