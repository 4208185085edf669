"""``pythtb`` — backend switch for existing PythTB scripts.

Put this directory's parent (``<repo>/shim``) and the repository root on ``PYTHONPATH`` ahead of an
installed PythTB; an unmodified script that does ``from pythtb import *`` (or ``tb_model, wf_array,
w90``) then runs on whichever implementation the flag names::

    PYTHTB_BACKEND=b200       the B200 engine (``pythtb_b200``: CUDA kernels behind the C ABI of include/tbk.h)
    PYTHTB_BACKEND=reference  the stock PythTB found further down ``sys.path`` (or under ``PYTHTB_REFERENCE``)

The default is ``b200``.  ``pythtb.set_backend(name)`` switches at run time (names imported with
``from pythtb import ...`` before the switch stay bound to the old classes), ``pythtb.get_backend()``
tells which one is active.  There is no silent fallback in either direction: a missing CUDA library /
device raises from the first numerical call of the b200 backend, a missing stock PythTB raises
``ImportError`` when the reference backend is asked for.
"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_PUBLIC = ("tb_model", "wf_array", "w90")
_backend = None
_impl = None


def _load_b200():
    root = os.path.dirname(os.path.dirname(_HERE))
    try:
        import pythtb_b200
    except ImportError:
        if root not in sys.path:
            sys.path.append(root)
        import pythtb_b200
    return pythtb_b200


def _load_reference():
    """The stock module: ``pythtb.py`` / ``pythtb/__init__.py`` in ``$PYTHTB_REFERENCE`` or in the first
    ``sys.path`` entry that is not this shim."""
    cands = []
    if os.environ.get("PYTHTB_REFERENCE"):
        cands.append(os.environ["PYTHTB_REFERENCE"])
    cands += [p or os.getcwd() for p in sys.path]
    for d in cands:
        for rel in ("pythtb.py", os.path.join("pythtb", "__init__.py")):
            path = os.path.join(d, rel)
            if os.path.isfile(path) and os.path.dirname(os.path.abspath(path)) != _HERE:
                spec = importlib.util.spec_from_file_location("_pythtb_reference", path)
                mod = importlib.util.module_from_spec(spec)
                sys.modules["_pythtb_reference"] = mod
                spec.loader.exec_module(mod)
                return mod
    raise ImportError("PYTHTB_BACKEND=reference: no stock PythTB (pythtb.py) on sys.path or under $PYTHTB_REFERENCE")


def set_backend(name):
    """Select the implementation behind ``pythtb.tb_model`` / ``wf_array`` / ``w90``: 'b200' or 'reference'."""
    global _backend, _impl
    name = str(name).lower()
    if name in ("b200", "gpu", "cuda"):
        impl, name = _load_b200(), "b200"
    elif name in ("reference", "numpy", "cpu"):
        impl, name = _load_reference(), "reference"
    else:
        raise ValueError("unknown PythTB backend %r (expected 'b200' or 'reference')" % (name,))
    g = globals()
    for key in _PUBLIC:
        g[key] = getattr(impl, key)
    g["__version__"] = getattr(impl, "__version__", "unknown")
    _backend, _impl = name, impl
    return name


def get_backend():
    return _backend


def __getattr__(attr):
    # anything else a script reaches for (helper functions of the stock module) comes from the active backend
    if _impl is not None and not attr.startswith("__") and hasattr(_impl, attr):
        return getattr(_impl, attr)
    raise AttributeError("module 'pythtb' (backend %s) has no attribute %r" % (_backend, attr))


set_backend(os.environ.get("PYTHTB_BACKEND", "b200"))
__all__ = list(_PUBLIC) + ["set_backend", "get_backend"]
