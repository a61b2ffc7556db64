"""Loads the UNMODIFIED reference coder compiled into oracle/_ref (see oracle/Makefile).

TEST INFRASTRUCTURE: used by tests/ and bench.py's cpu_baseline / --impl reference legs only.
oracle/_ref is git-ignored (built artefact) but travels to the GPU box with the snapshot.
"""
import glob
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def load(name="ans"):
    """Return the reference pybind11 module `name` ("ans" or "rans"), or None if it was never built."""
    if name in _cache:
        return _cache[name]
    import sys
    if "cbench." + name in sys.modules:      # already imported through the reference package (ref_shim)
        _cache[name] = sys.modules["cbench." + name]
        return _cache[name]
    mod = None
    hits = glob.glob(os.path.join(_HERE, "_ref", name + ".*.so"))
    if hits:
        try:
            spec = importlib.util.spec_from_file_location(name, hits[0])
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        except Exception:  # wrong python ABI on this box etc.
            mod = None
    _cache[name] = mod
    return mod
