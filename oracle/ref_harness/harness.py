"""Make the UNMODIFIED reference importable in this container (TEST INFRASTRUCTURE).

`load()` puts /root/reference and the stub modules (shims/) on sys.path, adapts SciPy's renamed
keyword (the reference passes `tol=` to gcrotmk/minres, numpyVector.py:161,163; SciPy >= 1.14
calls it `rtol` — same criterion) and returns the reference's modules.  Drivers must be called
with saveTNSsEachIteration=False (the default True needs a `.ttns` attribute) and
from a scratch working directory (they write *.out files into the CWD).

/root/reference exists only in the build container; nothing that runs on the GPU box uses this.
"""
import importlib
import os
import sys
import types

REFERENCE = os.environ.get("EIGENSOLVERS_REFERENCE", "/root/reference")
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available():
    return os.path.isfile(os.path.join(REFERENCE, "numpyVector.py"))


def _patch_scipy_tol():
    import scipy.sparse.linalg as spla
    if getattr(spla, "_tol_adapter_installed", False):
        return
    for name in ("gcrotmk", "minres"):
        orig = getattr(spla, name)

        def adapted(*args, _orig=orig, **kw):
            if "tol" in kw:
                kw["rtol"] = kw.pop("tol")
            return _orig(*args, **kw)
        setattr(spla, name, adapted)
    spla._tol_adapter_installed = True


def load():
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    for p in (SHIMS, REFERENCE):
        if p not in sys.path:
            sys.path.insert(0, p)
    _patch_scipy_tol()
    mods = types.SimpleNamespace()
    for name in ("abstractVector", "numpyVector", "util_funcs", "inexact_Lanczos", "feast"):
        setattr(mods, name, importlib.import_module(name))
    return mods
