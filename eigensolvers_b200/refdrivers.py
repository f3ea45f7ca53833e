"""The reference's own, unchanged drivers for `CudaVector` (the plug-in boundary is the
`AbstractVector` ABC; `inexact_Lanczos.py` / `feast.py` stay as they are and dispatch on
`typeClass = type(v0[0])`, inexact_Lanczos.py:284, feast.py:168).

`load()` imports `inexactLanczosDiagonalization` (inexact_Lanczos.py:229-443) and
`feastDiagonalization` (feast.py:126-244) from a reference checkout or installation:
    $EIGENSOLVERS_REFERENCE,  else  <repo>/baseline/_ref  (baseline/install_reference.py)
and registers `CudaVector` with the reference's ABC so `issubclass(type(v0), AbstractVector)`
(inexact_Lanczos.py:278) holds whichever module was imported first.  When no reference is present
`available()` is False and callers use the stand-alone mirrors `eigensolvers_b200.lanczos` /
`.contour` (same signatures, reproduce the reference's trajectories bit for bit on the CPU oracle,
tests/test_oracle.py).

The reference imports three in-house modules it does not ship (`util`, `magic`, `ttns2`); a
maintainer has them, the installation under baseline/_ref carries stubs in `_stubs/`.
"""
import importlib
import os
import sys
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CACHE = [None]


def reference_dir():
    for cand in (os.environ.get("EIGENSOLVERS_REFERENCE"), os.path.join(_ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "inexact_Lanczos.py")):
            return cand
    return None


def available():
    return reference_dir() is not None


def _adapt_scipy_tol():
    """The reference passes `tol=` to scipy.sparse.linalg.gcrotmk/minres (numpyVector.py:161,163);
    SciPy >= 1.14 names that argument `rtol` (same criterion).  Only NumpyVector.solve needs this —
    CudaVector never calls SciPy."""
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


def load(register_cuda=True, numpy_backend=False):
    """Namespace with the reference's modules and its two driver functions.
    register_cuda=False: do not import CudaVector (bench.py's CPU arm must not touch the GPU code);
    numpy_backend=True: also expose the reference's `numpyVector` module, with the SciPy keyword adapter."""
    if _CACHE[0] is not None:
        ns = _CACHE[0]
        if numpy_backend and not hasattr(ns, "numpyVector"):
            _adapt_scipy_tol()
            ns.numpyVector = importlib.import_module("numpyVector")
        if register_cuda:
            _register(ns)
        return ns
    ref = reference_dir()
    if ref is None:
        raise RuntimeError("no reference checkout: set EIGENSOLVERS_REFERENCE or run baseline/install_reference.py")
    sys.dont_write_bytecode = True
    stubs = os.path.join(ref, "_stubs")
    for p in (stubs, ref):
        if os.path.isdir(p) and p not in sys.path:
            sys.path.insert(0, p)
    ns = types.SimpleNamespace(path=ref)
    for name in ("abstractVector", "util_funcs", "inexact_Lanczos", "feast"):
        setattr(ns, name, importlib.import_module(name))
    if numpy_backend:
        _adapt_scipy_tol()
        ns.numpyVector = importlib.import_module("numpyVector")
    ns.inexactLanczosDiagonalization = ns.inexact_Lanczos.inexactLanczosDiagonalization
    ns.feastDiagonalization = ns.feast.feastDiagonalization
    _CACHE[0] = ns
    if register_cuda:
        _register(ns)
    return ns


def _register(ns):
    from .cudaVector import CudaVector
    if not issubclass(CudaVector, ns.abstractVector.AbstractVector):
        ns.abstractVector.AbstractVector.register(CudaVector)


def lanczos_driver():
    """(function, label): the reference's driver when present, else the stand-alone mirror."""
    if available():
        return load().inexactLanczosDiagonalization, "reference inexact_Lanczos.py (unchanged)"
    from .lanczos import inexactLanczosDiagonalization
    return inexactLanczosDiagonalization, "eigensolvers_b200.lanczos (mirror)"


def feast_driver():
    if available():
        return load().feastDiagonalization, "reference feast.py (unchanged)"
    from .contour import feastDiagonalization
    return feastDiagonalization, "eigensolvers_b200.contour (mirror)"
