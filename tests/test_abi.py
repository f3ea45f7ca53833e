"""The C-ABI library loads on a CPU-only machine and exports exactly what include/cudavec.h
declares; the ctypes table (eigensolvers_b200/_lib.py) lists the same symbols.  No compute calls."""
import ctypes
import os
import re

import numpy as np

from eigensolvers_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "cudavec.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cv_[a-z0-9_]+)\s*\(", text)))


def test_header_and_ctypes_table_agree():
    assert header_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.cv_abi_version() == 2
    assert lib.cv_ctx_scratch_bytes() > 0
    assert isinstance(lib.cv_last_error(), bytes)


def test_host_side_partition_routine_runs_without_gpu():
    lib = _lib.load()
    off = np.empty(4, dtype=np.int64)
    _lib.check(lib.cv_partition_rows(10, 3, off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))))
    assert list(off) == [0, 3, 6, 10]
    assert lib.cv_partition_rows(-1, 3, off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))) != 0
    assert b"cv_partition_rows" in lib.cv_last_error()


def test_product_has_no_cpu_fallback():
    """The product package never imports the oracle, and constructing a vector without a GPU
    fails loudly instead of computing on the CPU."""
    import torch
    pkg = os.path.join(ROOT, "eigensolvers_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
    if not torch.cuda.is_available():
        import pytest
        from eigensolvers_b200 import CudaVector
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            CudaVector(np.ones(3))


def test_vector_classes_conform_to_the_plugin_interface():
    """CudaVector and the oracle's NumpyVector expose every member of the reference's ABC
    (abstractVector.py:15-169) with the reference's argument names and defaults."""
    from eigensolvers_b200 import CudaVector
    from eigensolvers_b200.vector_api import INTERFACE, conformance, _VectorInterface
    from oracle.numpy_vector import NumpyVectorOracle as NumpyVector
    assert conformance(CudaVector) == []
    assert conformance(NumpyVector, strict_static=False) == []
    assert conformance(NumpyVector) != []     # declared as in numpyVector.py: plain functions, class-level use only
    assert conformance(_VectorInterface) == []
    assert len(INTERFACE) == 25

    class Broken(CudaVector):
        @staticmethod
        def solve(H, rhs, sigma):   # renamed argument, missing optionals
            return None
        norm = property(lambda self: 0.0)
    bad = conformance(Broken)
    assert any(p.startswith("solve") for p in bad) and any(p.startswith("norm") for p in bad)


def test_interface_table_matches_the_reference_abc():
    """The table is checked against the reference's own abstractVector.py when it is installed
    (baseline/_ref, git-ignored): same members, same kinds, same signatures."""
    import inspect
    import os
    import sys
    import pytest
    from eigensolvers_b200.vector_api import INTERFACE, conformance
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    if not os.path.exists(os.path.join(ref, "abstractVector.py")):
        pytest.skip("reference not installed under baseline/_ref")
    sys.path.insert(0, ref)
    try:
        import abstractVector as ref_mod
    finally:
        sys.path.remove(ref)
    assert conformance(ref_mod.AbstractVector) == []
    public = {n for n, v in vars(ref_mod.AbstractVector).items()
              if not n.startswith("_") or n in ("__mul__", "__rmul__", "__truediv__", "__imul__", "__itruediv__", "__len__")}
    public -= {"_abc_impl"}
    assert public == {row[0] for row in INTERFACE}
    assert ref_mod.LINDEP_DEFAULT_VALUE == 1e-14
    for name, kind, *_ in INTERFACE:
        raw = inspect.getattr_static(ref_mod.AbstractVector, name)
        abstract = getattr(raw.fget if isinstance(raw, property) else raw, "__isabstractmethod__", False)
        assert abstract == (kind != "static"), name
