"""The vector plug-in interface the eigensolver drivers are written against.

The reference states it as the ``AbstractVector`` ABC (abstractVector.py:15-169): three
properties, the scalar/BLAS-1 instance methods and eight static methods on lists of vectors.
Here the interface is DATA -- ``INTERFACE`` below, one row per member with its kind, its argument
names (the reference's, so keyword calls made by the unchanged drivers keep working) and the
reference lines it restates -- and the stand-alone base class is generated from that table.
``conformance(cls)`` checks any implementation (``CudaVector``, the oracle's ``NumpyVector``)
against the table; ``tests/test_abi.py`` also checks the table against the reference's own ABC
when ``baseline/_ref`` is installed.

When the reference package itself is importable (a maintainer running the unchanged
``inexact_Lanczos.py`` / ``feast.py`` with ``PYTHONPATH`` pointing at it), ``CudaVector``
subclasses *that* class so the driver's ``issubclass(type(v0), AbstractVector)`` check
(inexact_Lanczos.py:278) holds; otherwise it subclasses the generated mirror.
"""
import inspect
from abc import ABC, abstractmethod

# abstractVector.py:12
LINDEP_DEFAULT_VALUE = 1e-14

_LD = ("lindep", LINDEP_DEFAULT_VALUE)

# (name, kind, arguments after self / of the static method, reference lines, meaning)
# an argument is a name or a (name, default) pair
INTERFACE = (
    ("hasExactAddition", "property", (), "abstractVector.py:17-26",
     "True when c + c* is exactly 2 Re(c) for this vector type (FEAST uses it, feast.py:89)"),
    ("dtype", "property", (), "abstractVector.py:28-31", "numpy dtype of the elements"),
    ("maxD", "property", (), "abstractVector.py:33-37",
     "largest bond dimension; 0 for plain vectors (inexact_Lanczos.py:310)"),
    ("__mul__", "method", ("other",), "abstractVector.py:39-41", "vector * scalar"),
    ("__rmul__", "method", ("other",), "abstractVector.py:43-45", "scalar * vector"),
    ("__truediv__", "method", ("other",), "abstractVector.py:47-49", "vector / scalar"),
    ("__imul__", "method", ("other",), "abstractVector.py:51-53", "in-place scaling"),
    ("__itruediv__", "method", ("other",), "abstractVector.py:55-57", "in-place division"),
    ("__len__", "method", (), "abstractVector.py:59-61", "number of elements"),
    ("normalize", "method", (), "abstractVector.py:63-66", "normalise in place and return self"),
    ("norm", "method", (), "abstractVector.py:68-70", "Euclidean norm"),
    ("real", "method", (), "abstractVector.py:72-74", "real part as a new vector"),
    ("conjugate", "method", (), "abstractVector.py:76-78", "complex conjugate as a new vector"),
    ("vdot", "method", ("other", ("conjugate", True)), "abstractVector.py:80-82",
     "<self|other>, or the unconjugated product when conjugate is False"),
    ("copy", "method", (), "abstractVector.py:84-86", "deep copy"),
    ("applyOp", "method", ("other",), "abstractVector.py:88-91", "other @ self as a new vector"),
    ("compress", "method", (), "abstractVector.py:93-97", "compress if compressible; may return self"),
    ("linearCombination", "static", ("other", "coeff"), "abstractVector.py:99-109",
     "sum_i coeff[i] * other[i]"),
    ("orthogonalize", "static", ("xs", _LD), "abstractVector.py:111-113",
     "orthonormalise a list; drops linearly dependent members"),
    ("orthogonalize_against_set", "static", ("x", "xs", _LD), "abstractVector.py:115-125",
     "x made orthonormal to xs; None when it is linearly dependent"),
    ("solve", "static", ("H", "b", "sigma", ("x0", None), ("opType", "her"), ("reverseGF", False)),
     "abstractVector.py:127-139", "(sigma - H) x = b, or (H - sigma) x = b with reverseGF"),
    ("matrixRepresentation", "static", ("operator", "vectors"), "abstractVector.py:141-144",
     "<v_i|operator|v_j>"),
    ("overlapMatrix", "static", ("vectors",), "abstractVector.py:146-149", "<v_i|v_j>"),
    ("extendMatrixRepresentation", "static", ("operator", "vectors", "opMat"), "abstractVector.py:151-159",
     "grow opMat by the rows/columns of the vectors appended since it was built"),
    ("extendOverlapMatrix", "static", ("vectors", "overlap"), "abstractVector.py:161-169",
     "grow the overlap matrix likewise"),
)


def _signature(args, bound):
    params = [inspect.Parameter("self", inspect.Parameter.POSITIONAL_OR_KEYWORD)] if bound else []
    for a in args:
        name, default = (a, inspect.Parameter.empty) if isinstance(a, str) else a
        params.append(inspect.Parameter(name, inspect.Parameter.POSITIONAL_OR_KEYWORD, default=default))
    return inspect.Signature(params)


def _stub(name, args, bound, doc):
    def member(*a, **k):
        raise NotImplementedError(name)
    member.__name__ = member.__qualname__ = name
    member.__doc__ = doc
    member.__signature__ = _signature(args, bound)
    return member


def _build_interface():
    body = {"__doc__": "Stand-alone statement of the plug-in interface, generated from INTERFACE."}
    for name, kind, args, where, meaning in INTERFACE:
        doc = f"{meaning} ({where})"
        if kind == "property":
            body[name] = property(abstractmethod(_stub(name, (), True, doc)), doc=doc)
        elif kind == "method":
            body[name] = abstractmethod(_stub(name, args, True, doc))
        else:   # the reference leaves the static members concrete: they raise until overridden
            body[name] = staticmethod(_stub(name, args, False, doc))
    return type("_VectorInterface", (ABC,), body)


_VectorInterface = _build_interface()


def conformance(cls, strict_static=True):
    """Differences between ``cls`` and INTERFACE as a list of strings (empty = conforms): a missing
    member, a member of the wrong kind, a different number of required arguments, or optional
    arguments whose names/defaults differ from the reference's (the drivers pass those by keyword:
    ``conjugate=``, ``lindep=``, ``opType=``, ``reverseGF=``).  The names of the required arguments
    are not compared -- the reference's own NumpyVector (numpyVector.py:105,121) does not keep the
    ABC's.  ``strict_static=False`` accepts undecorated functions for the static members, which is
    how numpyVector.py:105-238 declares them (they then work on the class only, not on instances)."""
    problems = []
    for name, kind, args, where, _ in INTERFACE:
        raw = inspect.getattr_static(cls, name, None)
        if raw is None:
            problems.append(f"{name}: missing ({where})")
            continue
        if kind == "property":
            if not isinstance(raw, property):
                problems.append(f"{name}: must be a property ({where})")
            continue
        if kind == "static" and not strict_static and inspect.isfunction(raw):
            pass
        elif (kind == "static") != isinstance(raw, staticmethod):
            problems.append(f"{name}: must be {'a static' if kind == 'static' else 'an instance'} method ({where})")
            continue
        fn = raw.__func__ if isinstance(raw, staticmethod) else raw
        if not inspect.isfunction(fn):
            problems.append(f"{name}: must be a function, found {type(raw).__name__} ({where})")
            continue
        want = _signature(args, kind == "method")
        have = inspect.signature(fn)
        hp = [p for p in have.parameters.values() if p.kind is not inspect.Parameter.VAR_KEYWORD]
        wp = list(want.parameters.values())
        # an implementation may append extra optional arguments, never rename or reorder the reference's
        if len(hp) < len(wp):
            problems.append(f"{name}{have}: fewer arguments than {want} ({where})")
        elif any(h.default != w.default for h, w in zip(hp, wp)):
            problems.append(f"{name}{have}: defaults differ from {want} ({where})")
        elif any(h.name != w.name for h, w in zip(hp, wp) if w.default is not inspect.Parameter.empty):
            problems.append(f"{name}{have}: keyword arguments differ from {want} ({where})")
        elif any(p.default is inspect.Parameter.empty for p in hp[len(wp):]):
            problems.append(f"{name}{have}: extra arguments must be optional ({where})")
    return problems


def _resolve_base():
    """Prefer the reference's own ABC when it is on the path (drop-in use), else the mirror."""
    try:
        from abstractVector import AbstractVector as RefAbstractVector  # reference module
        if all(hasattr(RefAbstractVector, row[0]) for row in INTERFACE if row[1] == "static"):
            return RefAbstractVector
    except Exception:
        pass
    return _VectorInterface


AbstractVector = _resolve_base()
