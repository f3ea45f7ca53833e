"""The vector plug-in interface the eigensolver drivers are written against.

This mirrors the reference's ``AbstractVector`` ABC (abstractVector.py:15-169): three
properties, the scalar/BLAS-1 instance methods and eight static methods.  When the reference
package itself is importable (a maintainer running the unchanged ``inexact_Lanczos.py`` /
``feast.py`` with ``PYTHONPATH`` pointing at it), ``CudaVector`` subclasses *that* class so
the driver's ``issubclass(type(v0), AbstractVector)`` check (inexact_Lanczos.py:278) holds;
otherwise it subclasses the stand-alone mirror defined here.
"""
from abc import ABC, abstractmethod

# abstractVector.py:12
LINDEP_DEFAULT_VALUE = 1e-14


class _VectorInterface(ABC):
    """Stand-alone statement of the interface (names and argument meaning as in the reference)."""

    # -- properties (abstractVector.py:17-37) ------------------------------------------------
    @property
    @abstractmethod
    def hasExactAddition(self):
        """True when c + c* is exactly 2 Re(c) for this vector type (used by FEAST, feast.py:89)."""

    @property
    @abstractmethod
    def dtype(self):
        """numpy dtype of the elements."""

    @property
    @abstractmethod
    def maxD(self):
        """Largest bond dimension; 0 for plain vectors (inexact_Lanczos.py:310)."""

    # -- scalar algebra (abstractVector.py:39-57) ----------------------------------------------
    @abstractmethod
    def __mul__(self, other): ...

    @abstractmethod
    def __rmul__(self, other): ...

    @abstractmethod
    def __truediv__(self, other): ...

    @abstractmethod
    def __imul__(self, other): ...

    @abstractmethod
    def __itruediv__(self, other): ...

    @abstractmethod
    def __len__(self): ...

    # -- BLAS-1 (abstractVector.py:59-99) -----------------------------------------------------
    @abstractmethod
    def normalize(self):
        """Normalise in place and return self."""

    @abstractmethod
    def norm(self): ...

    @abstractmethod
    def real(self): ...

    @abstractmethod
    def conjugate(self): ...

    @abstractmethod
    def vdot(self, other, conjugate=True): ...

    @abstractmethod
    def copy(self): ...

    @abstractmethod
    def applyOp(self, other):
        """Return ``other @ self`` as a new vector."""

    @abstractmethod
    def compress(self):
        """Compress if compressible; may return self."""

    # -- static algebra on lists of vectors (abstractVector.py:101-169) -------------------------
    @staticmethod
    def linearCombination(other, coeff):
        raise NotImplementedError

    @staticmethod
    def orthogonalize(xs, lindep=LINDEP_DEFAULT_VALUE):
        raise NotImplementedError

    @staticmethod
    def orthogonalize_against_set(x, xs, lindep=LINDEP_DEFAULT_VALUE):
        raise NotImplementedError

    @staticmethod
    def solve(H, b, sigma, x0=None, opType="her", reverseGF=False):
        raise NotImplementedError

    @staticmethod
    def matrixRepresentation(operator, vectors):
        raise NotImplementedError

    @staticmethod
    def overlapMatrix(vectors):
        raise NotImplementedError

    @staticmethod
    def extendMatrixRepresentation(operator, vectors, opMat):
        raise NotImplementedError

    @staticmethod
    def extendOverlapMatrix(vectors, overlap):
        raise NotImplementedError


def _resolve_base():
    """Prefer the reference's own ABC when it is on the path (drop-in use), else the mirror."""
    try:
        from abstractVector import AbstractVector as RefAbstractVector  # reference module
        required = ("linearCombination", "orthogonalize_against_set", "solve", "overlapMatrix")
        if all(hasattr(RefAbstractVector, name) for name in required):
            return RefAbstractVector
    except Exception:
        pass
    return _VectorInterface


AbstractVector = _resolve_base()
