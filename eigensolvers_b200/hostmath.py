"""Host-side small dense algebra of the eigensolver drivers: everything that acts on the m x m
overlap / Hamiltonian matrices (m <= L*nBlock) or on lists of eigenvalues.  north_star keeps
this on the host; it mirrors the behaviour of the reference's util_funcs.py (lines cited per
function) with the same names and argument meaning so that driver-level parity tests read like
the reference's.  Nothing here touches length-N data except through the vector interface.
"""
import warnings

import numpy as np
import scipy.linalg as sla
from scipy import special

from .vector_api import LINDEP_DEFAULT_VALUE


# ---------------------------------------------------------------------------- selections
def find_nearest(array, value):
    """(index, element) of the entry closest to `value` (util_funcs.py:125-128)."""
    array = np.asarray(array)
    k = int(np.argmin(np.abs(array - value)))
    return k, array[k]


def select_within_range(in_arr, arr_min, arr_max):
    """Entries with arr_min <= x <= arr_max and their indices (util_funcs.py:108-122)."""
    keep = [i for i in range(len(in_arr)) if arr_min <= in_arr[i] <= arr_max]
    return np.array([in_arr[i] for i in keep]), keep


def calculateTarget(eigenvalues, indx, tol=1e-14):
    """Shift a quarter of the smaller neighbouring gap above level `indx` (util_funcs.py:292-303)."""
    below = eigenvalues[indx] - eigenvalues[indx - 1]
    above = eigenvalues[indx + 1] - eigenvalues[indx]
    assert min(below, above) > tol, "Got a degenerate eigenvalue"
    return eigenvalues[indx] + 0.25 * min(below, above)


# ---------------------------------------------------------------------------- quadrature
def _trapezoidal(nc):
    """The reference's "trapezoidal" rule verbatim in behaviour (util_funcs.py:14-27): nodes
    a + dx*(i-1), constant weights (b-a)/(nc+1) on [-1,1]."""
    a, b = -1.0, 1.0
    dx = (b - a) / nc
    pts = np.array([a + dx * (i - 1) for i in range(nc)])
    wts = np.full(nc, (b - a) / (nc + 1))
    return pts, wts


def quadraturePointsWeights(nc, quad, positiveHalf=True):
    """Nodes/weights on [-1,1]; positiveHalf keeps nodes > 0 only (util_funcs.py:146-166)."""
    if quad == "legendre":
        gk, wk = special.roots_legendre(nc)
    elif quad == "hermite":
        gk, wk = special.roots_hermite(nc)
    elif quad == "trapezoidal":
        gk, wk = _trapezoidal(nc)
    else:
        raise UnboundLocalError(f"unknown quadrature {quad!r}")  # the reference falls through unbound
    if positiveHalf:
        sel = gk > 0.0
        gk, wk = gk[sel], wk[sel]
    return gk, wk


# ---------------------------------------------------------------------------- Rayleigh-Ritz
def lowdinOrtho(oMat, tol=LINDEP_DEFAULT_VALUE):
    """Symmetric (Loewdin) orthogonalisation keeping overlap eigenvalues > tol
    (util_funcs.py:233-247).  Returns (kept mask, all-kept flag, S^(-1/2) columns)."""
    w, u = sla.eigh(oMat)
    keep = w > tol
    return keep, bool(np.all(keep)), u[:, keep] * w[keep] ** (-0.5)


def lowdinOrthoMatrix(S, status):
    """status["lindep"] := some overlap eigenvalue <= tol; returns (status, uS) (util_funcs.py:346-358)."""
    _, independent, uS = lowdinOrtho(S)
    status["lindep"] = not independent
    return status, uS


def diagonalizeHamiltonian(X, Hmat, printObj=None):
    """eigh of X^H Hmat X (util_funcs.py:360-385)."""
    if printObj is not None:
        printObj.writeFile("hamiltonian", Hmat, "beforeOrthogonalization")
    Hort = X.T.conj() @ Hmat @ X
    ev, uv = sla.eigh(Hort)
    if printObj is not None:
        printObj.writeFile("hamiltonian", Hort, "afterOrthogonalization")
        printObj.writeFile("eigenvalues", ev)
    return ev, uv


def eigenvalueResidual(ev, reference, eigenvalueRange=None):
    """sum|reference - ev| / sum|ev|, optionally restricted to reference values inside a window
    (util_funcs.py:249-289)."""
    if eigenvalueRange is not None:
        assert len(eigenvalueRange) == 2, "Eigenvalue range needs [min, max]"
        emin, emax = eigenvalueRange
        if emin > emax:
            warnings.warn("emin is greater than emax. Moving forward with swapped values")
            emin, emax = emax, emin
        inside = select_within_range(reference, emin, emax)[1]
        if len(inside) >= 1:
            reference = reference[inside]
            ev = ev[inside]
            assert len(reference) == len(ev), "Eigenvalues are not equal in number"
    num = 0.0
    den = 0.0
    for i in range(len(ev)):
        num += abs(reference[i] - ev[i])
        den += abs(ev[i])
    return num / den


def basisTransformation(bases, coeffs):
    """New vectors sum_i coeffs[i, j] bases[i] (util_funcs.py:208-231).  1-D coeffs give one
    vector; the degenerate 1-D case [1.0] returns the input list itself inside a list, as the
    reference does.  Vector classes that offer `linearCombinationBlock` get all columns in one
    batched call (one pass over the inputs per four outputs instead of one per output)."""
    typeClass = bases[0].__class__
    coeffs = np.asarray(coeffs)
    if coeffs.ndim == 1:
        if len(coeffs) == 1 and coeffs[0] == 1.0:
            return [bases]
        return [typeClass.linearCombination(bases, coeffs)]
    block = getattr(typeClass, "linearCombinationBlock", None)
    if block is not None:
        return list(block(bases, coeffs))
    return [typeClass.linearCombination(bases, coeffs[:, j]) for j in range(coeffs.shape[1])]


# ---------------------------------------------------------------------------- pick functions
def get_pick_function_close_to_sigma(toCompare):
    """Order Ritz pairs by |eigenvalue - toCompare| (util_funcs.py:329-344)."""
    def pick(transformMat, vectors, eigenvalues):
        return np.argsort(np.abs(eigenvalues - toCompare))
    return pick


def get_pick_function_maxOvlp(toCompare):
    """Order Ritz pairs by decreasing |<Ritz vector|toCompare>| (util_funcs.py:305-327)."""
    def pick(transformMat, vectors, eigenvalues):
        m = transformMat.shape[0]
        ovl = np.zeros(m, dtype=transformMat[0].dtype)
        for i in range(m):
            ovl[i] = vectors[i].vdot(toCompare)
        weight = abs(transformMat.T.conj() @ ovl)
        return np.argsort(-weight)
    return pick
