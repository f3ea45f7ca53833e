"""Stand-alone restarted (block) inexact shift-and-invert Lanczos driver for the vector plug-in.

The product path runs the REFERENCE'S OWN `inexactLanczosDiagonalization` unchanged on
`CudaVector` (eigensolvers_b200/refdrivers.py, tests/test_gpu_dropin.py, bench.py).  This module
exists for two things the reference's file cannot give:

  * a driver on machines without a reference checkout — same call signature, status dictionary,
    return values and failure behaviour as inexact_Lanczos.py:229-443 (tests/test_oracle.py: driven
    with a NumPy vector class it reproduces the reference's runs bit for bit, same Krylov steps);
  * LOCK-STEP block solves: the reference normalises each solve's result before it requests the
    next (inexact_Lanczos.py:319-327), which serialises the nBlock independent solves of a Krylov
    step; here they are requested together (`solveBlock`) when the vector class offers it.

Structure: a `_KrylovSpace` object owns the list of vectors with its overlap / Hamiltonian matrices
and knows how to grow, diagonalise (Rayleigh-Ritz in the Loewdin basis) and collapse itself; the
driver function is the schedule of outer (restart) and inner (growth) steps around it.

Deviations from the reference, both of which only remove crashes:
  * `saveTNSsEachIteration` defaults to False (the reference's default True raises AttributeError
    for non-TTNS vectors); when True, vectors exposing `.ttns.saveToHDF5` are saved.
  * a linear dependency or zero vector on the very first Krylov step returns NaN eigenvalues
    instead of raising UnboundLocalError.
"""
import os
import time
import warnings

import numpy as np
import scipy.linalg as sla

from .hostmath import (basisTransformation, diagonalizeHamiltonian, eigenvalueResidual,
                       get_pick_function_close_to_sigma, lowdinOrthoMatrix)
from .runlog import LanczosRunLog
from .vector_api import AbstractVector

_STATUS_DEFAULTS = dict(ref=None, residual=np.inf, nBlock=None, flagAddition=None, outerIter=0, innerIter=0,
                        cumIter=0, iBlock=0, zeroVector=False, isConverged=False, lindep=False,
                        futileRestarts=0, startTime=None, runTime=0.0, KSmaxD=None, fitmaxD=None, phase=1)


def _initial_status(user, first_vector, nBlock):
    """Status dictionary: defaults, then every key of the user's dict on top (no whitelist — the
    reference's tests pass unrelated keys through it; inexact_Lanczos.py:23-82)."""
    st = dict(_STATUS_DEFAULTS)
    st.update(ref=[], KSmaxD=[], nBlock=nBlock, flagAddition=first_vector.hasExactAddition, startTime=time.time())
    st.update(user or {})
    return st


class _KrylovSpace:
    """The Krylov list Y with S = <Y|Y> and Hm = <Y|H|Y>, kept current column by column."""

    def __init__(self, vecClass, H, vectors):
        self.vc, self.H = vecClass, H
        self.Y = list(vectors)
        self.S = vecClass.overlapMatrix(self.Y)
        self.Hm = None          # filled by measure(); the initial-guess check comes first
        self.ritz = None        # (values, coefficients) of the last Rayleigh-Ritz, ordered by `pick`

    def measure(self):
        self.Hm = self.vc.matrixRepresentation(self.H, self.Y)

    def is_orthonormal(self, tol):
        return np.allclose(self.S, np.eye(len(self.Y)), rtol=tol, atol=tol)

    def absorb(self, candidate):
        """Gram-Schmidt `candidate` against Y (reference semantics, numpyVector.py:121-145); append it
        and one column to S and Hm.  False when it is linearly dependent (GS returned None)."""
        q = self.vc.orthogonalize_against_set(candidate, self.Y)
        if q is None:
            return False
        self.Y.append(q.compress())
        both = getattr(self.vc, "extendBoth", None)
        if both is not None:      # one SpMV + one pass over the list for the two columns
            self.S, self.Hm = both(self.H, self.Y, self.S, self.Hm)
        else:
            self.S = self.vc.extendOverlapMatrix(self.Y, self.S)
            self.Hm = self.vc.extendMatrixRepresentation(self.H, self.Y, self.Hm)
        return True

    def rayleigh_ritz(self, pick, status, log):
        """Ritz pairs in the Loewdin-orthogonalised basis (util_funcs.py:346-385), ordered by `pick`."""
        status, lowdin = lowdinOrthoMatrix(self.S, status)
        assert not status["lindep"]          # absorb() has already refused dependent vectors
        values, rot = diagonalizeHamiltonian(lowdin, self.Hm, log)
        coeff = lowdin @ rot
        order = pick(coeff, self.Y, values)
        assert len(order) == len(values), f"{len(values)=} {len(order)=}"
        self.ritz = (values[order], coeff[:, order])
        return status

    def collapse_to_ritz_vectors(self, count=None):
        """Replace Y by Ritz vectors: all of them (final answer) or the first `count`, each normalised
        on its own and NOT re-orthogonalised (restart, inexact_Lanczos.py:414-425)."""
        coeff = self.ritz[1]
        if count is None:
            self.Y = basisTransformation(self.Y, coeff)
        else:
            self.Y = [self.vc.normalize(basisTransformation(self.Y, coeff[:, b])[0]) for b in range(count)]
        self.S = self.vc.overlapMatrix(self.Y)


def _shifted_solves(vc, Hsolve, sources, sigma, eConv, lockstep):
    """Normalised results of (sigma - H) w = y for the given sources, in order; the list stops in front of
    the first (numerically) zero result, whose norm is returned as second value
    (inexact_Lanczos.py:84-105, 319-327: norm <= 0.001*eConv counts as zero)."""
    together = getattr(vc, "solveBlock", None) if (lockstep and len(sources) > 1) else None
    raw = together(Hsolve, sources, sigma) if together is not None else None
    done = []
    for i, y in enumerate(sources):
        w = raw[i] if raw is not None else vc.solve(Hsolve, y, sigma)
        size = vc.norm(w)
        if not size > 0.001 * eConv:
            return done, size
        done.append(vc.normalize(w))
    return done, None


def _update_convergence(values, eConv, status, log):
    """Relative change of the sorted first nBlock picked values against the previous step's, from the
    second cumulative step on (inexact_Lanczos.py:115-143)."""
    now = np.sort(values[:status["nBlock"]])
    hit = False
    if status["cumIter"] > 1:
        status["residual"] = eigenvalueResidual(now, status["ref"][-1])
        hit = status["residual"] <= eConv
    status["isConverged"] = bool(hit)
    status["runTime"] = time.time() - status["startTime"]
    if log is not None:
        log.writeFile("summary", now, status)
    status["ref"] = (status["ref"] + [now])[-2:]
    return status


def _out_of_budget(status, maxit, L):
    last = status["outerIter"] == maxit - 1 and status["innerIter"] == L - 1
    if last and not status["isConverged"]:
        print("Alert: Lanczos iterations is not converged!")
    return last


def _futile(block_values, eConv, status, limit=3):
    """Restarts that do not move the block energies while lindep is flagged are counted; more than
    `limit` of them end the run (inexact_Lanczos.py:167-194)."""
    if status["lindep"] and eigenvalueResidual(block_values, status["ref"][0]) > max(1e-9, eConv):
        status["futileRestarts"] += 1
    if status["futileRestarts"] > limit:
        warnings.warn("Lindep and did not have fruitful restarts")
        return True
    return False


def _checkpoint(space, status, saveDir):
    os.makedirs(saveDir, exist_ok=True)
    extra = {"status": status, "eigencoefficients": space.ritz[1], "eigenvalues": space.ritz[0]}
    for i, vec in enumerate(space.Y):
        vec.ttns.saveToHDF5(f"{saveDir}/tns_{status['cumIter']}_{i}.h5", additionalInformation=extra)


def inexactLanczosDiagonalization(H, v0, sigma, L, maxit, eConv, checkFitTol=1e-7,
                                  Hsolve=None, pick=None, status=None,
                                  writeOut=True, eShift=0.0, convertUnit="au",
                                  outFileName=None, summaryFileName=None,
                                  saveTNSsEachIteration=False, saveDir="saveTNSs", lockstep=True):
    """Eigenpairs of H closest to `sigma` (or selected by `pick`).

    H        operator for the Rayleigh-Ritz matrices; Hsolve (default H) is used for the solves
    v0       guess vector or list of mutually orthogonal guess vectors (block Lanczos)
    L        Krylov vectors per block between restarts;  maxit  outer iterations
    eConv    relative eigenvalue-change tolerance
    lockstep request the nBlock solves of a step together when the vector class has `solveBlock`
    Returns (ev, Ylist, status): ALL Ritz values ordered by `pick`, the Ritz vectors, the status.
    """
    if issubclass(type(v0), AbstractVector):
        v0 = [v0]
    else:
        assert isinstance(v0, (list, tuple, np.ndarray)), f"{v0=} {type(v0)=}"
    Hsolve = H if Hsolve is None else Hsolve
    vc, nBlock = type(v0[0]), len(v0)

    space = _KrylovSpace(vc, H, v0)
    if not space.is_orthonormal(1e-3):
        if nBlock > 1:
            raise RuntimeError(f"Input vectors not orthogonalized: Smat={space.S}")
        space.Y[0].normalize()   # a single guess is normalised quietly (inexact_Lanczos.py:292-295)
        space.S[0, 0] = 1
    space.measure()

    status = _initial_status(status, space.Y[0], nBlock)
    pick = pick if pick is not None else get_pick_function_close_to_sigma(sigma)
    assert callable(pick)
    log = LanczosRunLog(space.Y[0], sigma, L, maxit, eConv, checkFitTol, writeOut, eShift,
                        convertUnit, pick, status, outFileName, summaryFileName)
    log.fileHeader()

    ev = np.full(len(space.Y), np.nan)
    for outer in range(maxit):
        status.update(outerIter=outer, KSmaxD=[space.Y[0].maxD], fitmaxD=None)
        space.ritz = None
        verdict = "restart"                      # how this outer iteration ends
        for inner in range(1, L):
            status["innerIter"] = inner
            status["cumIter"] += 1
            # new directions from the last nBlock vectors, visited back to front
            fresh, zero_norm = _shifted_solves(vc, Hsolve, [space.Y[-b] for b in range(1, nBlock + 1)], sigma, eConv,
                                               lockstep)
            if zero_norm is not None:
                status["zeroVector"] = True
                warnings.warn(f"Alert: zero vector: ||inv(H-sigma)vec||={zero_norm:5.3e}")
                # the reference leaves the Krylov loop and restarts from the last Ritz vectors; with none
                # yet (first step) there is nothing to restart from: stop
                verdict = "restart" if space.ritz is not None else "abort"
                break
            # orthogonalise in list order, append, grow S and Hm
            dependent = False
            for b, w in enumerate(fresh):
                status["iBlock"] = b
                if not space.absorb(w):
                    dependent = True
                    if log.writeOut:
                        warnings.warn(f"Linear dependency problem in iteration {outer} and microiteration {inner} "
                                      f"for block state {b}, abort current Lanczos iteration and restart.")
                    break
                status["KSmaxD"].append(space.Y[-1].maxD)
            log.writeFile("iteration", status)
            log.writeFile("overlap", space.S)
            log.writeFile("KSmaxD", status)
            if dependent:                        # inexact_Lanczos.py:356-359: NaN results, the solver returns
                ev = np.full(len(space.Y), np.nan)
                verdict = "abort"
                break
            status = space.rayleigh_ritz(pick, status, log)
            ev = space.ritz[0]
            status = _update_convergence(ev, eConv, status, log)
            if saveTNSsEachIteration:
                _checkpoint(space, status, saveDir)
            if status["isConverged"] or _out_of_budget(status, maxit, L):
                verdict = "finish"
                break
        if verdict == "abort":
            break
        if verdict == "finish":                  # back-transform ALL Ritz vectors, check orthonormality
            space.collapse_to_ritz_vectors()
            if not space.is_orthonormal(checkFitTol):
                warnings.warn(f"Alert:Final eigenvectors are not properly fitted. S=\n{space.S}")
            status["fitmaxD"] = [v.maxD for v in space.Y]
            log.writeFile("fitmaxD", status)
            break
        # plain restart from the nBlock picked Ritz vectors
        space.collapse_to_ritz_vectors(nBlock)
        space.measure()
        if not space.is_orthonormal(checkFitTol):
            warnings.warn(f"Alert:Final eigenvectors are not properly fitted. S=\n{space.S}")
            break
        if _futile(sla.eigvalsh(space.Hm, space.S), eConv, status):
            break
        status["fitmaxD"] = [v.maxD for v in space.Y]
        log.writeFile("fitmaxD", status)

    log.writeFile("results", ev)
    log.fileFooter()
    return ev, space.Y, status
