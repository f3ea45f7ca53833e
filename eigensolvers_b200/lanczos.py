"""Restarted (block) inexact shift-and-invert Lanczos — the host control flow that drives the
vector plug-in.  Same call signature, status dictionary, return values and failure behaviour as
the reference's ``inexactLanczosDiagonalization`` (inexact_Lanczos.py:229-443); written against
the ``AbstractVector`` interface only, so it runs with ``CudaVector`` on the GPU and with the
CPU oracle in the parity tests.  On a machine that has the reference checked out, the
reference's own unchanged driver can be used with ``CudaVector`` instead.

Algorithm (per outer iteration): grow a Krylov list Y by solving (sigma - H) w = y for the last
nBlock vectors, Gram-Schmidt the results against Y, extend the small overlap/Hamiltonian
matrices S, Hm by one column each, Rayleigh-Ritz in the Loewdin-orthogonalised basis, order the
Ritz pairs with `pick`, test the change of the nBlock picked eigenvalues, restart from the
picked Ritz vectors after L-1 steps.

Deviations from the reference, both of which only remove crashes:
  * `saveTNSsEachIteration` defaults to False (the reference's default True raises
    AttributeError for non-TTNS vectors); when True, vectors exposing `.ttns.saveToHDF5` are saved.
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


def _getStatus(status, guessVector, nBlock):
    """Defaults of the status dictionary, overridden key by key by the user's dict
    (inexact_Lanczos.py:23-82; no whitelist)."""
    out = {"ref": [], "residual": np.inf, "nBlock": nBlock,
           "flagAddition": guessVector.hasExactAddition,
           "outerIter": 0, "innerIter": 0, "cumIter": 0, "iBlock": 0,
           "zeroVector": False, "isConverged": False, "lindep": False,
           "futileRestarts": 0, "startTime": time.time(), "runTime": 0.0,
           "KSmaxD": [], "fitmaxD": None, "phase": 1}
    if status is not None:
        out.update(status)
    return out


def generateSubspace(Hop, vec, sigma, eConv):
    """One shift-invert step: solve, then normalise unless the result is (numerically) zero,
    i.e. norm <= 0.001*eConv (inexact_Lanczos.py:84-105)."""
    typeClass = type(vec)
    out = typeClass.solve(Hop, vec, sigma)
    if typeClass.norm(out) > 0.001 * eConv:
        return typeClass.normalize(out), True
    return out, False


def _solve_block(typeClass, Hop, vecs, sigma):
    """The nBlock shifted solves of one Krylov step.  They are independent (inexact_Lanczos.py:319-320),
    so a vector class that offers `solveBlock` may advance them together (CudaVector: lock-step GCROT,
    one pass over H per Arnoldi step for the whole block); the results equal the one-by-one solves."""
    fn = getattr(typeClass, "solveBlock", None)
    if fn is None:
        return None
    return fn(Hop, vecs, sigma)


def checkConvergence(ev, eConv, status, printObj=None):
    """Relative change of the sorted first nBlock picked eigenvalues against the previous
    step's, from the second cumulative step on (inexact_Lanczos.py:115-143)."""
    nBlock = status["nBlock"]
    current = np.sort(ev[0:nBlock])
    converged = False
    if status["cumIter"] > 1:
        residual = eigenvalueResidual(current, status["ref"][-1])
        status["residual"] = residual
        converged = residual <= eConv
    status["isConverged"] = bool(converged)
    status["runTime"] = time.time() - status["startTime"]
    if printObj is not None:
        printObj.writeFile("summary", current, status)
    status["ref"].append(current)
    if len(status["ref"]) > 2:
        status["ref"].pop(0)
    return status


def terminateRestart(blockEnergies, eConv, status, num=3):
    """Count restarts that did not improve the block energies while lindep is flagged; give up
    after more than `num` (inexact_Lanczos.py:167-194)."""
    previous = status["ref"][0]
    if status["lindep"]:
        if eigenvalueResidual(blockEnergies, previous) > max(1e-9, eConv):
            status["futileRestarts"] += 1
    if status["futileRestarts"] > num:
        warnings.warn("Lindep and did not have fruitful restarts")
        return True
    return False


def analyzeStatus(status, maxit, L):
    """Continue unless converged or the last inner step of the last outer iteration was reached
    (inexact_Lanczos.py:197-222)."""
    if status["isConverged"]:
        return False
    if status["outerIter"] == maxit - 1 and status["innerIter"] == L - 1:
        print("Alert: Lanczos iterations is not converged!")
        return False
    return True


def _extend_small_matrices(typeClass, H, Ylist, Smat, Hmat):
    fused = getattr(typeClass, "extendBoth", None)
    if fused is not None:  # one SpMV + one pass over the Krylov list for both columns
        return fused(H, Ylist, Smat, Hmat)
    Smat = typeClass.extendOverlapMatrix(Ylist, Smat)
    Hmat = typeClass.extendMatrixRepresentation(H, Ylist, Hmat)
    return Smat, Hmat


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
    Returns (ev, Ylist, status): ALL Ritz values ordered by `pick`, the Ritz vectors, the status.
    """
    if issubclass(type(v0), AbstractVector):
        v0 = [v0]
    else:
        assert isinstance(v0, (list, tuple, np.ndarray)), f"{v0=} {type(v0)=}"
    if Hsolve is None:
        Hsolve = H
    typeClass = type(v0[0])
    nBlock = len(v0)

    Ylist = list(v0)
    Smat = typeClass.overlapMatrix(Ylist)
    if not np.allclose(Smat, np.eye(nBlock), rtol=1e-3, atol=1e-3):
        if nBlock > 1:
            raise RuntimeError(f"Input vectors not orthogonalized: {Smat=}")
        Ylist[0].normalize()  # single guess: normalise quietly (inexact_Lanczos.py:292-295)
        Smat[0, 0] = 1
    Hmat = typeClass.matrixRepresentation(H, Ylist)

    status = _getStatus(status, Ylist[0], nBlock)
    if pick is None:
        pick = get_pick_function_close_to_sigma(sigma)
    assert callable(pick)
    printObj = LanczosRunLog(Ylist[0], sigma, L, maxit, eConv, checkFitTol, writeOut, eShift,
                             convertUnit, pick, status, outFileName, summaryFileName)
    printObj.fileHeader()

    ev = np.array([np.nan] * len(Ylist))
    aborted = False          # linear dependency (or an unusable zero vector): leave both loops
    keepGoing = True
    for outerIter in range(maxit):
        uSH = None           # Ritz coefficients of the current Krylov list, none yet
        status["outerIter"] = outerIter
        status["KSmaxD"] = [Ylist[0].maxD]
        status["fitmaxD"] = None
        for innerIter in range(1, L):
            status["innerIter"] = innerIter
            status["cumIter"] += 1
            # -- new directions: the last nBlock vectors, visited back to front
            fresh = []
            nonzero = True
            block = _solve_block(typeClass, Hsolve, [Ylist[-iBlock] for iBlock in range(1, nBlock + 1)], sigma) \
                if (nBlock > 1 and lockstep) else None
            for iBlock in range(1, nBlock + 1):
                if block is None:
                    out, nonzero = generateSubspace(Hsolve, Ylist[-iBlock], sigma, eConv)
                else:  # same test as generateSubspace, on the result of the lock-step solve
                    out = block[iBlock - 1]
                    nonzero = typeClass.norm(out) > 0.001 * eConv
                    if nonzero:
                        out = typeClass.normalize(out)
                if not nonzero:
                    status["zeroVector"] = True
                    warnings.warn(f"Alert: zero vector: ||inv(H-sigma)vec||={typeClass.norm(out):5.3e}")
                    break
                fresh.append(out)
            if not nonzero:
                # inexact_Lanczos.py:321-329: leave the Krylov loop; the code after it restarts
                # from the last Ritz vectors.  Without any (first step) there is nothing to
                # restart from (the reference fails on stale/unbound data there): stop.
                aborted = uSH is None
                break
            # -- orthogonalise in list order, append, grow S and H by one column each
            lindepProblem = False
            for iBlock in range(nBlock):
                status["iBlock"] = iBlock
                q = typeClass.orthogonalize_against_set(fresh[iBlock], Ylist)
                if q is None:
                    lindepProblem = True
                    if printObj.writeOut:
                        warnings.warn(f"Linear dependency problem in iteration {outerIter} "
                                      f"and microiteration {innerIter} for block state {iBlock},"
                                      f" abort current Lanczos iteration and restart.")
                    break
                Ylist.append(q.compress())
                status["KSmaxD"].append(Ylist[-1].maxD)
                Smat, Hmat = _extend_small_matrices(typeClass, H, Ylist, Smat, Hmat)
            printObj.writeFile("iteration", status)
            printObj.writeFile("overlap", Smat)
            printObj.writeFile("KSmaxD", status)
            if lindepProblem:
                # inexact_Lanczos.py:356-359: results are NaN, the solver returns
                ev = np.array([np.nan] * len(Ylist))
                aborted = True
                break
            # -- Rayleigh-Ritz in the Loewdin basis
            status, uS = lowdinOrthoMatrix(Smat, status)
            assert not status["lindep"]  # Gram-Schmidt above should have caught it
            ev, uv = diagonalizeHamiltonian(uS, Hmat, printObj)
            uSH = uS @ uv
            idx = pick(uSH, Ylist, ev)
            assert len(idx) == len(ev), f"{len(ev)=} {len(idx)=}"
            ev = ev[idx]
            uSH = uSH[:, idx]
            status = checkConvergence(ev, eConv, status, printObj)
            keepGoing = analyzeStatus(status, maxit, L)
            if saveTNSsEachIteration:
                os.makedirs(saveDir, exist_ok=True)
                extra = {"status": status, "eigencoefficients": uSH, "eigenvalues": ev}
                for iv, vec in enumerate(Ylist):
                    vec.ttns.saveToHDF5(f"{saveDir}/tns_{status['cumIter']}_{iv}.h5",
                                        additionalInformation=extra)
            if not keepGoing:
                break
        if aborted:
            break
        if not keepGoing:
            # final back-transformation of ALL Ritz vectors and orthonormality check
            Ylist = basisTransformation(Ylist, uSH)
            Smat = typeClass.overlapMatrix(Ylist)
            if not np.allclose(Smat, np.eye(len(Ylist)), rtol=checkFitTol, atol=checkFitTol):
                warnings.warn(f"Alert:Final eigenvectors are not properly fitted. S=\n{Smat}")
            status["fitmaxD"] = [item.maxD for item in Ylist]
            printObj.writeFile("fitmaxD", status)
            break
        # -- plain restart from the nBlock picked Ritz vectors (normalised, not re-orthogonalised)
        guesses = []
        for iBlock in range(nBlock):
            g = basisTransformation(Ylist, uSH[:, iBlock])
            guesses.append(typeClass.normalize(g[0]))
        Ylist = guesses
        Smat = typeClass.overlapMatrix(Ylist)
        Hmat = typeClass.matrixRepresentation(H, Ylist)
        if not np.allclose(Smat, np.eye(len(Ylist)), rtol=checkFitTol, atol=checkFitTol):
            warnings.warn(f"Alert:Final eigenvectors are not properly fitted. S=\n{Smat}")
            break
        evNew = sla.eigvalsh(Hmat, Smat)
        if terminateRestart(evNew, eConv, status):
            break
        status["fitmaxD"] = [item.maxD for item in Ylist]
        printObj.writeFile("fitmaxD", status)

    printObj.writeFile("results", ev)
    printObj.fileFooter()
    return ev, Ylist, status
