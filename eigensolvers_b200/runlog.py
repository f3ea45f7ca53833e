"""Plain-text run logs of the drivers (two files per run, like the reference's printUtils.py:
an iteration log with the small matrices and a summary table).  Logging is outside the hot path
(SURVEY §2 #6); the writers keep the reference's call protocol — ``fileHeader()``,
``writeFile(kind, *args)``, ``fileFooter()``, attribute ``writeOut`` — and its file names, with a
simplified layout.  Benchmarks run with ``writeOut=False``.
"""
import time

import numpy as np


def convert(arr, eShift=0.0, unit="au"):
    """Shift (and, for units other than atomic units, convert) energies for printing
    (printUtils.py:9-18).  Only 'au' is built in; other units need the reference's `util`."""
    arr = np.asarray(arr) + eShift
    if unit.lower() not in ("au", "a.u.", "hartree"):
        raise NotImplementedError(f"unit conversion to {unit!r} needs the reference's util.au2unit")
    return arr


class _RunLog:
    default_names = ("iterations.out", "summary.out")

    def __init__(self, writeOut, eShift, convertUnit, outFileName, summaryFileName):
        self.writeOut = bool(writeOut)
        self.eShift = eShift
        self.convertUnit = convertUnit
        self.outfile = self.sumfile = None
        if self.writeOut:
            self.outfile = open(outFileName or self.default_names[0], "w")
            self.sumfile = open(summaryFileName or self.default_names[1], "w")

    def _w(self, fh, text):
        if fh is not None:
            fh.write(text)

    def fileFooter(self):
        if self.writeOut:
            self._w(self.outfile, "\n*** computation complete ***\n")
            self.outfile.close()
            self.sumfile.close()
            self.outfile = self.sumfile = None

    def _matrix(self, title, mat):
        self._w(self.outfile, f"{title}\n{np.array2string(np.asarray(mat), precision=8, max_line_width=200)}\n")


class LanczosRunLog(_RunLog):
    """Protocol of printUtils.LanczosPrintUtils (printUtils.py:23-274)."""
    default_names = ("iterations_lanczos.out", "summary_lanczos.out")

    def __init__(self, guessVector, sigma, L, maxit, eConv, checkFitTol, writeOut, eShift,
                 convertUnit, pick, status, outFileName=None, summaryFileName=None):
        super().__init__(writeOut, eShift, convertUnit, outFileName, summaryFileName)
        self.info = dict(sigma=sigma, L=L, maxit=maxit, eConv=eConv, checkFitTol=checkFitTol,
                         vector=type(guessVector).__name__, options=getattr(guessVector, "options", {}))

    def fileHeader(self):
        if not self.writeOut:
            return
        self._w(self.outfile, "*** inexact Lanczos ***\n")
        for k, v in self.info.items():
            self._w(self.outfile, f"{k:14s} {v}\n")
        self._w(self.sumfile, "# it i nCum target Evalue(s) residual time(seconds)\n")

    def writeFile(self, kind, *args):
        if not self.writeOut:
            return
        if kind == "iteration":
            st = args[0]
            self._w(self.outfile, f"\n--- outer {st['outerIter']} inner {st['innerIter']} cumulative {st['cumIter']} ---\n")
        elif kind == "overlap":
            S = np.asarray(args[0])
            self._matrix(f"overlap matrix (condition number {np.linalg.cond(S):.3e})", S)
        elif kind == "hamiltonian":
            self._matrix(f"Hamiltonian matrix {args[1]}", convert(args[0], 0.0, self.convertUnit))
        elif kind == "eigenvalues":
            self._matrix("eigenvalues", convert(args[0], self.eShift, self.convertUnit))
        elif kind == "summary":
            evs, st = args
            vals = " ".join(f"{e:.12f}" for e in convert(evs, self.eShift, self.convertUnit))
            self._w(self.sumfile, f"{st['outerIter']} {st['innerIter']} {st['cumIter']} {self.info['sigma']} "
                                  f"{vals} {st['residual']:.6e} {st['runTime']:.3f}\n")
        elif kind == "results":
            self._matrix("final eigenvalues", convert(args[0], self.eShift, self.convertUnit))
        # "KSmaxD" / "fitmaxD" concern tensor-network bond dimensions only


class FeastRunLog(_RunLog):
    """Protocol of printUtils.FeastPrintUtils (printUtils.py:279-499)."""
    default_names = ("iterations_feast.out", "summary_feast.out")

    def __init__(self, guessVectors, nc, quad, eMin, eMax, eConv, maxit, writeOut, eShift,
                 convertUnit, status, outFileName=None, summaryFileName=None):
        super().__init__(writeOut, eShift, convertUnit, outFileName, summaryFileName)
        self.info = dict(nc=nc, quad=quad, eMin=eMin, eMax=eMax, eConv=eConv, maxit=maxit,
                         m0=len(guessVectors), vector=type(guessVectors[0]).__name__,
                         options=getattr(guessVectors[0], "options", {}))

    def fileHeader(self):
        if not self.writeOut:
            return
        self._w(self.outfile, "*** FEAST ***\n")
        for k, v in self.info.items():
            self._w(self.outfile, f"{k:14s} {v}\n")
        self._w(self.sumfile, "# it Evalue(s) residual time(seconds)\n")

    def writeFile(self, kind, *args):
        if not self.writeOut:
            return
        if kind == "iteration":
            self._w(self.outfile, f"\n--- FEAST iteration {args[0]['outerIter']} ---\n")
        elif kind == "overlap":
            self._matrix("overlap matrix", args[0])
        elif kind == "hamiltonian":
            self._matrix(f"Hamiltonian matrix {args[1]}", args[0])
        elif kind == "eigenvalues":
            self._matrix("eigenvalues", convert(args[0], self.eShift, self.convertUnit))
        elif kind == "summary":
            evs, residual, st = args
            vals = " ".join(f"{e:.12f}" for e in convert(evs, self.eShift, self.convertUnit))
            self._w(self.sumfile, f"{st['outerIter']} {vals} {residual:.6e} {st['runTime']:.3f}\n")
        elif kind == "results":
            self._matrix("final eigenvalues", convert(args[0], self.eShift, self.convertUnit))
