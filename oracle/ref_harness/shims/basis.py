"""Restatement of the one piece of the in-house `basis` module the NumpyVector examples use
(examples/stateFollowingHO.py:15-19): an infinite-range sinc DVR (Colbert & Miller, J. Chem.
Phys. 96, 1982 (1992)) with N points on [xmin, xmax].  `mat_dx2` is d^2/dx^2, `xi` the grid.
The exact grid convention of the original is not published: parity of this one example is
therefore pinned only up to the harness (oracle/__init__.py)."""
import numpy as np


class SincInfInf:
    @staticmethod
    def getOptions(N, xRange):
        return {"N": N, "xRange": xRange}

    def __init__(self, opts):
        N = opts["N"]
        a, b = opts["xRange"]
        self.N = N
        self.xi = np.linspace(a, b, N)
        dx = self.xi[1] - self.xi[0]
        i = np.arange(N)
        d = i[:, None] - i[None, :]
        with np.errstate(divide="ignore"):
            off = -2.0 * (-1.0) ** d / (d.astype(float) ** 2)
        off[d == 0] = -np.pi ** 2 / 3.0
        self.mat_dx2 = off / dx ** 2
