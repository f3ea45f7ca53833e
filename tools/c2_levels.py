"""One-off: lowest eigenvalues of the C2 Hamiltonian (100^3 Laplacian + random potential) by ARPACK
shift-invert at sigma = 0 with the GPU GCROT as the inverse, to place bench.py's sigma a quarter-gap
above the 11th level (util_funcs.calculateTarget) without a minutes-long CPU eigsh in every run."""
import sys
import time

import numpy as np
import scipy.sparse.linalg as spla

sys.path.insert(0, ".")
from eigensolvers_b200 import CudaVector, DeviceOperator, hamiltonians as hm  # noqa: E402
from eigensolvers_b200.hostmath import calculateTarget  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
H = hm.laplacian3d(n, seed=2, W=1.0)
N = H.shape[0]
op = DeviceOperator.from_host(H)
opts = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 5000, "linear_tol": 1e-11, "linear_atol": 0.0}}
count = [0]


def opinv(b):  # (H - 0)^-1 b = -(0 - H)^-1 b
    count[0] += 1
    return -CudaVector.solve(op, CudaVector(np.ascontiguousarray(b), dict(opts)), 0.0).array


t0 = time.time()
ev = spla.eigsh(H, k=24, sigma=0.0, which="LM", OPinv=spla.LinearOperator((N, N), matvec=opinv, dtype=float),
                tol=1e-9, return_eigenvectors=False)
ev = np.sort(ev)
print("N", N, "format", op.format, "solves", count[0], "seconds", round(time.time() - t0, 1))
print("levels", [float(f"{e:.12f}") for e in ev])
print("gaps", np.diff(ev)[:16])
print("sigma(k=10)", repr(float(calculateTarget(ev, 10))))
