"""Per-solve statistics of single solves vs the same solves in lock step (development aid)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from eigensolvers_b200 import CudaVector, DeviceOperator, Runtime, hamiltonians as hm  # noqa: E402

rt = Runtime.get()
H = hm.laplacian3d(16, seed=2, W=1.0)
op = DeviceOperator.from_host(H)
n = H.shape[0]
rng = np.random.default_rng(2)
o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 3000, "linear_tol": 1e-9, "linear_atol": 0.0}}
bs = [rng.standard_normal(n) for _ in range(2)]
for rep in range(2):
    for q in range(2):
        x = CudaVector.solve(op, CudaVector(bs[q], dict(o)), 0.9)
        s = rt.last_solve
        xa = x.array
        print("single", q, "matvec", s.n_matvec, "outer", s.n_outer, "reorth", s.n_reorth, "safe", s.n_safe, "resid", s.resid,
              "true", np.linalg.norm(bs[q] - (0.9 * xa - H @ xa)) / np.linalg.norm(bs[q]), "loss", s.orth_loss)
import ctypes as C
from eigensolvers_b200 import _lib
for same in (False, True):
    B = [CudaVector(bs[0 if same else q], dict(o)) for q in range(2)]
    out = CudaVector.solveBlock(op, B, 0.9)
    print("lockstep same=%s" % same, rt.last_block_matvecs, [float(np.linalg.norm(bs[0 if same else q] - (0.9 * out[q].array - H @ out[q].array)) / np.linalg.norm(bs[q])) for q in range(2)],
          "reorth", rt.stats.get("reorth"), "safe", rt.stats.get("safe_solves"))

# the general-sparsity case of tests/test_gpu_solvers.py::test_lockstep_solves_match_single_solves
import scipy.sparse as sp
A = sp.random(3000, 3000, density=0.003, random_state=3, format="csr")
H2 = (A + A.T + sp.diags(np.linspace(1.0, 9.0, 3000))).tocsr()
op2 = DeviceOperator.from_host(H2)
ev = np.linalg.eigvalsh(H2.toarray())
print("format", op2.format, "nearest eigenvalue distance to 4.3:", np.min(np.abs(ev - 4.3)))
rng = np.random.default_rng(2)
for q in range(3):
    b = rng.standard_normal(3000)
    try:
        x = CudaVector.solve(op2, CudaVector(b, dict(o)), 4.3)
        s = rt.last_solve
        print("general single", q, "info", s.info, "matvec", s.n_matvec, "outer", s.n_outer, "resid", s.resid)
    except Exception as e:
        s = rt.last_solve
        print("general single", q, "EXC", type(e).__name__, "info", s.info, "matvec", s.n_matvec, "outer", s.n_outer, "resid", s.resid, "bnorm", s.b_norm)
    import warnings; warnings.resetwarnings()
