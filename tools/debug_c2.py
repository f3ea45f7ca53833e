import sys, warnings
import numpy as np
sys.path.insert(0, ".")
from eigensolvers_b200 import CudaVector, DeviceOperator, Runtime, hamiltonians as hm
from eigensolvers_b200.lanczos import inexactLanczosDiagonalization
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
H = hm.laplacian3d(n, seed=2, W=1.0)
sigma = 0.49075197166174706
o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 5000, "linear_tol": 1e-4, "linear_atol": 1e-4}}
guess = hm.orthonormal_block(H.shape[0], 4, seed=3)
rt = Runtime.get()
import os, ctypes
if os.environ.get("ETA"):
    from eigensolvers_b200 import _lib
    _lib.check(rt.lib.cv_ctx_set_reorth_eta(rt.ctx, ctypes.c_double(float(os.environ["ETA"]))))
op = DeviceOperator.from_host(H)
orig = CudaVector.solve
count = [0]
def traced(Hop, b, sg, *a, **k):
    try:
        return orig(Hop, b, sg, *a, **k)
    finally:
        st = rt.last_solve
        count[0] += 1
        print(count[0], "info", st.info, "outer", st.n_outer, "mv", st.n_matvec, "reorth", st.n_reorth, "resid", st.resid, "orth_loss", st.orth_loss, "safe", st.n_safe, flush=True)
CudaVector.solve = staticmethod(traced)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    try:
        ev, vecs, st = inexactLanczosDiagonalization(op, [CudaVector(g, dict(o)) for g in guess], sigma, 12, 20, 1e-8, writeOut=False)
        print("converged", st["isConverged"], ev[:4], st["cumIter"])
    except Exception as e:
        print("EXC", type(e).__name__, e)
