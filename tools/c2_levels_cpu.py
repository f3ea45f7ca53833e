"""Lowest eigenvalues of the C2 Hamiltonian (n^3 Laplacian + random potential) on the CPU with
scipy.sparse.linalg.eigsh (ARPACK, which="SA", no GPU code involved) -> tests/golden/c2_levels.json.

The GPU parity test of BASELINE config 2 (tests/test_gpu_drivers.py) compares the block-Lanczos
eigenvalues with these levels, and bench.py's sigma for c2 is calculateTarget(levels, 10).

    python tools/c2_levels_cpu.py [n=100]
"""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eigensolvers_b200 import hamiltonians as hm  # noqa: E402
from eigensolvers_b200.hostmath import calculateTarget  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
H = hm.laplacian3d(n, seed=2, W=1.0)
t0 = time.time()
ev, vec = spla.eigsh(H, k=24, which="SA", ncv=120, tol=1e-11, maxiter=200000)
order = np.argsort(ev)
ev, vec = ev[order], vec[:, order]
res = [float(np.linalg.norm(H @ vec[:, i] - ev[i] * vec[:, i])) for i in range(len(ev))]
out = {"n": n, "N": int(H.shape[0]), "levels": [float(e) for e in ev], "residuals": res,
       "sigma_k10": float(calculateTarget(ev, 10)), "seconds": time.time() - t0,
       "how": "scipy.sparse.linalg.eigsh(H, k=24, which='SA', ncv=120, tol=1e-11) on the CPU"}
path = os.path.join(ROOT, "tests", "golden", f"c2_levels_{n}.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out))
