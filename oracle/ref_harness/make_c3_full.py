"""Full-size parity anchor for BASELINE config 3 (TEST INFRASTRUCTURE).

Runs the UNMODIFIED reference (`/root/reference`, through harness.py) ONCE on bench.py's `c3`
workload at full size — the N = 2e7 coupled-oscillator Hamiltonian, the same seeded guess, sigma
and solver options — with `NumpyVector` on the CPU (hours on 8 cores), and stores what a GPU test
can compare without holding a 160 MB vector in git:

  eigenvalues, status scalars (cumIter, outerIter, innerIter, isConverged), matvecs per solve,
  top_idx/top_val   the largest-magnitude components of the converged eigenvector (>= 1 - 1e-12 of
                    its squared norm, capped at 400 000 entries) -> |<v_ref|v>| to ~1e-10
  sample_idx/sample_val   4096 seeded-index samples of the eigenvector
  probe_overlaps    overlaps with 8 seeded probe vectors (rng(77).standard_normal(N), normalised)

    python oracle/ref_harness/make_c3_full.py [workload]      (default c3; c3small for a dry run)

Output: tests/golden/<workload>_full.npz.  Progress goes to stdout (one line per solve).
"""
import json
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import harness  # noqa: E402

ref = harness.load()
from eigensolvers_b200.workloads import build_workload, solver_options  # noqa: E402

NumpyVector = ref.numpyVector.NumpyVector


class CountingOperator:
    """H with a matvec counter (shape, dtype, @ — all NumpyVector needs)."""

    def __init__(self, H):
        self.H, self.shape, self.dtype = H, H.shape, H.dtype
        self.count = 0

    def __matmul__(self, x):
        self.count += 1
        return self.H @ x


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    w = build_workload(name)
    print(f"[{name}] N={w['N']} nnz={w['nnz']} sigma={w['sigma']!r} generated in {w['gen_seconds']:.1f}s", flush=True)
    Hc = CountingOperator(w["H"])
    per_solve, t_solve = [], []
    orig_solve = NumpyVector.solve

    def counted(H, b, sigma, *a, **k):
        c0, t0 = Hc.count, time.time()
        out = orig_solve(H, b, sigma, *a, **k)
        per_solve.append(Hc.count - c0)
        t_solve.append(time.time() - t0)
        print(f"  solve {len(per_solve)}: {per_solve[-1]} matvecs, {t_solve[-1]:.0f}s", flush=True)
        return out
    NumpyVector.solve = staticmethod(counted)
    opts = solver_options(w)
    guess = NumpyVector(w["guesses"][0].copy(), opts)
    t0 = time.time()
    with warnings.catch_warnings():
        warnings.simplefilter("default")
        ev, vecs, st = ref.inexact_Lanczos.inexactLanczosDiagonalization(
            Hc, guess, w["sigma"], w["L"], w["maxit"], w["eConv"], writeOut=False, saveTNSsEachIteration=False)
    wall = time.time() - t0
    v = np.asarray(vecs[0].array)
    v = v / np.linalg.norm(v)
    N = w["N"]
    order = np.argsort(-np.abs(v), kind="stable")
    csum = np.cumsum(v[order] ** 2)
    k = int(np.searchsorted(csum, 1.0 - 1e-12)) + 1
    k = min(k, 400_000, N)
    top_idx = order[:k].astype(np.int64)
    rng = np.random.default_rng(123)
    sample_idx = np.sort(rng.choice(N, size=min(4096, N), replace=False)).astype(np.int64)
    prng = np.random.default_rng(77)
    probes = []
    for _ in range(8):
        p = prng.standard_normal(N)
        probes.append(float(np.dot(p, v) / np.linalg.norm(p)))
    true_res = float(np.linalg.norm(w["H"] @ v - ev[0] * v))
    meta = dict(workload=name, N=N, sigma=w["sigma"], L=w["L"], maxit=w["maxit"], eConv=w["eConv"], tol=w["tol"],
                wall_seconds=wall, cpu_count=os.cpu_count(), total_matvecs=int(Hc.count), true_residual=true_res,
                captured_norm2=float(csum[k - 1]), status={kk: (bool(st[kk]) if isinstance(st[kk], (bool, np.bool_)) else int(st[kk]))
                                                           for kk in ("outerIter", "innerIter", "cumIter", "isConverged")})
    out = os.path.join(ROOT, "tests", "golden", f"{name}_full.npz")
    np.savez_compressed(out, eigenvalues=np.asarray(ev, dtype=np.float64), per_solve_matvecs=np.asarray(per_solve),
                        per_solve_seconds=np.asarray(t_solve), top_idx=top_idx, top_val=v[top_idx],
                        sample_idx=sample_idx, sample_val=v[sample_idx], probe_overlaps=np.asarray(probes),
                        meta=json.dumps(meta))
    print("wrote", out, json.dumps(meta), flush=True)


if __name__ == "__main__":
    main()
