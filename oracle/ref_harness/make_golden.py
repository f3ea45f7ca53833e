"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in this
container.  TEST INFRASTRUCTURE; run as  `python oracle/ref_harness/make_golden.py`.

Every fixture stores the seeded inputs' recipe (in `meta`) and the reference's outputs, so the
tests can rebuild the inputs without the reference and compare the oracle / the CUDA path.
Cases:
  ops          per-operation outputs of NumpyVector on seeded vectors (numpyVector.py:57-238)
  solve_*      NumpyVector.solve outputs (gcrotmk, minres, exact "pardiso")
  lanczos_c1   examples/driver_numpyVector.py (BASELINE config 1)
  lanczos_t1   unittests/test_lanczos.py setup
  lanczos_blk  unittests/test_lanczosBlock.py setup
  lanczos_ho   unittests/test_stateFollowingHO.py setup (sinc-DVR shim, harness-dependent)
  feast_t1     unittests/test_feast.py setup
  fortran      the numbers of unittests/data_fortranCode.out (Polizzi's Fortran FEAST) plus the
               reference's calculateQuadrature/updateQ outputs for them
  lap_blk      block Lanczos on a 12^3 Laplacian + potential (C2's generator at small N)
  osc_1        single-vector Lanczos on a 600-dim coupled-oscillator Hamiltonian (C3's generator)
  feast_osc    FEAST (nc = 16 -> 8 retained nodes, m0 = 4) on the sparse 600-dim oscillator Hamiltonian
               (C5's structure at small N)
  lanczos_lindep  unittests/test_lanczosLINDEP.py setup (n=1200, loose solves rtol 1e-1, L=100: a
               30-vector Krylov list; with SciPy 1.18 the run converges WITHOUT tripping LINDEP —
               the reference's own test notes "may fail on some machines")

`python oracle/ref_harness/make_golden.py [case ...]` regenerates only the named cases (outputs
depend on the BLAS in the last digits, so untouched fixtures are left alone).
"""
import json
import math
import os
import sys
import tempfile
import warnings

import numpy as np
import scipy.linalg as la

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import harness  # noqa: E402

ref = harness.load()
from eigensolvers_b200 import hamiltonians as hm  # noqa: E402

NumpyVector = ref.numpyVector.NumpyVector
GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
summary = {}


def save(name, **arrays):
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **arrays)
    print("wrote", name, {k: np.asarray(v).shape for k, v in arrays.items()})


def opts(solver="gcrotmk", tol=1e-4, it=1000):
    return {"linearSystemArgs": {"linearSolver": solver, "linearIter": it, "linear_tol": tol}}


def run_lanczos(H, guess, sigma, L, maxit, eConv, pick=None, status=None):
    with warnings.catch_warnings():
        warnings.simplefilter("default")
        ev, vecs, st = ref.inexact_Lanczos.inexactLanczosDiagonalization(
            H, guess, sigma, L, maxit, eConv, pick=pick, status=status, writeOut=False,
            saveTNSsEachIteration=False)
    warnings.resetwarnings()
    return ev, vecs, st


def status_scalars(st):
    return {k: (bool(st[k]) if isinstance(st[k], (bool, np.bool_)) else int(st[k]))
            for k in ("outerIter", "innerIter", "cumIter", "isConverged", "lindep", "zeroVector", "futileRestarts")}


# ------------------------------------------------------------------------------------- ops
def case_ops():
    rng = np.random.default_rng(2024)
    n, m = 257, 5
    V = rng.standard_normal((m, n))
    Z = rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
    A, _, _ = hm.prescribed_spectrum(n, 300.0, seed=7)
    vs = [NumpyVector(V[i].copy(), opts()) for i in range(m)]
    zs = [NumpyVector(Z[i].copy(), opts()) for i in range(m)]
    out = dict(V=V, Z=Z, A=A)
    out["mul"] = (vs[0] * 1.7).array
    out["div"] = (vs[0] / 1.7).array
    out["cmul"] = ((0.3 - 0.8j) * zs[0]).array
    out["norm"] = vs[1].norm()
    out["znorm"] = zs[1].norm()
    c = vs[2].copy()
    c.normalize()
    out["normalized"] = c.array
    out["vdot"] = vs[0].vdot(vs[1])
    out["zvdot"] = zs[0].vdot(zs[1])
    out["zdot_unconj"] = zs[0].vdot(zs[1], conjugate=False)
    out["real"] = zs[0].real().array
    out["conj"] = zs[0].conjugate().array
    out["applyOp"] = vs[0].applyOp(A).array
    coeffs = [0.5, -1.25, 2.0, 0.125, -3.0]
    out["coeffs"] = np.array(coeffs)
    out["lincomb"] = NumpyVector.linearCombination(vs, coeffs).array
    out["zlincomb"] = NumpyVector.linearCombination(zs, [c * (1 + 0.5j) for c in coeffs]).array
    g = NumpyVector.orthogonalize_against_set(vs[4], vs[:4])
    out["gs"] = g.array
    gz = NumpyVector.orthogonalize_against_set(zs[4], zs[:4])
    out["zgs"] = gz.array
    dep = NumpyVector.linearCombination(vs[:3], [1.0, 2.0, -1.0])
    qs = [NumpyVector(q, opts()) for q in np.linalg.qr(V[:3].T)[0].T]
    out["gs_dep_is_none"] = NumpyVector.orthogonalize_against_set(dep / dep.norm(), qs) is None
    out["overlap"] = NumpyVector.overlapMatrix(vs)
    out["zoverlap"] = NumpyVector.overlapMatrix(zs)
    out["matrep"] = NumpyVector.matrixRepresentation(A, vs)
    out["ext_overlap"] = NumpyVector.extendOverlapMatrix(vs, NumpyVector.overlapMatrix(vs[:-1]))
    out["ext_matrep"] = NumpyVector.extendMatrixRepresentation(A, vs, NumpyVector.matrixRepresentation(A, vs[:-1]))
    save("ops", **out)


# ------------------------------------------------------------------------------------- solve
def case_solve():
    A, ev, _ = hm.prescribed_spectrum(100, 300.0, seed=10)
    rng = np.random.default_rng(5)
    b = rng.standard_normal(100)
    out = dict(A=A, b=b, sigma=30.0)
    for solver, tol in (("gcrotmk", 1e-4), ("gcrotmk", 1e-10), ("minres", 1e-4), ("minres", 1e-10)):
        x = NumpyVector.solve(A, NumpyVector(b.copy(), opts(solver, tol)), 30.0)
        out[f"x_{solver}_{tol:g}"] = x.array
    xr = NumpyVector.solve(A, NumpyVector(b.copy(), opts("gcrotmk", 1e-10)), 30.0, reverseGF=True)
    out["x_gcrotmk_reverse"] = xr.array
    z = 163.0 + 2.0j
    xz = NumpyVector.solve(A, NumpyVector(b.copy(), opts("gcrotmk", 1e-10)), z, opType="gen")
    out["z"] = z
    out["x_gcrotmk_complex"] = xz.array
    xe = NumpyVector.solve(A, NumpyVector(b.copy(), {"linearSystemArgs": {"linearSolver": "pardiso"}}), 30.0)
    out["x_exact"] = np.asarray(xe.array).ravel()
    save("solve", **out)


# ------------------------------------------------------------------------------------- drivers
def case_c1():
    # examples/driver_numpyVector.py:27-43
    n = 100
    ev = np.linspace(1, 300, n)
    np.random.seed(10)
    Q = la.qr(np.random.rand(n, n))[0]
    A = Q.T @ np.diag(ev) @ Q
    Y0 = np.random.random(n)
    lf, xf, st = run_lanczos(A, NumpyVector(Y0.copy(), opts()), 30, 6, 4, 1e-8)
    save("lanczos_c1", A=A, Y0=Y0, ev=lf, vecs=np.array([v.array for v in xf]), exact=ev)
    summary["lanczos_c1"] = dict(status_scalars(st), nearest=float(lf[np.argmin(abs(lf - 30))]))


def case_t1():
    # unittests/test_lanczos.py:14-41
    n = 100
    ev = np.linspace(1, 200, n)
    np.random.seed(1212)
    Q = la.qr(np.random.rand(n, n))[0]
    A = Q.T @ np.diag(ev) @ Q
    Y0 = np.random.random(n)
    lf, xf, st = run_lanczos(A, NumpyVector(Y0.copy(), opts()), 30, 6, 4, 1e-6,
                             pick=ref.util_funcs.get_pick_function_close_to_sigma(30))
    save("lanczos_t1", A=A, Y0=Y0, ev=lf, vecs=np.array([v.array for v in xf]), exact=ev)
    summary["lanczos_t1"] = status_scalars(st)


def case_blk():
    # unittests/test_lanczosBlock.py:13-46
    n, nBlock, iBlock = 100, 3, 5
    ev = np.linspace(1, 200, n)
    ev[iBlock:iBlock + nBlock] = ev[iBlock]
    np.random.seed(1212)
    Q = la.qr(np.random.rand(n, n))[0]
    A = Q.T @ np.diag(ev) @ Q
    Ys = la.qr(np.random.rand(n, nBlock), mode="economic")[0]
    sigma = ev[iBlock] + nBlock / 2
    guess = [NumpyVector(Ys[:, i].copy(), opts()) for i in range(nBlock)]
    lf, xf, st = run_lanczos(A, guess, sigma, 6, 4, 1e-6,
                             pick=ref.util_funcs.get_pick_function_close_to_sigma(sigma))
    save("lanczos_blk", A=A, Ys=Ys, sigma=sigma, ev=lf, vecs=np.array([v.array for v in xf]), exact=ev)
    summary["lanczos_blk"] = status_scalars(st)


def case_ho():
    # unittests/test_stateFollowingHO.py:13-43 with the sinc-DVR shim
    import basis
    N = 45
    b = basis.SincInfInf(basis.SincInfInf.getOptions(N=N, xRange=[-10, 10]))
    H = -b.mat_dx2 + np.diag(b.xi ** 2)
    evE, uvE = la.eigh(H)
    sigma = 13.1
    o = opts("gcrotmk", 1e-4, 30000)
    idx = ref.util_funcs.find_nearest(evE, sigma)[0]
    ovlpRef = NumpyVector(uvE[:, idx + 1].copy(), o)
    np.random.seed(13)
    Y0 = np.random.random(N)
    lf, xf, st = run_lanczos(H, NumpyVector(Y0.copy(), o), sigma, 16, 200, 1e-10,
                             pick=ref.util_funcs.get_pick_function_maxOvlp(ovlpRef))
    save("lanczos_ho", H=H, Y0=Y0, sigma=sigma, ovlpRef=ovlpRef.array, energyRef=evE[idx + 1], ev=lf,
         vecs=np.array([v.array for v in xf]))
    summary["lanczos_ho"] = status_scalars(st)


def case_feast():
    # unittests/test_feast.py:14-50
    n = 100
    ev = np.linspace(1, 200, n)
    np.random.seed(10)
    Q = la.qr(np.random.rand(n, n))[0]
    A = Q.T @ np.diag(ev) @ Q
    m0 = 6
    Y0 = np.random.random((n, m0))
    for i in range(m0):
        Y0[:, i] = np.ones(n) * (i + 1)
    Y1 = la.qr(Y0, mode="economic")[0]
    Y = [NumpyVector(Y1[:, i].copy(), opts("gcrotmk", 1e-2)) for i in range(m0)]
    evf, uvf, st = ref.feast.feastDiagonalization(A, Y, 8, "legendre", 160.0, 166.0, 1e-10, 20, writeOut=False)
    warnings.resetwarnings()
    save("feast_t1", A=A, Y1=Y1, ev=evf, vecs=np.array([v.array for v in uvf]), exact=ev)
    summary["feast_t1"] = dict(outerIter=int(st["outerIter"]), residual=float(st["residual"]),
                               isConverged=bool(st["isConverged"]))


def case_fortran():
    # unittests/test_feast_fortran.py:14-24 parser, data file lines 1-157
    fn = os.path.join(harness.REFERENCE, "unittests", "data_fortranCode.out")
    amat = np.loadtxt(fn, dtype=float, skiprows=1, max_rows=4)
    guess = np.loadtxt(fn, dtype=complex, skiprows=6, max_rows=3)
    xe = np.loadtxt(fn, dtype=float, skiprows=12, max_rows=8)
    we = np.loadtxt(fn, dtype=float, skiprows=22, max_rows=8)
    theta = np.loadtxt(fn, dtype=float, skiprows=32, max_rows=8)
    zne = np.loadtxt(fn, dtype=complex, skiprows=42, max_rows=8)
    Qe = np.array([np.loadtxt(fn, dtype=complex, skiprows=62 + k * 5, max_rows=3) for k in range(8)])
    Qf = np.array([np.loadtxt(fn, dtype=float, skiprows=102 + k * 5, max_rows=3) for k in range(8)])
    # reference's own quadrature accumulation for the same data (test_Q, :105-127)
    order = [4, 3, 5, 2, 6, 1, 7, 0]
    gk, wk = ref.util_funcs.quadraturePointsWeights(8, "legendre", positiveHalf=False)
    r = 1.0
    th = np.array([-(np.pi * 0.5) * (g - 1) for g in gk])[order]
    wk = wk[order]
    Y = [NumpyVector(guess[i].copy(), {"linearSystemArgs": {"linearSolver": "pardiso"}}) for i in range(3)]
    Q = [np.nan] * 3
    Qref = []
    for k in range(8):
        z = 4.0 + r * math.cos(th[k]) + r * 0.3 * 1.0j * math.sin(th[k])
        for im0 in range(3):
            qk = ref.feast.calculateQuadrature(amat, Y[im0], z, r, th[k], wk[k], 0.3)
            Q = ref.feast.updateQ(Q, im0, qk, k)
        Qref.append(np.array([Q[i].array for i in range(3)]))
    save("fortran", amat=amat, guess=guess, xe=xe, we=we, theta=theta, zne=zne, Qe=Qe, Q=Qf,
         Q_reference=np.array(Qref), order=np.array(order))


def case_lap_blk():
    H = hm.laplacian3d(12, seed=2, W=1.0)
    evs = np.linalg.eigvalsh(H.toarray())
    sigma = ref.util_funcs.calculateTarget(evs, 10)
    guess = hm.orthonormal_block(H.shape[0], 4, seed=3)
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-4, "linear_atol": 1e-4}}
    lf, xf, st = run_lanczos(H, [NumpyVector(g.copy(), o) for g in guess], sigma, 6, 20, 1e-8)
    save("lap_blk", sigma=sigma, ev=lf, vecs=np.array([v.array for v in xf]), exact=evs)
    summary["lap_blk"] = status_scalars(st)


def case_osc():
    H, om = hm.coupled_oscillators((6, 5, 5, 4), coupling=0.1, seed=1)
    evs = np.linalg.eigvalsh(H.toarray())
    sigma = ref.util_funcs.calculateTarget(evs, 8)
    rng = np.random.default_rng(4)
    y0 = rng.standard_normal(H.shape[0])
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-4, "linear_atol": 1e-4}}
    lf, xf, st = run_lanczos(H, NumpyVector(y0.copy(), o), sigma, 8, 20, 1e-10)
    save("osc_1", sigma=sigma, y0=y0, ev=lf, vecs=np.array([v.array for v in xf]), exact=evs)
    summary["osc_1"] = status_scalars(st)


def case_lindep():
    # unittests/test_lanczosLINDEP.py:9-31
    n = 1200
    ev = np.linspace(1, 400, n)
    np.random.seed(10)
    Q = la.qr(np.random.rand(n, n))[0]
    A = Q.T @ np.diag(ev) @ Q
    Y0 = np.random.random(n)
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 500, "linear_tol": 1e-1}}
    lf, xf, st = run_lanczos(A, NumpyVector(Y0.copy(), o), 390, 100, 1000, 1e-12,
                             status={"writeOut": False, "writePlot": False})
    # A is rebuilt from the seed by the tests (11 MB dense); only outputs are stored
    save("lanczos_lindep", Y0=Y0, ev=lf, vecs=np.array([v.array for v in xf[:4]]), exact=ev)
    summary["lanczos_lindep"] = dict(status_scalars(st), n_vectors=len(xf))


def case_feast_osc():
    # BASELINE config 5's structure at small N: FEAST on the sparse coupled-oscillator Hamiltonian,
    # nc = 16 -> the reference's positiveHalf rule keeps 8 nodes (upper-right quadrant), m0 = 4,
    # gcrotmk rtol 1e-2 (unittests/test_feast.py:33), window around two analytic levels
    H, om = hm.coupled_oscillators((6, 5, 5, 4), coupling=0.1, seed=1)
    lev = hm.oscillator_levels(om, 0.1, 12, max_quanta=6)
    eMin, eMax = 0.5 * (lev[0] + lev[1]), 0.5 * (lev[2] + lev[3])
    Q = np.linalg.qr(np.random.default_rng(7).standard_normal((H.shape[0], 4)))[0]
    Y = [NumpyVector(np.ascontiguousarray(Q[:, i]), opts("gcrotmk", 1e-2, 2000)) for i in range(4)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        evf, uvf, st = ref.feast.feastDiagonalization(H, Y, 16, "legendre", eMin, eMax, 1e-8, 12, writeOut=False)
    warnings.resetwarnings()
    save("feast_osc", Q=Q, eMin=eMin, eMax=eMax, ev=evf, vecs=np.array([v.array for v in uvf]), levels=lev)
    summary["feast_osc"] = dict(outerIter=int(st["outerIter"]), residual=float(st["residual"]),
                                isConverged=bool(st["isConverged"]))


if __name__ == "__main__":
    all_cases = dict(ops=case_ops, solve=case_solve, c1=case_c1, t1=case_t1, blk=case_blk, ho=case_ho,
                     feast=case_feast, fortran=case_fortran, lap_blk=case_lap_blk, osc=case_osc,
                     lindep=case_lindep, feast_osc=case_feast_osc)
    chosen = sys.argv[1:] or list(all_cases)
    if sys.argv[1:] and os.path.exists(os.path.join(GOLD, "summary.json")):
        summary.update(json.load(open(os.path.join(GOLD, "summary.json"))))
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)  # the drivers open files / saveTNSs in the CWD
        for name in chosen:
            all_cases[name]()
    import scipy
    summary.setdefault("_versions", dict(numpy=np.__version__, scipy=scipy.__version__))
    with open(os.path.join(GOLD, "summary.json"), "w") as fh:
        json.dump(summary, fh, indent=1, sort_keys=True)
    print(json.dumps(summary, indent=1, sort_keys=True))
