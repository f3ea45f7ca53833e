"""Pin the CPU oracle (oracle/numpy_vector.py, oracle/krylov.py) and the host drivers
(eigensolvers_b200/lanczos.py, contour.py, hostmath.py) against golden vectors produced by the
UNMODIFIED reference (tests/golden/*.npz, made by oracle/ref_harness/make_golden.py) and against
the reference's own Fortran-FEAST golden file.  Runs on CPU; no GPU, no /root/reference needed.
"""
import json
import math
import os
import warnings

import numpy as np
import pytest
import scipy.linalg as la

from eigensolvers_b200 import hamiltonians as hm
from eigensolvers_b200.contour import calculateQuadrature, feastDiagonalization, updateQ
from eigensolvers_b200.hostmath import (calculateTarget, find_nearest, get_pick_function_close_to_sigma,
                                        get_pick_function_maxOvlp, quadraturePointsWeights)
from eigensolvers_b200.lanczos import inexactLanczosDiagonalization
from oracle.numpy_vector import NumpyVectorOracle as NV

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def summary():
    with open(os.path.join(GOLD, "summary.json")) as fh:
        return json.load(fh)


def opts(solver="gcrotmk", tol=1e-4, it=1000):
    return {"linearSystemArgs": {"linearSolver": solver, "linearIter": it, "linear_tol": tol}}


@pytest.fixture(autouse=True)
def _reset_warning_filters():
    yield
    warnings.resetwarnings()  # a non-converged solve leaves 'error' mode behind


def test_ops_match_reference_bitwise():
    """Same numpy expressions in the same order => identical bits (numpyVector.py:57-238)."""
    g = gold("ops")
    V, Z, A = g["V"], g["Z"], g["A"]
    vs = [NV(V[i].copy(), opts()) for i in range(5)]
    zs = [NV(Z[i].copy(), opts()) for i in range(5)]
    coeffs = list(g["coeffs"])
    checks = {
        "mul": (vs[0] * 1.7).array, "div": (vs[0] / 1.7).array, "cmul": ((0.3 - 0.8j) * zs[0]).array,
        "norm": vs[1].norm(), "znorm": zs[1].norm(), "vdot": vs[0].vdot(vs[1]), "zvdot": zs[0].vdot(zs[1]),
        "zdot_unconj": zs[0].vdot(zs[1], conjugate=False), "real": zs[0].real().array,
        "conj": zs[0].conjugate().array, "applyOp": vs[0].applyOp(A).array,
        "lincomb": NV.linearCombination(vs, coeffs).array,
        "zlincomb": NV.linearCombination(zs, [c * (1 + 0.5j) for c in coeffs]).array,
        "gs": NV.orthogonalize_against_set(vs[4], vs[:4]).array,
        "zgs": NV.orthogonalize_against_set(zs[4], zs[:4]).array,
        "overlap": NV.overlapMatrix(vs), "zoverlap": NV.overlapMatrix(zs),
        "matrep": NV.matrixRepresentation(A, vs),
        "ext_overlap": NV.extendOverlapMatrix(vs, NV.overlapMatrix(vs[:-1])),
        "ext_matrep": NV.extendMatrixRepresentation(A, vs, NV.matrixRepresentation(A, vs[:-1])),
    }
    c = vs[2].copy()
    c.normalize()
    checks["normalized"] = c.array
    for key, val in checks.items():
        np.testing.assert_array_equal(np.asarray(val), g[key], err_msg=key)
    dep = NV.linearCombination(vs[:3], [1.0, 2.0, -1.0])
    qs = [NV(q, opts()) for q in np.linalg.qr(V[:3].T)[0].T]
    assert (NV.orthogonalize_against_set(dep / dep.norm(), qs) is None) == bool(g["gs_dep_is_none"])


def test_solve_matches_reference():
    g = gold("solve")
    A, b = g["A"], g["b"]
    for solver, tol in (("gcrotmk", 1e-4), ("gcrotmk", 1e-10), ("minres", 1e-4), ("minres", 1e-10)):
        x = NV.solve(A, NV(b.copy(), opts(solver, tol)), 30.0).array
        np.testing.assert_array_equal(x, g[f"x_{solver}_{tol:g}"], err_msg=f"{solver} {tol}")
    xr = NV.solve(A, NV(b.copy(), opts("gcrotmk", 1e-10)), 30.0, reverseGF=True).array
    np.testing.assert_array_equal(xr, g["x_gcrotmk_reverse"])
    xz = NV.solve(A, NV(b.copy(), opts("gcrotmk", 1e-10)), complex(g["z"]), opType="gen").array
    np.testing.assert_array_equal(xz, g["x_gcrotmk_complex"])
    xe = NV.solve(A, NV(b.copy(), {"linearSystemArgs": {"linearSolver": "pardiso"}}), 30.0).array
    np.testing.assert_array_equal(np.asarray(xe).ravel(), g["x_exact"])
    with pytest.raises(Exception, match="other than gcrotmk"):
        NV.solve(A, NV(b.copy(), {"linearSystemArgs": {"linearSolver": "cg"}}), 30.0)


def _run(H, guess, sigma, L, maxit, eConv, pick=None, status=None):
    with warnings.catch_warnings():
        warnings.simplefilter("default")
        return inexactLanczosDiagonalization(H, guess, sigma, L, maxit, eConv, pick=pick, status=status,
                                             writeOut=False)


def _check_status(st, name):
    ref = summary()[name]
    for key in ("outerIter", "innerIter", "cumIter", "isConverged", "lindep", "zeroVector", "futileRestarts"):
        assert st[key] == ref[key], (name, key, st[key], ref[key])


def test_lanczos_c1_matches_reference():
    """BASELINE config 1 (examples/driver_numpyVector.py:27-43): our host driver + the oracle vector
    reproduce the unmodified reference run bit for bit (same trajectory, same numbers)."""
    g = gold("lanczos_c1")
    ev, vecs, st = _run(g["A"], NV(g["Y0"].copy(), opts()), 30, 6, 4, 1e-8)
    _check_status(st, "lanczos_c1")
    np.testing.assert_array_equal(ev, g["ev"])
    np.testing.assert_array_equal(np.array([v.array for v in vecs]), g["vecs"])
    assert abs(find_nearest(ev, 30)[1] - find_nearest(g["exact"], 30)[1]) < 1e-8


def test_lanczos_unit_test_setups_match_reference():
    g = gold("lanczos_t1")
    ev, vecs, st = _run(g["A"], NV(g["Y0"].copy(), opts()), 30, 6, 4, 1e-6, pick=get_pick_function_close_to_sigma(30))
    _check_status(st, "lanczos_t1")
    np.testing.assert_array_equal(ev, g["ev"])
    np.testing.assert_array_equal(np.array([v.array for v in vecs]), g["vecs"])
    # the reference's own assertions (unittests/test_lanczos.py:78-93)
    assert abs(find_nearest(ev, 30)[1] - find_nearest(g["exact"], 30)[1]) <= 1e-4

    g = gold("lanczos_blk")
    sigma = float(g["sigma"])
    guess = [NV(g["Ys"][:, i].copy(), opts()) for i in range(3)]
    ev, vecs, st = _run(g["A"], guess, sigma, 6, 4, 1e-6, pick=get_pick_function_close_to_sigma(sigma))
    _check_status(st, "lanczos_blk")
    np.testing.assert_array_equal(ev, g["ev"])
    np.testing.assert_allclose(ev[:3], g["exact"][5:8], rtol=1e-6)  # test_lanczosBlock.py:54


def _lindep_matrix():
    # unittests/test_lanczosLINDEP.py:12-16 (seeded; the 11 MB dense matrix is not stored)
    n = 1200
    ev = np.linspace(1, 400, n)
    np.random.seed(10)
    Q = la.qr(np.random.rand(n, n))[0]
    return Q.T @ np.diag(ev) @ Q


def test_lanczos_lindep_setup_matches_reference():
    """unittests/test_lanczosLINDEP.py set-up (loose rtol 1e-1 solves, L=100): a 30-vector Krylov
    list, every new vector orthogonalised against all previous ones (numpyVector.py:121-145)."""
    g = gold("lanczos_lindep")
    A = _lindep_matrix()
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 500, "linear_tol": 1e-1}}
    ev, vecs, st = _run(A, NV(g["Y0"].copy(), o), 390, 100, 1000, 1e-12)
    _check_status(st, "lanczos_lindep")
    assert len(vecs) == summary()["lanczos_lindep"]["n_vectors"]
    # the solves are loose (rtol 1e-1) and the BLAS thread count changes their rounding, so only the
    # converged pair is pinned tightly; the unconverged Ritz values agree to the eigenvalue criterion
    np.testing.assert_allclose(ev, g["ev"], rtol=1e-7, atol=0)
    i, j = np.argmin(abs(ev - 390)), np.argmin(abs(g["ev"] - 390))
    assert i == j and abs(ev[i] - g["ev"][j]) <= 1e-11 * abs(g["ev"][j])
    assert abs(np.vdot(vecs[i].array, g["vecs"][j])) >= 1 - 1e-8


def test_lanczos_state_following_matches_reference():
    g = gold("lanczos_ho")
    o = opts("gcrotmk", 1e-4, 30000)
    pick = get_pick_function_maxOvlp(NV(g["ovlpRef"].copy(), o))
    ev, vecs, st = _run(g["H"], NV(g["Y0"].copy(), o), float(g["sigma"]), 16, 200, 1e-10, pick=pick)
    _check_status(st, "lanczos_ho")
    np.testing.assert_array_equal(ev, g["ev"])
    assert abs(ev[0] - g["energyRef"]) / abs(g["energyRef"]) <= 1e-4  # test_stateFollowingHO.py:50-52


def test_lanczos_sparse_generators_match_reference():
    """The C2/C3 generators at small N, including the reference's LINDEP abort (NaN result)."""
    g = gold("lap_blk")
    H = hm.laplacian3d(12, seed=2, W=1.0)
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-4, "linear_atol": 1e-4}}
    guess = [NV(v.copy(), o) for v in hm.orthonormal_block(H.shape[0], 4, seed=3)]
    ev, vecs, st = _run(H, guess, float(g["sigma"]), 6, 20, 1e-8)
    _check_status(st, "lap_blk")
    assert np.all(np.isnan(ev)) and np.all(np.isnan(g["ev"])) and len(ev) == len(g["ev"])

    g = gold("osc_1")
    H, om = hm.coupled_oscillators((6, 5, 5, 4), coupling=0.1, seed=1)
    ev, vecs, st = _run(H, NV(g["y0"].copy(), o), float(g["sigma"]), 8, 20, 1e-10)
    _check_status(st, "osc_1")
    np.testing.assert_array_equal(ev, g["ev"])
    np.testing.assert_array_equal(np.array([v.array for v in vecs]), g["vecs"])
    assert abs(ev[0] - g["exact"][8]) < 1e-8
    # the analytic normal-mode levels approximate the low states of the truncated basis
    lev = hm.oscillator_levels(om, 0.1, 12)
    np.testing.assert_allclose(lev[:3], g["exact"][:3], rtol=2e-3)


def test_feast_matches_reference():
    g = gold("feast_t1")
    Y = [NV(g["Y1"][:, i].copy(), opts("gcrotmk", 1e-2)) for i in range(6)]
    ev, vecs, st = feastDiagonalization(g["A"], Y, 8, "legendre", 160.0, 166.0, 1e-10, 20, writeOut=False)
    ref = summary()["feast_t1"]
    assert st["outerIter"] == ref["outerIter"] and st["isConverged"] == ref["isConverged"]
    assert st["residual"] == ref["residual"]
    np.testing.assert_array_equal(ev, g["ev"])
    np.testing.assert_array_equal(np.array([v.array for v in vecs]), g["vecs"])
    inside = [e for e in g["exact"] if 160.0 <= e <= 166.0]
    for e in inside:  # unittests/test_feast.py:113-119
        assert abs(find_nearest(ev, e)[1] - e) <= 1e-4


def test_feast_sparse_oscillator_matches_reference():
    """C5's structure at small N (sparse H, nc = 16 -> 8 retained nodes, m0 = 4): our FEAST driver on
    the oracle vector against the unmodified reference's run; the Ritz values inside the window are
    the analytic oscillator levels."""
    g = gold("feast_osc")
    H, om = hm.coupled_oscillators((6, 5, 5, 4), coupling=0.1, seed=1)
    Y = [NV(np.ascontiguousarray(g["Q"][:, i]), opts("gcrotmk", 1e-2, 2000)) for i in range(4)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ev, vecs, st = feastDiagonalization(H, Y, 16, "legendre", float(g["eMin"]), float(g["eMax"]), 1e-8, 12,
                                            writeOut=False)
    ref = summary()["feast_osc"]
    assert st["outerIter"] == ref["outerIter"]
    # the complex solves are loose (rtol 1e-2) and their rounding depends on the BLAS thread count:
    # same trajectory, values equal far below the solver tolerance
    np.testing.assert_allclose(ev, g["ev"], rtol=1e-9)
    lev = g["levels"]
    inside = lev[(lev > g["eMin"]) & (lev < g["eMax"])]
    got = np.sort([e for e in ev if g["eMin"] < e < g["eMax"]])
    assert len(got) == len(inside) == 2
    np.testing.assert_allclose(got, inside, rtol=0, atol=5e-6)


def test_fortran_feast_golden_vectors():
    """unittests/test_feast_fortran.py:56-127 against Polizzi's Fortran FEAST numbers
    (data_fortranCode.out), with our quadrature / contour code and the oracle's exact solve."""
    g = gold("fortran")
    order = list(g["order"])
    gk, wk = quadraturePointsWeights(8, "legendre", positiveHalf=False)
    np.testing.assert_allclose(g["xe"], gk[order], rtol=1e-5, atol=0)
    np.testing.assert_allclose(g["we"], wk[order], rtol=1e-5, atol=0)
    theta = np.array([-(np.pi * 0.5) * (x - 1) for x in gk])[order]
    np.testing.assert_allclose(g["theta"], theta, rtol=1e-5, atol=0)
    r, f = 1.0, 0.3
    zne = np.array([4.0 + r * math.cos(t) + r * f * 1.0j * math.sin(t) for t in theta])
    np.testing.assert_allclose(g["zne"], zne, rtol=1e-5, atol=0)
    Y = [NV(g["guess"][i].copy(), {"linearSystemArgs": {"linearSolver": "pardiso"}}) for i in range(3)]
    wko = wk[order]
    Q = [None] * 3
    for k in range(8):
        Qe = np.array([NV.solve(g["amat"], Y[i], zne[k]).array for i in range(3)])
        np.testing.assert_allclose(Qe, g["Qe"][k], rtol=1e-5, atol=0)       # test_Qe
        for i in range(3):
            qk = calculateQuadrature(g["amat"], Y[i], zne[k], r, theta[k], wko[k], f)
            Q = updateQ(Q, i, qk, k)
        Qk = np.array([Q[i].array for i in range(3)])
        np.testing.assert_allclose(Qk, g["Q"][k], rtol=1e-5, atol=0)        # test_Q (Fortran numbers)
        np.testing.assert_array_equal(Qk, g["Q_reference"][k])               # reference's own output


def test_hostmath_small_cases():
    ev = np.array([1.0, 2.0, 4.0, 4.5])
    assert calculateTarget(ev, 1) == 2.0 + 0.25 * 1.0
    with pytest.raises(AssertionError):
        calculateTarget(np.array([1.0, 1.0, 2.0]), 1)
    gk, wk = quadraturePointsWeights(8, "legendre")
    assert len(gk) == 4 and np.all(gk > 0)
    gk, wk = quadraturePointsWeights(4, "trapezoidal", positiveHalf=False)
    np.testing.assert_allclose(gk, [-1.5, -1.0, -0.5, 0.0])                 # util_funcs.py:14-27 quirk
    np.testing.assert_allclose(wk, 0.4)


def test_driver_error_paths():
    A, ev, Y0 = hm.prescribed_spectrum(60, 100.0, seed=3)
    two = [NV(Y0.copy(), opts()), NV(Y0.copy() * 1.001, opts())]
    with pytest.raises(RuntimeError, match="not orthogonalized"):             # inexact_Lanczos.py:289-291
        _run(A, two, 30, 4, 2, 1e-6)
    with pytest.raises(AssertionError):
        _run(A, "nonsense", 30, 4, 2, 1e-6)
    # non-converged inner solve raises (numpyVector.py:175-177)
    with pytest.raises(UserWarning):
        _run(A, NV(Y0.copy(), opts("gcrotmk", 1e-14, it=1)), 30.3, 4, 2, 1e-6)


def test_mirror_drivers_write_their_run_logs(tmp_path, monkeypatch):
    """writeOut=True (the drivers' default): the mirror drivers write the two files per run the
    reference's printUtils.py writes (iteration log + summary table, same default names), one summary
    line per cumulative Krylov step / FEAST iteration; the results do not depend on logging."""
    from eigensolvers_b200.contour import feastDiagonalization
    from eigensolvers_b200.lanczos import inexactLanczosDiagonalization
    from oracle.numpy_vector import NumpyVectorOracle as NV
    monkeypatch.chdir(tmp_path)
    g = gold("lanczos_c1")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ev, vecs, st = inexactLanczosDiagonalization(g["A"], NV(g["Y0"].copy(), opts()), 30, 6, 4, 1e-8)   # writeOut default
    warnings.resetwarnings()
    np.testing.assert_array_equal(ev, g["ev"])
    it_log, summ = (tmp_path / "iterations_lanczos.out").read_text(), (tmp_path / "summary_lanczos.out").read_text()
    assert "inexact Lanczos" in it_log and "overlap matrix (condition number" in it_log and "final eigenvalues" in it_log
    rows = [ln for ln in summ.splitlines() if ln and not ln.startswith("#")]
    assert len(rows) == st["cumIter"] and rows[-1].split()[2] == str(st["cumIter"])
    assert abs(float(rows[-1].split()[4]) - np.sort(ev[:1])[0]) < 1e-10
    gf = gold("feast_t1")
    Y = [NV(gf["Y1"][:, i].copy(), opts("gcrotmk", 1e-2)) for i in range(6)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        evf, vf, stf = feastDiagonalization(gf["A"], Y, 8, "legendre", 160.0, 166.0, 1e-10, 3, outFileName="f_it.out",
                                            summaryFileName="f_sum.out")
    warnings.resetwarnings()
    assert "FEAST" in (tmp_path / "f_it.out").read_text()
    assert len([ln for ln in (tmp_path / "f_sum.out").read_text().splitlines() if ln and not ln.startswith("#")]) == 2
