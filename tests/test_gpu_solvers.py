"""GPU parity of CudaVector.solve against the reference's call into SciPy
(numpyVector.py:147-178 -> scipy.sparse.linalg.gcrotmk / minres) on the same (H, b, sigma,
options).  Per SURVEY §8c nothing in the reference pins the iterative solves at vector level, so
the bar is: same convergence verdict, residual within the requested tolerance, and solutions
that agree to the solver tolerance.
"""
import warnings

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu


def _opts(solver, tol, it=1000):
    return {"linearSystemArgs": {"linearSolver": solver, "linearIter": it, "linear_tol": tol, "linear_atol": 1e-4}}


def _scipy_solve(H, b, sigma, solver, tol, atol, maxiter, reverse=False):
    n = H.shape[0]
    dtype = np.result_type(sigma, H.dtype, b.dtype)
    if not reverse:
        lin = spla.LinearOperator((n, n), matvec=lambda x: sigma * x - H @ x, dtype=dtype)
    else:
        lin = spla.LinearOperator((n, n), matvec=lambda x: H @ x - sigma * x, dtype=dtype)
    if solver == "gcrotmk":
        return spla.gcrotmk(lin, b, None, rtol=tol, atol=atol, maxiter=maxiter)
    return spla.minres(lin, b, None, rtol=tol, maxiter=maxiter)


def _cases():
    from eigensolvers_b200 import hamiltonians as hm
    A, ev, _ = hm.prescribed_spectrum(100)
    yield "dense100", A, 30.0
    yield "lap12", hm.laplacian3d(12), 0.9
    H, om = hm.coupled_oscillators((6, 5, 5, 4))
    lev = hm.oscillator_levels(om, 0.1, 12)
    yield "osc600", H, lev[4] + 0.25 * (lev[5] - lev[4])


@pytest.mark.parametrize("solver,tol", [("gcrotmk", 1e-4), ("gcrotmk", 1e-8), ("minres", 1e-4), ("minres", 1e-9)])
def test_solve_matches_scipy(rt, solver, tol):
    from eigensolvers_b200 import CudaVector
    for name, H, sigma in _cases():
        n = H.shape[0]
        rng = np.random.default_rng(11)
        b = rng.standard_normal(n)
        b /= np.linalg.norm(b)
        x_ref, info_ref = _scipy_solve(H, b, sigma, solver, tol, 1e-4, 1000)
        assert info_ref == 0, name
        X = CudaVector.solve(H, CudaVector(b, _opts(solver, tol)), sigma)
        x = X.array
        st = rt.last_solve
        assert st.info == 0, name
        r = b - (sigma * x - H @ x)
        r_ref = b - (sigma * x_ref - H @ x_ref)
        if solver == "gcrotmk":
            bound = max(1e-4, tol * np.linalg.norm(b))
            assert np.linalg.norm(r) <= bound * (1 + 1e-8), (name, np.linalg.norm(r))
            # both satisfy the same residual bound => they differ by at most A^-1 (r - r_ref)
            assert np.linalg.norm(r - r_ref) <= 2 * bound
        else:
            # MINRES stops on ||r|| / (||A|| ||x||) <= rtol (minres.py:292,327)
            assert np.linalg.norm(r) <= 5 * max(np.linalg.norm(r_ref), 1e-12), (name, np.linalg.norm(r), np.linalg.norm(r_ref))
        rel = np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)
        assert rel <= (50 * tol if tol >= 1e-6 else 1e-5), (name, rel)


def test_gcrotmk_tight_matches_direct(rt):
    """With a tight tolerance the iterative solve equals the exact solve of (sigma - H)."""
    from eigensolvers_b200 import CudaVector, hamiltonians as hm
    A, ev, _ = hm.prescribed_spectrum(100)
    b = np.random.default_rng(2).standard_normal(100)
    opts = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-13, "linear_atol": 0.0}}
    x = CudaVector.solve(A, CudaVector(b, opts), 30.0).array
    np.testing.assert_allclose(x, np.linalg.solve(30.0 * np.eye(100) - A, b), rtol=1e-8, atol=1e-10)
    xr = CudaVector.solve(A, CudaVector(b, opts), 30.0, reverseGF=True).array
    np.testing.assert_allclose(xr, -x, rtol=1e-8, atol=1e-10)


def test_gcrotmk_complex_shift(rt):
    """FEAST's complex contour point: (z - H) x = b with real b (feast.py:90)."""
    from eigensolvers_b200 import CudaVector, hamiltonians as hm
    A, ev, _ = hm.prescribed_spectrum(100, 200.0)
    z = 163.0 + 2.0j
    b = np.random.default_rng(3).standard_normal(100)
    opts = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-10, "linear_atol": 0.0}}
    X = CudaVector.solve(A, CudaVector(b, opts), z, opType="gen")
    assert X.dtype == np.complex128
    np.testing.assert_allclose(X.array, np.linalg.solve(z * np.eye(100) - A, b), rtol=1e-7, atol=1e-9)
    x_ref, info = _scipy_solve(A, b.astype(complex), z, "gcrotmk", 1e-10, 0.0, 1000)
    assert info == 0
    np.testing.assert_allclose(X.array, x_ref, rtol=1e-6, atol=1e-9)


def test_x0_and_zero_rhs(rt):
    from eigensolvers_b200 import CudaVector, hamiltonians as hm
    H = hm.laplacian3d(10)
    n = H.shape[0]
    b = np.random.default_rng(5).standard_normal(n)
    for solver in ("gcrotmk", "minres"):
        B = CudaVector(b, _opts(solver, 1e-10))
        x = CudaVector.solve(H, B, 0.5)
        x2 = CudaVector.solve(H, B, 0.5, x0=x)       # already converged start
        r1 = np.linalg.norm(b - (0.5 * x.array - H @ x.array))
        r2 = np.linalg.norm(b - (0.5 * x2.array - H @ x2.array))
        assert r2 <= max(r1 * (1 + 1e-6), 1e-10)     # a converged start is not made worse
        np.testing.assert_allclose(x2.array, x.array, rtol=1e-5, atol=1e-7)
        z = CudaVector.solve(H, CudaVector(np.zeros(n), _opts(solver, 1e-10)), 0.5)
        assert np.all(z.array == 0.0)


def test_nonconvergence_raises_like_reference(rt):
    """numpyVector.py:175-177: the warning is escalated to an exception (and the filter stays)."""
    from eigensolvers_b200 import CudaVector, hamiltonians as hm
    H = hm.laplacian3d(12)
    b = np.random.default_rng(6).standard_normal(H.shape[0])
    with warnings.catch_warnings():
        with pytest.raises(UserWarning):
            CudaVector.solve(H, CudaVector(b, _opts("gcrotmk", 1e-12, it=1)), 3.0)
    with pytest.raises(Exception):
        CudaVector.solve(H, CudaVector(b, {"linearSystemArgs": {"linearSolver": "cg"}}), 3.0)


def test_iteration_counts_close_to_scipy(rt):
    """Matvec count of the device GCROT vs SciPy's on the same problem (CGS2 vs MGS: the
    Arnoldi recurrences agree in exact arithmetic, so the counts must be within a few %)."""
    from eigensolvers_b200 import CudaVector, hamiltonians as hm
    H, om = hm.coupled_oscillators((8, 6, 5, 5))
    lev = hm.oscillator_levels(om, 0.1, 12)
    sigma = lev[6] + 0.25 * (lev[7] - lev[6])
    n = H.shape[0]
    b = np.random.default_rng(8).standard_normal(n)
    b /= np.linalg.norm(b)
    count = [0]

    def mv(x):
        count[0] += 1
        return sigma * x - H @ x
    lin = spla.LinearOperator((n, n), matvec=mv, dtype=float)
    x_ref, info = spla.gcrotmk(lin, b, rtol=1e-6, atol=1e-4 * 0, maxiter=1000)
    assert info == 0
    opts = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-6, "linear_atol": 0.0}}
    x = CudaVector.solve(H, CudaVector(b, opts), sigma).array
    st = rt.last_solve
    assert abs(st.n_matvec - count[0]) <= max(3, 0.05 * count[0]), (st.n_matvec, count[0])
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-4


@pytest.mark.parametrize("eta,expect_second_pass", [(0.0, False), (0.1, None), (2.0, True)])
@pytest.mark.parametrize("n_odd", [False, True])
def test_fused_arnoldi_step_both_passes(rt, eta, expect_second_pass, n_odd):
    """The fused Arnoldi-step kernel (csrc/kernels_orth.cuh): never / selectively / always taking
    its second Gram-Schmidt pass gives the same solution to solver tolerance (the norm of the
    projected vector comes from |w|^2 - sum|h|^2 in pass 1 and from |w'|^2 - sum|h2|^2 in pass 2);
    odd n exercises the unpaired tail element of the 128-bit path."""
    import ctypes as C
    from eigensolvers_b200 import CudaVector, _lib, hamiltonians as hm
    H = hm.laplacian3d(13 if n_odd else 12)
    n = H.shape[0]
    assert (n % 2 == 1) == n_odd
    rng = np.random.default_rng(5)
    b = rng.standard_normal(n)
    sigma = 0.9
    tol = 1e-10
    x_ref, info_ref = _scipy_solve(H, b, sigma, "gcrotmk", tol, 0.0, 1000)
    assert info_ref == 0
    _lib.check(rt.lib.cv_ctx_set_reorth_eta(rt.ctx, C.c_double(eta)))
    try:
        o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": tol, "linear_atol": 0.0}}
        x = CudaVector.solve(H, CudaVector(b, o), sigma).array
        st = rt.last_solve
    finally:
        _lib.check(rt.lib.cv_ctx_set_reorth_eta(rt.ctx, C.c_double(0.1)))
    assert st.info == 0
    if expect_second_pass is True:
        assert st.n_reorth >= st.n_matvec - st.n_outer - 2
    if expect_second_pass is False:
        assert st.n_reorth == 0
    r = b - (sigma * x - H @ x)
    assert np.linalg.norm(r) <= tol * np.linalg.norm(b) * (1 + 1e-6)
    assert np.linalg.norm(x - x_ref) <= 1e-7 * np.linalg.norm(x_ref)
    if eta > 0:
        assert abs(st.n_matvec - 0) > 0


def test_gcrotmk_recycling_across_solves(rt):
    """Opt-in GCROT recycling (options["linearSystemArgs"]["recycle"] = True; SciPy's CU= argument,
    unused by the reference): a Lanczos-like sequence of solves with one operator and shift needs
    about as few matvecs as SciPy with CU and fewer than without recycling; every solve still
    meets its residual tolerance; a different shift starts from an empty ring."""
    from eigensolvers_b200 import CudaVector, DeviceOperator, hamiltonians as hm
    H, om = hm.coupled_oscillators((8, 6, 6, 5))
    lev = hm.oscillator_levels(om, 0.1, 12)
    sigma = lev[6] + 0.25 * (lev[7] - lev[6])
    n = H.shape[0]
    b0 = np.random.default_rng(2).standard_normal(n)
    b0 /= np.linalg.norm(b0)
    op = DeviceOperator.from_host(H)

    def sequence(solver):
        basis, counts = [b0], []
        for _ in range(5):
            x, nmv = solver(basis[-1])
            assert np.linalg.norm(basis[-1] - (sigma * x - H @ x)) <= 1e-6 * (1 + 1e-6)
            counts.append(nmv)
            y = x.copy()
            for q in basis:
                y -= (y @ q) * q
            basis.append(y / np.linalg.norm(y))
        return counts

    def gpu_solver(recycle):
        o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-6, "linear_atol": 0.0,
                                  "recycle": recycle}}

        def solve(b):
            x = CudaVector.solve(op, CudaVector(b, o), sigma).array
            return x, rt.last_solve.n_matvec
        return solve

    CU = []

    def scipy_solve(b):
        cnt = [0]

        def mv(v):
            cnt[0] += 1
            return sigma * v - H @ v
        lin = spla.LinearOperator((n, n), matvec=mv, dtype=float)
        x, info = spla.gcrotmk(lin, b, rtol=1e-6, atol=0.0, maxiter=1000, CU=CU, discard_C=False)
        assert info == 0
        return x, cnt[0]

    plain = sequence(gpu_solver(False))
    assert rt.last_solve.n_recycled == 0
    rec = sequence(gpu_solver(True))
    assert rt.last_solve.n_recycled > 0
    ref = sequence(scipy_solve)
    assert rec[0] == plain[0] or abs(rec[0] - plain[0]) <= 2          # first solve: nothing to recycle yet
    assert sum(rec) < 0.7 * sum(plain), (rec, plain)
    assert abs(sum(rec) - sum(ref)) <= 0.2 * sum(ref), (rec, ref)
    # another shift must not see the old ring
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-6, "linear_atol": 0.0, "recycle": True}}
    x = CudaVector.solve(op, CudaVector(b0, o), sigma + 0.01).array
    assert rt.last_solve.n_recycled == 0
    assert np.linalg.norm(b0 - ((sigma + 0.01) * x - H @ x)) <= 1e-6 * (1 + 1e-6)


# --------------------------------------------------------------------------------------------
# lock-step solves (cv_solve_batch / CudaVector.solveBlock): the independent solves of a block-Lanczos
# step (inexact_Lanczos.py:319-320) or of FEAST nodes (feast.py:190-201) advanced together
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nrhs", [2, 3, 4, 6])
@pytest.mark.parametrize("kind", ["dia", "general"])
@pytest.mark.parametrize("cplx", [False, True])
def test_lockstep_solves_match_single_solves(rt, nrhs, kind, cplx):
    """Every solve of a lock-step group is the same algorithm as a single solve: same verdict, solutions
    equal to solver accuracy, residuals within tolerance, a comparable number of operator applications
    (GCROT's path on these indefinite systems is sensitive to the reduction order: the SAME right-hand
    side twice in one group gives 1326 / 1324 applications, alone 1587 — tools/lockstep_diag.py) — with
    ONE shift for the block and with one shift per right-hand side."""
    from eigensolvers_b200 import CudaVector, DeviceOperator, hamiltonians as hm
    if kind == "dia":
        H = hm.laplacian3d(16, seed=2, W=1.0)
        op = DeviceOperator.from_host(H)
        assert op.format == "dia"
        base = 0.9
    else:
        rng = np.random.default_rng(3)
        A = sp.random(3000, 3000, density=0.003, random_state=3, format="csr")
        H = (0.1 * (A + A.T) + sp.diags(np.linspace(1.0, 9.0, 3000))).tocsr()
        op = DeviceOperator.from_host(H)
        assert op.format in ("sell", "csr")
        # shift in the middle of a gap with ~1 % of the spectrum below it: interior, but a system the
        # one-problem GCROT(20,20) itself solves to 1e-9 (at sigma = 4.3 on the unscaled matrix it
        # stagnates at 1e-6 after 60 000 applications — as SciPy's does: not a lock-step question)
        ev = np.linalg.eigvalsh(H.toarray())
        k = 30 + int(np.argmax(np.diff(ev[30:45])))
        base = 0.5 * (ev[k] + ev[k + 1])
    n = H.shape[0]
    rng = np.random.default_rng(nrhs)
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 3000, "linear_tol": 1e-9, "linear_atol": 0.0}}
    bs = [rng.standard_normal(n) for _ in range(nrhs)]
    for sigmas in ([base + (0.02j if cplx else 0.0)] * nrhs,
                   [base + 0.013 * q + (0.03j * (q + 1) if cplx else 0.0) for q in range(nrhs)]):
        B = [CudaVector(b, dict(o)) for b in bs]
        mv0 = rt.stats["matvecs"]
        singles = [CudaVector.solve(op, B[q], sigmas[q]).array for q in range(nrhs)]
        mv_single = rt.stats["matvecs"] - mv0
        mv0, ls0 = rt.stats["matvecs"], rt.stats.get("lockstep_solves", 0)
        block = CudaVector.solveBlock(op, B, sigmas if len(set(sigmas)) > 1 else sigmas[0])
        mv_block = rt.stats["matvecs"] - mv0
        assert rt.stats.get("lockstep_solves", 0) - ls0 == nrhs          # the batched path really ran
        assert abs(mv_block - mv_single) <= 0.3 * mv_single + 2, (mv_block, mv_single)
        for q in range(nrhs):
            xb = block[q].array
            res = np.linalg.norm(bs[q] - (sigmas[q] * xb - H @ xb)) / np.linalg.norm(bs[q])
            assert res < 5e-9, (q, res)
            # two solutions with residual <= 1e-9 |b| differ by up to cond(sigma - H) * 1e-9: comparable only
            # where the shift keeps the system well conditioned (complex shifts, distance >= 0.02 from the axis)
            if cplx:
                assert np.linalg.norm(xb - singles[q]) <= 1e-6 * np.linalg.norm(singles[q])


def test_lockstep_hard_shift(rt):
    """sigma inside a dense spectrum (BASELINE config 2's regime at 24^3): thousands of operator
    applications per solve, the solves of the group finish at different times and do their outer
    updates out of phase."""
    from eigensolvers_b200 import CudaVector, DeviceOperator
    from eigensolvers_b200.workloads import build_workload, solver_options
    w = build_workload("c2small")
    op = DeviceOperator.from_host(w["H"])
    o = solver_options(w)
    B = [CudaVector(g.copy(), dict(o)) for g in w["guesses"]]
    ls0 = rt.stats.get("lockstep_solves", 0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        block = CudaVector.solveBlock(op, B, w["sigma"])
    warnings.resetwarnings()
    assert rt.stats.get("lockstep_solves", 0) - ls0 == len(B)
    la = o["linearSystemArgs"]
    for q, g in enumerate(w["guesses"]):
        x = block[q].array
        res = np.linalg.norm(g - (w["sigma"] * x - w["H"] @ x))
        assert res <= 1.05 * max(la["linear_atol"], la["linear_tol"] * np.linalg.norm(g)), (q, res)


def test_lockstep_fallbacks_and_errors(rt):
    """solveBlock covers what the batched path does not by solving one at a time, and raises like solve."""
    from eigensolvers_b200 import CudaVector, hamiltonians as hm
    H = hm.laplacian3d(10)
    n = H.shape[0]
    rng = np.random.default_rng(0)
    bs = [rng.standard_normal(n) for _ in range(3)]
    ls0 = rt.stats.get("lockstep_solves", 0)
    out = CudaVector.solveBlock(H, [CudaVector(b, _opts("minres", 1e-8)) for b in bs], 0.5)      # MINRES: one at a time
    assert rt.stats.get("lockstep_solves", 0) == ls0 and len(out) == 3
    for b, x in zip(bs, out):
        xa = x.array
        assert np.linalg.norm(b - (0.5 * xa - H @ xa)) <= 1e-5 * np.linalg.norm(b)
    one = CudaVector.solveBlock(H, [CudaVector(bs[0], _opts("gcrotmk", 1e-8))], 0.5)             # a block of one
    assert len(one) == 1
    tight = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1, "linear_tol": 1e-14, "linear_atol": 0.0}}
    with pytest.raises(UserWarning):                                                             # numpyVector.py:175-177
        CudaVector.solveBlock(H, [CudaVector(b, dict(tight)) for b in bs], 3.1)
    warnings.resetwarnings()


def test_block_lanczos_lockstep_equals_sequential(rt):
    """The mirror driver with lock-step block solves against the same driver solving one by one:
    same number of Krylov steps, same eigenvalues (C2's generator at 12^3, four guesses)."""
    from eigensolvers_b200 import CudaVector, hamiltonians as hm
    from eigensolvers_b200.hostmath import calculateTarget
    from eigensolvers_b200.lanczos import inexactLanczosDiagonalization
    H = hm.laplacian3d(12, seed=2, W=1.0)
    evs = np.linalg.eigvalsh(H.toarray())
    sigma = calculateTarget(evs, 10)
    guess = hm.orthonormal_block(H.shape[0], 4, seed=3)
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-4, "linear_atol": 1e-4}}
    out = []
    for lock in (True, False):
        ls0 = rt.stats.get("lockstep_solves", 0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out.append(inexactLanczosDiagonalization(H, [CudaVector(v.copy(), dict(o)) for v in guess], sigma, 10, 20, 1e-8,
                                                     writeOut=False, lockstep=lock))
        warnings.resetwarnings()
        assert (rt.stats.get("lockstep_solves", 0) > ls0) == lock
    (ev_a, Y_a, st_a), (ev_b, Y_b, st_b) = out
    assert st_a["isConverged"] and st_b["isConverged"] and abs(st_a["cumIter"] - st_b["cumIter"]) <= 1
    np.testing.assert_allclose(np.sort(ev_a[:4]), np.sort(ev_b[:4]), rtol=1e-8)


# --------------------------------------------------------------------------------------------
# opt-in diagonal right preconditioner (SciPy's M= argument)
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("kind", ["csr", "kron"])
def test_jacobi_preconditioned_solve_matches_scipy(rt, cplx, kind):
    """options["linearSystemArgs"]["preconditioner"] = "jacobi": GCROT with M = diag(1/(sigma - H_ii)) as the
    right preconditioner, against scipy.sparse.linalg.gcrotmk(..., M=M) on the same inputs — the same
    stopping rule on the TRUE residual, a comparable number of operator applications, and far fewer than
    without (the oscillator Hamiltonian is diagonally dominant)."""
    from eigensolvers_b200 import CudaVector, KroneckerSumOperator, hamiltonians as hm
    from eigensolvers_b200.hostmath import calculateTarget
    dims = (12, 10, 8, 6)
    H, om = hm.coupled_oscillators(dims, coupling=0.1, seed=1)
    op = H if kind == "csr" else KroneckerSumOperator.coupled_oscillators(dims, coupling=0.1, seed=1)
    lev = hm.oscillator_levels(om, 0.1, 20, max_quanta=6)
    sigma = float(calculateTarget(lev, 8)) + (0.01j if cplx else 0.0)
    n = H.shape[0]
    b = np.random.default_rng(5).standard_normal(n)
    count = [0]

    def shifted(x):
        count[0] += 1
        return sigma * x - H @ x
    dtype = np.complex128 if cplx else np.float64
    lin = spla.LinearOperator((n, n), matvec=shifted, dtype=dtype)
    den = sigma - H.diagonal()
    M = spla.LinearOperator((n, n), matvec=lambda x: x / den, dtype=dtype)
    x_ref, info = spla.gcrotmk(lin, b.astype(dtype), M=M, rtol=1e-8, atol=0.0, maxiter=1000)
    assert info == 0
    mv_ref = count[0]
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-8, "linear_atol": 0.0,
                              "preconditioner": "jacobi"}}
    mv0 = rt.stats["matvecs"]
    x = CudaVector.solve(op, CudaVector(b, dict(o)), sigma).array
    mv_pre = rt.stats["matvecs"] - mv0
    res = np.linalg.norm(b - (sigma * x - H @ x)) / np.linalg.norm(b)
    assert res <= 1.05e-8, res
    assert np.linalg.norm(x - x_ref) <= 1e-6 * np.linalg.norm(x_ref)
    assert abs(mv_pre - mv_ref) <= 0.15 * mv_ref + 3, (mv_pre, mv_ref)
    o2 = {"linearSystemArgs": dict(o["linearSystemArgs"], preconditioner=None)}
    mv0 = rt.stats["matvecs"]
    CudaVector.solve(op, CudaVector(b, dict(o2)), sigma)
    mv_plain = rt.stats["matvecs"] - mv0
    assert mv_pre * 3 < mv_plain, (mv_pre, mv_plain)
    # a user-supplied diagonal (the diagonal of M itself) gives the same solve
    o3 = {"linearSystemArgs": dict(o["linearSystemArgs"], preconditioner=(1.0 / den).astype(dtype))}
    x3 = CudaVector.solve(op, CudaVector(b, dict(o3)), sigma).array
    assert np.linalg.norm(x3 - x) <= 1e-10 * np.linalg.norm(x)


def test_preconditioned_lanczos_reaches_the_same_eigenpair(rt):
    """The inexact Lanczos run with Jacobi-preconditioned solves converges to the eigenpair of the plain run
    (same eConv, same solver tolerance) with a fraction of the operator applications."""
    from eigensolvers_b200 import CudaVector, DeviceOperator
    from eigensolvers_b200.lanczos import inexactLanczosDiagonalization
    from eigensolvers_b200.workloads import build_workload, solver_options
    w = build_workload("c3small")
    op = DeviceOperator.from_host(w["H"])
    runs = []
    for pre in (None, "jacobi"):
        o = solver_options(w)
        o["linearSystemArgs"]["preconditioner"] = pre
        mv0 = rt.stats["matvecs"]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ev, vecs, st = inexactLanczosDiagonalization(op, CudaVector(w["guesses"][0].copy(), o), w["sigma"], w["L"], w["maxit"],
                                                         w["eConv"], writeOut=False)
        warnings.resetwarnings()
        assert st["isConverged"]
        runs.append((ev[0], vecs[0].array, rt.stats["matvecs"] - mv0))
    (e0, v0, mv0), (e1, v1, mv1) = runs
    assert abs(e0 - e1) <= max(w["eConv"], 1e-10) * abs(e0)
    assert abs(np.vdot(v0, v1)) >= 1 - 1e-8
    assert mv1 * 5 < mv0, (mv1, mv0)
