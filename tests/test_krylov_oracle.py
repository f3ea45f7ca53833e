"""Pin oracle/krylov.py (the numpy restatement of SciPy's gcrotmk / minres that the CUDA solvers
mirror) against SciPy itself, the third-party code the reference calls (numpyVector.py:161,163)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from eigensolvers_b200 import hamiltonians as hm
from oracle import krylov


def _problems():
    A, ev, _ = hm.prescribed_spectrum(100, 300.0, seed=10)
    yield "dense100", A, 30.0
    yield "lap10", hm.laplacian3d(10), 0.8
    H, om = hm.coupled_oscillators((6, 5, 5, 4))
    yield "osc600", H, 4.586


def _counting(H, sigma, dtype=float):
    cnt = [0]

    def mv(x):
        cnt[0] += 1
        return sigma * x - H @ x
    return spla.LinearOperator(H.shape, matvec=mv, dtype=dtype), mv, cnt


@pytest.mark.parametrize("tol", [1e-4, 1e-10])
def test_gcrotmk_mgs_equals_scipy(tol):
    for name, H, sigma in _problems():
        b = np.random.default_rng(1).standard_normal(H.shape[0])
        lin, mv, cnt = _counting(H, sigma)
        xs, info_s = spla.gcrotmk(lin, b, rtol=tol, atol=1e-4, maxiter=1000)
        n_scipy = cnt[0]
        x, info, nmv = krylov.gcrotmk(lambda v: sigma * v - H @ v, b, rtol=tol, atol=1e-4, maxiter=1000)
        assert info == info_s == 0, name
        assert nmv == n_scipy, (name, nmv, n_scipy)
        np.testing.assert_allclose(x, xs, rtol=1e-9, atol=1e-11, err_msg=name)


def test_gcrotmk_complex_equals_scipy():
    A, ev, _ = hm.prescribed_spectrum(100, 200.0, seed=10)
    z = 163.0 + 2.0j
    b = np.random.default_rng(3).standard_normal(100).astype(complex)
    lin = spla.LinearOperator(A.shape, matvec=lambda v: z * v - A @ v, dtype=complex)
    xs, info_s = spla.gcrotmk(lin, b, rtol=1e-10, atol=0.0, maxiter=1000)
    x, info, _ = krylov.gcrotmk(lambda v: z * v - A @ v, b, rtol=1e-10, atol=0.0, maxiter=1000)
    assert info == info_s == 0
    np.testing.assert_allclose(x, xs, rtol=1e-8, atol=1e-12)


def test_gcrotmk_cgs2_variant_tracks_mgs():
    """The device algorithm (CGS2 projections) against SciPy's MGS: same matvec count within a
    few steps and the same solution to the solver tolerance."""
    for name, H, sigma in _problems():
        b = np.random.default_rng(1).standard_normal(H.shape[0])
        mv = lambda v: sigma * v - H @ v  # noqa: E731
        x1, i1, n1 = krylov.gcrotmk(mv, b, rtol=1e-8, atol=0.0, maxiter=1000, orth="mgs")
        x2, i2, n2 = krylov.gcrotmk(mv, b, rtol=1e-8, atol=0.0, maxiter=1000, orth="cgs2")
        assert i1 == i2 == 0
        assert abs(n1 - n2) <= max(2, 0.03 * n1), (name, n1, n2)
        assert np.linalg.norm(x1 - x2) <= 1e-6 * np.linalg.norm(x1), name


def test_gcrotmk_nonconvergence_info():
    H = hm.laplacian3d(10)
    b = np.random.default_rng(1).standard_normal(H.shape[0])
    lin = spla.LinearOperator(H.shape, matvec=lambda v: 3.0 * v - H @ v, dtype=float)
    xs, info_s = spla.gcrotmk(lin, b, rtol=1e-12, atol=0.0, maxiter=2)
    x, info, _ = krylov.gcrotmk(lambda v: 3.0 * v - H @ v, b, rtol=1e-12, atol=0.0, maxiter=2)
    assert info == info_s == 2
    np.testing.assert_allclose(x, xs, rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("tol", [1e-4, 1e-10])
def test_minres_equals_scipy(tol):
    for name, H, sigma in _problems():
        b = np.random.default_rng(2).standard_normal(H.shape[0])
        lin, mv, cnt = _counting(H, sigma)
        xs, info_s = spla.minres(lin, b, rtol=tol, maxiter=1000)
        n_scipy = cnt[0]
        x, info, nmv = krylov.minres(lambda v: sigma * v - H @ v, b, rtol=tol, maxiter=1000)
        assert info == info_s, name
        assert nmv == n_scipy, (name, nmv, n_scipy)
        np.testing.assert_allclose(x, xs, rtol=1e-12, atol=1e-14, err_msg=name)


def test_gcrotmk_recycling_tracks_scipy():
    """GCROT recycling across successive solves with the same operator (SciPy's CU= argument,
    _gcrotmk.py:227-236, which the reference leaves unused): the restatement skips SciPy's
    re-orthogonalising QR (the kept c's are orthonormal already) and must still need about as few
    matvecs as SciPy with CU, fewer than without recycling, and reach the same residuals."""
    H, om = hm.coupled_oscillators((8, 6, 6, 5))
    lev = hm.oscillator_levels(om, 0.1, 12)
    sigma = lev[6] + 0.25 * (lev[7] - lev[6])
    n = H.shape[0]
    mv = lambda v: sigma * v - H @ v  # noqa: E731
    rng = np.random.default_rng(2)
    b0 = rng.standard_normal(n)
    b0 /= np.linalg.norm(b0)

    def sequence(solver):
        basis, counts = [b0], []
        for _ in range(5):
            x, nmv = solver(basis[-1])
            assert np.linalg.norm(basis[-1] - mv(x)) <= 1e-6 * (1 + 1e-8)
            counts.append(nmv)
            y = x.copy()
            for q in basis:
                y -= (y @ q) * q
            basis.append(y / np.linalg.norm(y))
        return counts

    def scipy_solver(CU):
        def solve(b):
            lin, _, cnt = _counting(H, sigma)
            x, info = spla.gcrotmk(lin, b, rtol=1e-6, atol=0.0, maxiter=1000, CU=CU, discard_C=False)
            assert info == 0
            return x, cnt[0]
        return solve

    def oracle_solver(CU):
        def solve(b):
            x, info, nmv = krylov.gcrotmk(mv, b, rtol=1e-6, atol=0.0, maxiter=1000, orth="cgs2", CU=CU)
            assert info == 0
            return x, nmv
        return solve

    plain = sum(sequence(oracle_solver(None)))
    ours = sum(sequence(oracle_solver([])))
    ref = sum(sequence(scipy_solver([])))
    assert ours < plain, (ours, plain)
    assert abs(ours - ref) <= 0.15 * ref, (ours, ref, plain)


def test_oracle_jacobi_option_is_scipys_M():
    """The oracle's counterpart of CudaVector's opt-in preconditioner option passes M = diag(1/(sigma - H_ii))
    to SciPy (gcrotmk's `M=`): same stopping rule on the true residual, far fewer operator applications on the
    diagonally dominant oscillator Hamiltonian."""
    import numpy as np
    from eigensolvers_b200 import hamiltonians as hm
    from eigensolvers_b200.hostmath import calculateTarget
    from oracle.numpy_vector import NumpyVectorOracle as NV
    H, om = hm.coupled_oscillators((8, 6, 5, 4), coupling=0.1, seed=1)
    lev = hm.oscillator_levels(om, 0.1, 20, max_quanta=6)
    sigma = float(calculateTarget(lev, 8))
    b = np.random.default_rng(1).standard_normal(H.shape[0])
    counts = []
    for pre in (None, "jacobi"):
        o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-8, "linear_atol": 0.0,
                                  "preconditioner": pre}}
        NV.matvec_count = 0
        x = NV.solve(H, NV(b.copy(), o), sigma).array
        counts.append(NV.matvec_count)
        assert np.linalg.norm(b - (sigma * x - H @ x)) <= 1.01e-8 * np.linalg.norm(b)
    assert counts[1] * 3 < counts[0], counts
