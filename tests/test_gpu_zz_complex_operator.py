"""Complex-valued (Hermitian) H through CudaVector.

numpyVector.py:98-100 applies whatever `other @ self.array` accepts and :147-178 hands any H to
SciPy, so NumpyVector's applyOp / solve / matrixRepresentation work on a complex Hermitian matrix.
Here the operator carries its imaginary parts as a second CSR value stream (cv_op_set_imag) and
every operator application promotes real vectors to complex, as numpy does.  Checked against
numpy/scipy on the same seeded inputs.

The reference's DRIVERS do not run on such a matrix (checked with the unmodified files under
baseline/_ref): inexact_Lanczos.py aborts in its linear-dependency branch because
orthogonalize_against_set uses unconjugated products (numpyVector.py:133-145), and feast.py
integrates over the upper half contour only, which assumes a real symmetric matrix
(feast.py:185-201).  So there is no driver-level parity to assert and none is claimed.

(The file name sorts last on purpose: the newest feature runs after the established suite.)
"""
import warnings

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu


def _opts(tol=1e-4, solver="gcrotmk", atol=1e-4):
    return {"linearSystemArgs": {"linearSolver": solver, "linearIter": 1000, "linear_tol": tol, "linear_atol": atol}}


def _hermitian_sparse(n, density, seed, diag=None):
    """Seeded complex Hermitian CSR matrix: random off-diagonal entries (O(nnz) generator -- scipy's
    sp.random permutes all n*n positions with a legacy RandomState, minutes at n = 7e4) + real diagonal."""
    rng = np.random.default_rng(seed)
    k = max(1, int(density * n * n))
    rows, cols = rng.integers(0, n, k), rng.integers(0, n, k)
    vals = rng.random(k) + 1j * rng.random(k)
    Z = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()     # duplicates are summed
    H = (Z + Z.conj().T) * 0.5
    d = np.arange(1, n + 1, dtype=np.float64) if diag is None else diag
    H = (H + sp.diags(d)).tocsr()
    assert abs(H - H.conj().T).max() == 0
    return H


@pytest.fixture(autouse=True)
def _reset_warning_filters():
    yield
    warnings.resetwarnings()


@pytest.mark.parametrize("n,density", [(2, 1.0), (37, 0.3), (3000, 0.004), (70001, 0.0001)])
def test_apply_complex_operator(rt, n, density):
    from eigensolvers_b200 import CudaVector
    H = _hermitian_sparse(n, density, 3)
    rng = np.random.default_rng(5)
    xr = rng.standard_normal(n)
    xc = xr + 1j * rng.standard_normal(n)
    op = rt.operator_for(H)
    assert op.format == "csr" and op.dtype == np.complex128
    for x in (xr, xc):
        Y = CudaVector(x, _opts()).applyOp(H)
        ref = H @ x
        assert Y.dtype == np.complex128                      # numpy promotes a real x the same way
        np.testing.assert_allclose(Y.array, ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max())
    dense = np.asarray(H.todense()) if n <= 3000 else None
    if dense is not None:                                        # dense ndarray H takes the same path
        Y = CudaVector(xc, _opts()).applyOp(dense)
        np.testing.assert_allclose(Y.array, dense @ xc, rtol=1e-12, atol=1e-12 * np.abs(dense @ xc).max())


def test_complex_dtype_with_real_values_keeps_fast_formats(rt):
    from eigensolvers_b200 import CudaVector, hamiltonians as hm
    H = hm.laplacian3d(16)
    Hz = H.astype(np.complex128)
    op = rt.operator_for(Hz)
    assert op.dtype == np.float64 and op.format in ("dia", "sell")
    x = np.random.default_rng(1).standard_normal(H.shape[0])
    np.testing.assert_allclose(CudaVector(x, _opts()).applyOp(Hz).array, H @ x, rtol=1e-13, atol=1e-13)


def test_complex_operator_rejects_other_formats(rt):
    from eigensolvers_b200 import DeviceOperator
    H = _hermitian_sparse(2000, 0.003, 8)
    for fmt in ("sell", "dia"):
        with pytest.raises(NotImplementedError):
            DeviceOperator.from_host(H, fmt=fmt)
    op = DeviceOperator.from_host(H)
    with pytest.raises(RuntimeError):
        op.set_format("sell")


def test_matrix_representation_complex_operator(rt):
    from eigensolvers_b200 import CudaVector
    n = 900
    H = _hermitian_sparse(n, 0.01, 21)
    rng = np.random.default_rng(2)
    Vr = np.linalg.qr(rng.standard_normal((n, 5)))[0]
    Vc = np.linalg.qr(rng.standard_normal((n, 5)) + 1j * rng.standard_normal((n, 5)))[0]
    for V in (Vr, Vc):
        vs = [CudaVector(V[:, i].copy(), _opts()) for i in range(5)]
        M = CudaVector.matrixRepresentation(H, vs)
        ref = V.conj().T @ (H @ V)
        np.testing.assert_allclose(M, ref, rtol=1e-11, atol=1e-11 * np.abs(ref).max())
        M1 = CudaVector.matrixRepresentation(H, vs[:-1])
        S1 = CudaVector.overlapMatrix(vs[:-1])
        np.testing.assert_allclose(CudaVector.extendMatrixRepresentation(H, vs, M1), ref, atol=1e-9 * np.abs(ref).max())
        S2, M2 = CudaVector.extendBoth(H, vs, S1, M1)
        np.testing.assert_allclose(M2, ref, atol=1e-9 * np.abs(ref).max())
        np.testing.assert_allclose(S2, V.conj().T @ V, atol=1e-9)


@pytest.mark.parametrize("sigma", [511.5, 511.5 + 0.8j, -3.7])
def test_solve_with_complex_operator(rt, sigma):
    """(sigma - H) x = b with the device GCROT (numpyVector.py:147-178) against a direct dense solve and
    SciPy's GCROT on the same system: a shift inside a spectral gap (indefinite), the same with an
    imaginary part (FEAST's case) and one below the spectrum (definite); 190-290 applications each."""
    from eigensolvers_b200 import CudaVector
    n = 1500
    d = np.arange(1, n + 1, dtype=np.float64)
    d[411:] += 200.0                    # a gap of +-100 around 511.5
    H = _hermitian_sparse(n, 0.004, 31, diag=d)
    rng = np.random.default_rng(7)
    dense = sigma * np.eye(n) - np.asarray(H.todense())
    tol = 1e-10
    for b in (rng.standard_normal(n), rng.standard_normal(n) + 1j * rng.standard_normal(n)):
        b = b / np.linalg.norm(b)
        X = CudaVector.solve(H, CudaVector(b, _opts(tol, atol=1e-12)), sigma)
        assert X.dtype == np.complex128 and rt.last_solve.info == 0
        x = X.array
        x_direct = np.linalg.solve(dense, b)
        r = b - (sigma * x - H @ x)
        assert np.linalg.norm(r) <= 2e-10, np.linalg.norm(r)      # stops on ||r|| <= max(atol, tol ||b||)
        assert np.linalg.norm(x - x_direct) / np.linalg.norm(x_direct) <= 1e-7
        lin = spla.LinearOperator((n, n), matvec=lambda v: sigma * v - H @ v, dtype=np.complex128)
        x_ref, info = spla.gcrotmk(lin, b.astype(np.complex128), None, rtol=tol, atol=1e-12, maxiter=1000)
        assert info == 0
        assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) <= 1e-7
        assert rt.last_solve.n_matvec <= 600                      # SciPy: 190-290
        Xr = CudaVector.solve(H, CudaVector(b, _opts(tol, atol=1e-12)), sigma, reverseGF=True)
        np.testing.assert_allclose(Xr.array, -x, rtol=0, atol=1e-7 * np.linalg.norm(x))


def test_minres_rejects_complex_systems(rt):
    """The device MINRES is the real symmetric recurrence (minres.py:98-379); a complex-valued H or
    complex vectors are refused loudly rather than solved wrongly -- GCROT is the solver for those."""
    from eigensolvers_b200 import CudaVector
    H = _hermitian_sparse(400, 0.01, 5)
    b = np.random.default_rng(1).standard_normal(400)
    with pytest.raises(RuntimeError, match="real symmetric"):
        CudaVector.solve(H, CudaVector(b, _opts(1e-8, "minres")), -2.0)


@pytest.mark.xfail(strict=False, reason="row-sharded complex-valued H was written after this round's GPU minutes "
                                        "ran out: the single-GPU tests above ran on a B200, this one has not run yet")
def test_complex_operator_row_sharded(rt):
    """Two ranks (tests/multirank_worker.py, case zherm): the imaginary value stream follows the
    general halo plan -- applyOp on real / complex vectors, GCROT, matrixRepresentation.
    Non-strict xfail until it has been seen green on hardware (an XPASS in the report means it is)."""
    from test_gpu_multirank import _spawn
    rc, out = _spawn(2, ("zherm",), timeout=150)
    assert rc == 0, out[-6000:]
    assert out.count("PASS (all ranks: PASS") == 2, out[-6000:]
