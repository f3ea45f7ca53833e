"""What the UNMODIFIED reference does outside the real symmetric case -- the evidence behind
DESIGN §7's statement that complex-valued H is a vector-level capability only.  Runs the files under
baseline/_ref (installed by baseline/install_reference.py from the read-only checkout); skipped
where they are absent."""
import warnings

import numpy as np
import pytest

from eigensolvers_b200 import hamiltonians as hm, refdrivers

pytestmark = pytest.mark.skipif(not refdrivers.available(), reason="reference not installed under baseline/_ref")


def _opts(tol=1e-4):
    return {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": tol}}


def _complex_hermitian_twin(seed):
    """D A D^H with random unit phases D: Hermitian, complex, the spectrum of the reference's
    prescribed-spectrum test matrix (examples/driver_numpyVector.py:27-39)."""
    A, ev, _ = hm.prescribed_spectrum(100)
    rng = np.random.default_rng(seed)
    ph = np.exp(1j * rng.uniform(0, 2 * np.pi, 100))
    return (ph[:, None] * A) * ph.conj()[None, :], ev, rng


def test_numpy_vector_level_calls_accept_a_complex_hermitian_matrix():
    """applyOp / solve / matrixRepresentation of the reference's NumpyVector work on complex H
    (numpyVector.py:98-100, 147-178, 180-190): the behaviour CudaVector matches on the GPU
    (tests/test_gpu_zz_complex_operator.py)."""
    ns = refdrivers.load(register_cuda=False, numpy_backend=True)
    NV = ns.numpyVector.NumpyVector
    Hc, ev, rng = _complex_hermitian_twin(3)
    x = rng.standard_normal(100)
    y = NV(x.copy(), _opts()).applyOp(Hc)
    assert y.dtype == np.complex128
    np.testing.assert_allclose(y.array, Hc @ x, rtol=1e-14)
    sol = NV.solve(Hc, NV(x.copy() + 0j, _opts(1e-10)), 30.0)
    r = x - (30.0 * sol.array - Hc @ sol.array)
    assert np.linalg.norm(r) <= 1e-4 * 1.01                      # linear_atol default 1e-4 (numpyVector.py:31-36)
    V = np.linalg.qr(rng.standard_normal((100, 4)) + 1j * rng.standard_normal((100, 4)))[0]
    M = NV.matrixRepresentation(Hc, [NV(V[:, i].copy(), _opts()) for i in range(4)])
    np.testing.assert_allclose(M, V.conj().T @ Hc @ V, atol=1e-10)


def test_reference_lanczos_driver_breaks_on_a_complex_hermitian_matrix():
    """inexact_Lanczos.py on the same matrix: the unconjugated Gram-Schmidt (numpyVector.py:133-145)
    sends the run into the linear-dependency branch, whose checkpoint line needs a `.ttns` attribute
    NumpyVector does not have (inexact_Lanczos.py:392) -- or, with checkpoints off, returns NaN."""
    ns = refdrivers.load(register_cuda=False, numpy_backend=True)
    NV = ns.numpyVector.NumpyVector
    Hc, ev, rng = _complex_hermitian_twin(3)
    y0 = rng.standard_normal(100) + 1j * rng.standard_normal(100)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(AttributeError, match="ttns"):
            ns.inexactLanczosDiagonalization(Hc, NV(y0.copy(), _opts()), 30.0, 6, 6, 1e-10, writeOut=False)
        lam, _, st = ns.inexactLanczosDiagonalization(Hc, NV(y0.copy(), _opts()), 30.0, 6, 6, 1e-10, writeOut=False,
                                                      saveTNSsEachIteration=False)
    warnings.resetwarnings()
    assert not st["isConverged"] and np.all(np.isnan(lam))


def test_reference_feast_half_contour_needs_a_real_symmetric_matrix():
    """feast.py integrates over the upper half of the contour and doubles the real part
    (feast.py:185-201): exact for real symmetric A, not for a complex Hermitian one -- the eigenvalues
    inside the window are not found."""
    ns = refdrivers.load(register_cuda=False, numpy_backend=True)
    NV = ns.numpyVector.NumpyVector
    Hc, ev, rng = _complex_hermitian_twin(4)
    inside = np.sort(ev[(ev > 25.0) & (ev < 35.0)])
    guess = [NV(rng.standard_normal(100) + 1j * rng.standard_normal(100), _opts(1e-8)) for _ in range(len(inside) + 2)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lam, _, st = ns.feastDiagonalization(Hc, guess, 8, "legendre", 25.0, 35.0, 1e-9, 6, writeOut=False)
    warnings.resetwarnings()
    lam = np.sort(np.real(lam))
    got = lam[(lam > 25.0) & (lam < 35.0)]
    assert not st["isConverged"]
    assert len(got) != len(inside) or np.max(np.abs(got - inside)) > 1e-3
