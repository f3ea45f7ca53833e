"""GPU parity tests of every C-ABI entry point against the reference's own expressions
(numpyVector.py lines cited per test) on seeded inputs.  Tolerances: these are fp64 sums whose
order differs from numpy's, so elementwise agreement is asserted to a few ulps of the
accumulated magnitude (rtol 1e-13), far inside north_star's 1e-10 eigenvalue bar.
"""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

RTOL = 1e-13


def _rand(n, cplx, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(n)
    if cplx:
        a = a + 1j * rng.standard_normal(n)
    return a


def _opts():
    return {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-4}}


SIZES = [1, 2, 31, 32, 33, 1000, 4097, 262144 + 3]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("cplx", [False, True])
def test_blas1(rt, n, cplx):
    from eigensolvers_b200 import CudaVector
    x, y = _rand(n, cplx, 1), _rand(n, cplx, 2)
    X, Y = CudaVector(x, _opts()), CudaVector(y, _opts())
    assert len(X) == n and X.size == n and X.shape == (n,)
    assert X.dtype == x.dtype
    np.testing.assert_array_equal(X.array, x)                       # round trip
    np.testing.assert_allclose((X * 2.5).array, x * 2.5, rtol=1e-15)  # numpyVector.py:57
    np.testing.assert_allclose((2.5 * X).array, 2.5 * x, rtol=1e-15)  # :60
    np.testing.assert_allclose((X / 3.0).array, x / 3.0, rtol=1e-15)  # :63
    z = (X * (0.5 - 2j)).array
    np.testing.assert_allclose(z, x * (0.5 - 2j), rtol=1e-15)
    np.testing.assert_allclose(X.norm(), np.linalg.norm(x), rtol=RTOL)  # :80
    np.testing.assert_allclose(X.vdot(Y), np.vdot(x, y), rtol=1e-12, atol=1e-12 * np.sqrt(n))  # :91
    np.testing.assert_allclose(X.vdot(Y, conjugate=False), np.dot(x, y), rtol=1e-12, atol=1e-12 * np.sqrt(n))  # :93
    np.testing.assert_array_equal(X.real().array, np.real(x))       # :83
    np.testing.assert_array_equal(X.conjugate().array, x.conj())    # :86
    c = X.copy()
    c.normalize()                                                    # :76
    np.testing.assert_allclose(c.array, x / np.linalg.norm(x), rtol=RTOL)
    np.testing.assert_array_equal(X.array, x)                       # copy is independent
    assert X.compress() is X and X.maxD == 0 and X.hasExactAddition
    with pytest.raises(NotImplementedError):
        X *= 2.0


def test_mixed_dot(rt):
    from eigensolvers_b200 import CudaVector
    x, y = _rand(1000, False, 1), _rand(1000, True, 2)
    X, Y = CudaVector(x), CudaVector(y)
    np.testing.assert_allclose(X.vdot(Y), np.vdot(x, y), rtol=1e-12)
    np.testing.assert_allclose(Y.vdot(X), np.vdot(y, x), rtol=1e-12)


def test_options_default_and_shared(rt):
    """numpyVector.py:25-36: defaults are written INTO the caller's dict and shared."""
    from eigensolvers_b200 import CudaVector
    inner = {"linearSolver": "gcrotmk"}
    v = CudaVector(np.ones(4), {"linearSystemArgs": inner})
    assert inner["linearIter"] == 1000 and inner["linear_tol"] == 1e-4 and inner["linear_atol"] == 1e-4
    assert (v * 2.0).options["linearSystemArgs"] is inner
    assert CudaVector(np.ones(4)).options["linearSystemArgs"]["linearSolver"] == "minres"


@pytest.mark.parametrize("n", [5, 1000, 100003])
@pytest.mark.parametrize("m,k", [(1, 1), (2, 1), (7, 3), (24, 24), (61, 2)])
def test_lincomb(rt, n, m, k):
    from eigensolvers_b200 import CudaVector
    rng = np.random.default_rng(5)
    V = rng.standard_normal((m, n))
    Cm = rng.standard_normal((m, k))
    vs = [CudaVector(V[i]) for i in range(m)]
    out = CudaVector.linearCombinationBlock(vs, Cm)
    ref = Cm.T @ V
    for j in range(k):
        np.testing.assert_allclose(out[j].array, ref[j], rtol=1e-12, atol=1e-12)
    one = CudaVector.linearCombination(vs, list(Cm[:, 0]))          # numpyVector.py:105-119
    np.testing.assert_allclose(one.array, ref[0], rtol=1e-12, atol=1e-12)


def test_lincomb_complex_and_errors(rt):
    from eigensolvers_b200 import CudaVector
    n = 777
    a, b = _rand(n, True, 1), _rand(n, True, 2)
    A, B = CudaVector(a), CudaVector(b)
    c = [0.3 - 1j, 2.0 + 0.5j]
    np.testing.assert_allclose(CudaVector.linearCombination([A, B], c).array, c[0] * a + c[1] * b, rtol=1e-13)
    np.testing.assert_allclose(CudaVector.linearCombination([A, B], [1.0, -1.0]).array, a - b, rtol=1e-13)
    R = CudaVector(a.real.copy())
    with pytest.raises(TypeError):  # complex term on a real accumulator
        CudaVector.linearCombination([R, R], [1.0, 1j])
    with pytest.raises(AssertionError):
        CudaVector.linearCombination([R, R], [1.0])


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("m", [1, 3, 17, 40])
def test_overlap_and_tsdot(rt, cplx, m):
    from eigensolvers_b200 import CudaVector
    n = 20011
    V = np.array([_rand(n, cplx, 10 + i) for i in range(m)])
    vs = [CudaVector(V[i]) for i in range(m)]
    S = CudaVector.overlapMatrix(vs)                                  # numpyVector.py:192-203
    ref = V.conj() @ V.T
    np.testing.assert_allclose(S, ref, rtol=1e-12, atol=1e-10)
    np.testing.assert_array_equal(np.triu(S, 1), np.tril(S, -1).conj().T)  # mirrored exactly
    assert S.dtype == V.dtype


def _sym_sparse(n, density, seed):
    rng = np.random.default_rng(seed)
    A = sp.random(n, n, density=density, random_state=rng, format="csr")
    A = (A + A.T) * 0.5 + sp.diags(rng.standard_normal(n))
    return A.tocsr()


@pytest.mark.parametrize("fmt", ["csr", "sell", "dia"])
@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("case", ["lap", "rand", "dense", "osc", "empty_rows"])
def test_spmv(rt, fmt, cplx, case):
    """H@x, sigma*x-H@x, H@x-sigma*x (numpyVector.py:100,152,154) and the fused dots."""
    import ctypes as C
    from eigensolvers_b200 import CudaVector, DeviceOperator, _lib, hamiltonians as hm
    if case == "lap":
        H = hm.laplacian3d(17)
    elif case == "rand":
        H = _sym_sparse(3001, 0.004, 3)
    elif case == "dense":
        H = hm.prescribed_spectrum(100)[0]
    elif case == "osc":
        H = hm.coupled_oscillators((6, 5, 4, 4, 3))[0]
    else:
        H = sp.csr_matrix(([1.0, 2.0, 3.0], ([0, 5, 70], [3, 5, 1])), shape=(75, 75))
    if fmt == "dia" and case in ("rand", "dense"):
        with pytest.raises(ValueError):      # not banded: DIA is refused, never silently wrong
            DeviceOperator.from_host(H, fmt=fmt)
        return
    op = DeviceOperator.from_host(H, fmt=fmt)
    assert op.format == fmt
    Hd = H if not sp.issparse(H) else H
    n = H.shape[0]
    x = _rand(n, cplx, 7)
    X = CudaVector(x)
    np.testing.assert_allclose(X.applyOp(op).array, Hd @ x, rtol=1e-12, atol=1e-12)
    sigma = (0.7 + 0.3j) if cplx else 0.7
    for mode, ref in ((1, sigma * x - Hd @ x), (2, Hd @ x - sigma * x)):
        y = rt.empty(n, cplx)
        out = _lib.dbl_array(3)
        s = complex(sigma)
        _lib.check(rt.lib.cv_spmv_dots(rt.ctx, op.handle, int(cplx), mode, s.real, s.imag, X._ptr,
                                       y.data_ptr(), out, rt.stream))
        yh = y.cpu().numpy()
        np.testing.assert_allclose(yh, ref, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(complex(out[0], out[1]), np.vdot(x, ref), rtol=1e-11, atol=1e-10)
        np.testing.assert_allclose(out[2], np.vdot(ref, ref).real, rtol=1e-12)


def test_format_selection(rt):
    """auto: DIA for banded structure, SELL for short irregular rows, CSR for tiny/dense."""
    from eigensolvers_b200 import DeviceOperator, hamiltonians as hm
    assert DeviceOperator.from_host(hm.laplacian3d(17)).format == "dia"
    assert DeviceOperator.from_host(hm.coupled_oscillators((12, 10, 10))[0]).format == "dia"
    # heavy basis-edge truncation: DIA padding would cost more than the CSR index stream
    assert DeviceOperator.from_host(hm.coupled_oscillators((6, 5, 4, 4, 3))[0]).format in ("sell", "csr")
    assert DeviceOperator.from_host(_sym_sparse(3001, 0.004, 3)).format in ("sell", "csr")
    assert DeviceOperator.from_host(hm.prescribed_spectrum(100)[0]).format == "csr"
    # a banded matrix with stray entries that the host-side row sample does not see: the device fill
    # kernel checks EVERY entry, the format falls back, and the product stays right
    H = hm.laplacian3d(30).tolil()
    n = H.shape[0]
    sampled = set(np.concatenate([np.arange(0, 2048), np.arange(n - 2048, n),
                                  np.linspace(0, n - 1, num=4096, dtype=np.int64)]).tolist())
    row = next(r for r in range(n // 2, n) if r not in sampled and (r + 7777) % n not in sampled)
    col = (row + 7777) % n
    H[row, col] = 1.0
    H[col, row] = 1.0
    H = H.tocsr()
    op = DeviceOperator.from_host(H)
    assert op.format in ("sell", "csr")
    x = np.random.default_rng(0).standard_normal(n)
    from eigensolvers_b200 import CudaVector
    np.testing.assert_allclose(CudaVector(x).applyOp(op).array, H @ x, rtol=1e-12, atol=1e-12)


def test_operator_cache_and_types(rt):
    from eigensolvers_b200 import CudaVector
    from scipy.sparse.linalg import LinearOperator
    H = _sym_sparse(500, 0.01, 1)
    op1, op2 = rt.operator_for(H), rt.operator_for(H)
    assert op1 is op2
    X = CudaVector(np.ones(500))
    np.testing.assert_allclose(X.applyOp(H).array, H @ np.ones(500), rtol=1e-13, atol=1e-13)
    with pytest.raises(TypeError):
        X.applyOp(LinearOperator((500, 500), matvec=lambda v: v))
    with pytest.raises(ValueError):
        CudaVector(np.ones(3)).applyOp(H)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("m", [0, 1, 5, 23])
def test_gram_schmidt(rt, cplx, m):
    """orthogonalize_against_set (numpyVector.py:121-145): unconjugated products, /q.q, LINDEP."""
    from eigensolvers_b200 import CudaVector
    n = 5003
    x = _rand(n, cplx, 100)
    qs = [_rand(n, cplx, 200 + i) * (1.0 + 0.1 * i) for i in range(m)]   # NOT normalised on purpose
    ref = x.copy()
    for q in qs:
        ref = ref - q * (np.dot(ref, q) / np.dot(q, q))
    ip = np.dot(ref, ref)
    out = CudaVector.orthogonalize_against_set(CudaVector(x), [CudaVector(q) for q in qs])
    if not (ip > 1e-14):      # numpyVector.py:141 — for complex data numpy orders by the real part first
        assert out is None
        return
    ref = ref / np.sqrt(ip)
    np.testing.assert_allclose(out.array, ref, rtol=1e-10, atol=1e-12)


def test_gram_schmidt_lindep(rt):
    from eigensolvers_b200 import CudaVector
    n = 4000
    rng = np.random.default_rng(0)
    Q = np.linalg.qr(rng.standard_normal((n, 4)))[0]
    qs = [CudaVector(Q[:, i].copy()) for i in range(4)]
    dep = Q @ np.array([0.5, -0.5, 0.5, 0.5])                  # unit vector inside span(qs)
    assert CudaVector.orthogonalize_against_set(CudaVector(dep), qs) is None
    nearly = dep + 1e-5 * rng.standard_normal(n) / np.sqrt(n)   # |residual|^2 ~ 1e-10 > 1e-14
    assert CudaVector.orthogonalize_against_set(CudaVector(nearly), qs) is not None
    assert CudaVector.orthogonalize_against_set(CudaVector(nearly), qs, lindep=1e-6) is None


@pytest.mark.parametrize("kind", ["dense", "sparse"])
def test_matrix_representation_and_extension(rt, kind):
    """matrixRepresentation / extend* (numpyVector.py:180-238); the reference's own test of
    "extension == full rebuild" (unittests/test_lanczos.py:67-76, atol 1e-9)."""
    from eigensolvers_b200 import CudaVector, hamiltonians as hm
    H = hm.prescribed_spectrum(100)[0] if kind == "dense" else hm.laplacian3d(12)
    n = H.shape[0]
    rng = np.random.default_rng(4)
    V = np.linalg.qr(rng.standard_normal((n, 6)))[0]
    vs = [CudaVector(V[:, i].copy()) for i in range(6)]
    M = CudaVector.matrixRepresentation(H, vs)
    np.testing.assert_allclose(M, V.T @ (H @ V), rtol=1e-11, atol=1e-11)
    S = CudaVector.overlapMatrix(vs)
    M1 = CudaVector.matrixRepresentation(H, vs[:-1])
    S1 = CudaVector.overlapMatrix(vs[:-1])
    np.testing.assert_allclose(CudaVector.extendMatrixRepresentation(H, vs, M1), M, atol=1e-9)
    np.testing.assert_allclose(CudaVector.extendOverlapMatrix(vs, S1), S, atol=1e-9)
    S2, M2 = CudaVector.extendBoth(H, vs, S1, M1)
    np.testing.assert_allclose(S2, S, atol=1e-9)
    np.testing.assert_allclose(M2, M, atol=1e-9)


@pytest.mark.parametrize("cplx", [False, True])
def test_more_vectors_than_one_launch_takes(rt, cplx):
    """The reference has no cap on the length of the Krylov list (L*nBlock = 200 in
    unittests/test_lanczosLINDEP.py's set-up); a launch takes 128 pointers, so linear combinations,
    tall-skinny products, Gram-Schmidt and the extend* columns chunk internally (m = 150, 257)."""
    from eigensolvers_b200 import CudaVector
    n = 3001
    for m in (150, 257):
        V = np.stack([_rand(n, cplx, 1000 + i) for i in range(m)], axis=1) / np.sqrt(n)
        vs = [CudaVector(V[:, i].copy()) for i in range(m)]
        rng = np.random.default_rng(m)
        C = rng.standard_normal((m, 3)) + (1j * rng.standard_normal((m, 3)) if cplx else 0)
        ys = CudaVector.linearCombinationBlock(vs, C)                       # numpyVector.py:105-119
        for k in range(3):
            np.testing.assert_allclose(ys[k].array, V @ C[:, k], rtol=1e-11, atol=1e-12)
        one = CudaVector.linearCombination(vs, C[:, 0])
        np.testing.assert_allclose(one.array, V @ C[:, 0], rtol=1e-11, atol=1e-12)
        S = CudaVector.overlapMatrix(vs)                                    # :192-203
        np.testing.assert_allclose(S, V.conj().T @ V, rtol=1e-10, atol=1e-12)
        S1 = CudaVector.overlapMatrix(vs[:-1])
        np.testing.assert_allclose(CudaVector.extendOverlapMatrix(vs, S1), S, atol=1e-9)   # :223-238
        H = sp.diags([np.linspace(1, 2, n)], [0]).tocsr()
        M = CudaVector.matrixRepresentation(H, vs)
        M1 = CudaVector.matrixRepresentation(H, vs[:-1])
        np.testing.assert_allclose(CudaVector.extendMatrixRepresentation(H, vs, M1), M, atol=1e-9)  # :205-221
        x = _rand(n, cplx, 7)
        ref = x.copy()
        for i in range(m):                                                  # :121-145
            q = V[:, i]
            ref = ref - q * (np.dot(ref, q) / np.dot(q, q))
        ref = ref / np.sqrt(np.dot(ref, ref))
        out = CudaVector.orthogonalize_against_set(CudaVector(x), vs)
        np.testing.assert_allclose(out.array, ref, rtol=1e-9, atol=1e-11)


def test_operator_cache_sees_in_place_edits(rt):
    """the reference evaluates H @ x on every call, so an in-place edit of H must not be
    served from a stale device copy."""
    from eigensolvers_b200 import CudaVector
    H = _sym_sparse(600, 0.01, 3)
    x = np.arange(600.0)
    X = CudaVector(x)
    np.testing.assert_allclose(X.applyOp(H).array, H @ x, rtol=1e-13, atol=1e-10)
    H.data *= 1.5
    np.testing.assert_allclose(X.applyOp(H).array, H @ x, rtol=1e-13, atol=1e-10)
    A = np.diag(np.arange(1.0, 51.0))
    Xd = CudaVector(np.ones(50))
    np.testing.assert_allclose(Xd.applyOp(A).array, A @ np.ones(50), rtol=1e-13)
    A[3, 7] = A[7, 3] = 2.0
    rt.invalidate_operator(A)       # single-entry edits: the explicit route
    np.testing.assert_allclose(Xd.applyOp(A).array, A @ np.ones(50), rtol=1e-13)


@pytest.mark.parametrize("cplx", [False, True])
def test_kronecker_sum_operator(rt, cplx):
    """Matrix-free sum-of-products operator (unittests/test_lanczosBlockTTNS.py:21-35 is the
    reference's template for such Hamiltonians) against the same operator assembled with scipy.sparse.kron:
    diagonal and dense single-factor terms, two-factor terms with odd table widths, all three shift modes."""
    from eigensolvers_b200 import CudaVector, KroneckerSumOperator, _lib
    rng = np.random.default_rng(11)
    dims = (5, 4, 6, 3)

    def band(d, bw):
        h = rng.standard_normal((d, d))
        h = h + h.T
        for i in range(d):
            for j in range(d):
                if abs(i - j) > bw:
                    h[i, j] = 0.0
        return h
    terms = [(1.3, {0: np.diag(rng.standard_normal(5))}), (0.7, {2: np.diag(rng.standard_normal(6))}),
             (0.5, {1: band(4, 3)}),                       # dense single factor (width 4)
             (-0.8, {3: band(3, 1)}),                      # tridiagonal single factor (width 3: odd)
             (0.25, {0: band(5, 1), 1: band(4, 1)}),       # 3 x 3 entries per row
             (0.4, {1: band(4, 2), 3: band(3, 0)}),        # 4-wide x diagonal
             (-0.3, {0: band(5, 2), 3: band(3, 2)})]       # 5 x 3, non-adjacent modes
    op = KroneckerSumOperator(dims, terms)
    H = op.to_csr()
    n = H.shape[0]
    assert op.shape == (n, n) and op.format == "kron"
    x = _rand(n, cplx, 5)
    X = CudaVector(x)
    np.testing.assert_allclose(X.applyOp(op).array, H @ x, rtol=1e-13, atol=1e-13)
    y = rt.empty(n, int(cplx))
    for mode, want in ((_lib.CV_SPMV_SHIFT, 0.37 * x - H @ x), (_lib.CV_SPMV_RSHIFT, H @ x - 0.37 * x)):
        _lib.check(rt.lib.cv_spmv(rt.ctx, op.handle, int(cplx), mode, 0.37, 0.0, X._ptr, y.data_ptr(), rt.stream))
        np.testing.assert_allclose(y.cpu().numpy(), want, rtol=1e-13, atol=1e-13)
    M = CudaVector.matrixRepresentation(op, [X, CudaVector(_rand(n, cplx, 6))])
    assert M.shape == (2, 2) and abs(M[0, 0] - np.vdot(x, H @ x)) <= 1e-11 * abs(M[0, 0])


def test_kronecker_oscillator_equals_assembled_path(rt):
    """The matrix-free twin of BASELINE config 3's generator: identical to the assembled CSR->DIA path to
    1e-13 relative in the product, and the same Lanczos eigenpair through the shifted solves."""
    import warnings
    from eigensolvers_b200 import CudaVector, DeviceOperator, KroneckerSumOperator, hamiltonians as hm
    from eigensolvers_b200.hostmath import calculateTarget
    from eigensolvers_b200.lanczos import inexactLanczosDiagonalization
    dims = (12, 10, 8, 6, 5)
    H, om = hm.coupled_oscillators(dims, coupling=0.1, seed=1)
    dia = DeviceOperator.from_host(H)
    assert dia.format == "dia"
    kron = KroneckerSumOperator.coupled_oscillators(dims, coupling=0.1, seed=1)
    assert kron.nnz == H.nnz and kron.max_offset == int(np.max(np.abs(dia.dia_offsets)))
    x = _rand(H.shape[0], False, 9)
    X = CudaVector(x)
    a, b = X.applyOp(kron).array, X.applyOp(dia).array
    assert np.max(np.abs(a - b)) <= 1e-13 * np.max(np.abs(b))
    levels = hm.oscillator_levels(om, 0.1, 40, max_quanta=6)
    sigma = float(calculateTarget(levels, 8))
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 5000, "linear_tol": 1e-4, "linear_atol": 0.0}}
    out = []
    for op in (kron, dia):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out.append(inexactLanczosDiagonalization(op, CudaVector(x, dict(o)), sigma, 8, 20, 1e-10, writeOut=False))
        warnings.resetwarnings()
    (ev_k, Y_k, st_k), (ev_d, Y_d, st_d) = out
    assert st_k["isConverged"] and st_d["isConverged"]
    assert abs(ev_k[0] - ev_d[0]) <= 1e-10 * abs(ev_d[0])
    assert abs(np.vdot(Y_k[0].array, Y_d[0].array)) >= 1 - 1e-8
