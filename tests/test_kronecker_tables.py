"""Host side of the matrix-free Kronecker-sum operator, checked without a GPU.

`KroneckerSumOperator.__init__` (eigensolvers_b200/kronecker.py) turns the 1-D factors into the
tables `k_spmv_kron` reads (csrc/kernels_kron.cuh: merged per-mode diagonal table, ELL tables of
(value, element offset), seven-integer term descriptors) and registers them through
cv_op_create_kron, which is a pure host routine.  Here the constructor runs against a stand-in
runtime that keeps the "uploaded" arrays on the host and records the C call; a numpy restatement of
the kernel's per-row arithmetic -- mixed-radix digits, diag = sum_d dtab[..], entries at
row + off_a (+ off_b) with value coef * val_a (* val_b) -- applied to exactly those tables must
reproduce scipy's assembled operator (unittests/test_lanczosBlockTTNS.py:21-35 builds its H the
same way, from 1-D factors).  The device kernel itself is covered by tests/test_gpu_kernels.py.
"""
import ctypes as C

import numpy as np
import pytest

from eigensolvers_b200 import _lib
from eigensolvers_b200.kronecker import KroneckerSumOperator, assemble_csr, oscillator_terms


class _HostArray:
    def __init__(self, a):
        self.a = np.ascontiguousarray(a)

    def data_ptr(self):
        return self.a.ctypes.data


class _RecordingLib:
    """Forwards to the real library and keeps the arguments of cv_op_create_kron."""

    def __init__(self):
        self.real = _lib.load()
        self.call = None

    def cv_op_create_kron(self, *args):
        self.call = args
        return self.real.cv_op_create_kron(*args)

    def __getattr__(self, name):
        return getattr(self.real, name)


class _HostRuntime:
    """What the constructor needs from a Runtime, minus the device: one rank, host 'uploads'."""
    world, rank = 1, 0

    def __init__(self):
        self.lib = _RecordingLib()
        self.ctx = C.c_void_p(0x10)          # cv_op_create_kron only checks it is non-null
        self.uploads = []

    def upload(self, a):
        self.uploads.append(_HostArray(a))
        return self.uploads[-1]


def _build(dims, terms):
    rt = _HostRuntime()
    op = KroneckerSumOperator(dims, terms, runtime=rt)
    (ctx, n_rows, row0, ndim, dims_p, nterm, desc_p, coef_p, val_p, col_p, tab_len, dtab_p, dtab_off_p, dtab_len,
     max_off, nnz_equiv, out) = rt.lib.call
    tab_val, tab_col, dtab = (u.a for u in rt.uploads[:3])
    assert (val_p, col_p, dtab_p) == (rt.uploads[0].data_ptr(), rt.uploads[1].data_ptr(), rt.uploads[2].data_ptr())
    assert tab_col.dtype == np.int32 and tab_val.dtype == np.float64 and dtab.dtype == np.float64
    assert (tab_len, dtab_len) == (len(tab_val), len(dtab)) and ndim == len(dims)

    def ints(ptr, n):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int32)), shape=(n,)).copy()
    desc = ints(desc_p, 7 * max(nterm, 1)).reshape(-1, 7)[:nterm]
    coef = np.ctypeslib.as_array(C.cast(coef_p, C.POINTER(C.c_double)), shape=(max(nterm, 1),)).copy()[:nterm]
    return op, dict(dims=ints(dims_p, ndim), dtab_off=ints(dtab_off_p, 8), desc=desc, coef=coef, tab_val=tab_val,
                    tab_col=tab_col, dtab=dtab, n_rows=n_rows, row0=row0, max_off=max_off, nnz_equiv=nnz_equiv)


def _emulate(t, x):
    """y = H x from the tables, the way k_spmv_kron computes a row (kernels_kron.cuh)."""
    dims = [int(d) for d in t["dims"]]
    N = int(np.prod(dims))
    stride = [int(np.prod(dims[i + 1:])) for i in range(len(dims))]
    rows = np.arange(N, dtype=np.int64)
    digit = [(rows // stride[d]) % dims[d] for d in range(len(dims))]          # last mode fastest
    diag = np.zeros(N)
    for d in range(len(dims)):
        diag += t["dtab"][t["dtab_off"][d] + digit[d]]
    y = diag * x
    for (ma, mb, ta, tb, wa, wb, _), c in zip(t["desc"], t["coef"]):
        ra = ta + digit[ma] * wa
        for ja in range(wa):
            va, oa = c * t["tab_val"][ra + ja], t["tab_col"][ra + ja].astype(np.int64)
            if mb < 0:
                idx = rows + oa
                assert np.all((idx >= 0) & (idx < N))                                # no clamp in the kernel
                y += va * x[idx]
                continue
            rb = tb + digit[mb] * wb
            for jb in range(wb):
                idx = rows + oa + t["tab_col"][rb + jb]
                assert np.all((idx >= 0) & (idx < N))
                y += va * t["tab_val"][rb + jb] * x[idx]
    return y


def _random_terms(dims, rng, n_terms):
    terms = []
    for _ in range(n_terms):
        kind = rng.integers(0, 4)
        m = int(rng.integers(0, len(dims)))
        d = dims[m]
        if kind == 0:                                           # single-factor diagonal (merged into dtab)
            terms.append((float(rng.standard_normal()), {m: np.diag(rng.standard_normal(d))}))
        elif kind == 1:                                         # single-factor banded, not symmetric
            h = np.triu(np.tril(rng.standard_normal((d, d)), 1), -2)
            terms.append((float(rng.standard_normal()), {m: h}))
        else:                                                   # two factors, ragged rows (ELL padding)
            m2 = int(rng.choice([i for i in range(len(dims)) if i != m])) if len(dims) > 1 else None
            if m2 is None:
                continue
            ha = rng.standard_normal((d, d)) * (rng.random((d, d)) < 0.5)
            hb = rng.standard_normal((dims[m2], dims[m2])) * (rng.random((dims[m2], dims[m2])) < 0.4)
            terms.append((float(rng.standard_normal()), {m: ha, m2: hb}))
    return terms


@pytest.mark.parametrize("dims", [(6, 5, 5, 4), (4, 3, 5), (7, 2), (3,), (2, 2, 2, 2, 2, 2, 2, 2)])
def test_oscillator_tables_reproduce_the_assembled_operator(dims):
    terms, omega = oscillator_terms(dims, coupling=0.1, seed=1)
    op, t = _build(dims, terms)
    H = assemble_csr(dims, terms)
    N = H.shape[0]
    assert op.shape == (N, N) and op.format == "kron" and t["n_rows"] == N and t["row0"] == 0
    x = np.random.default_rng(0).standard_normal(N)
    np.testing.assert_allclose(_emulate(t, x), H @ x, rtol=1e-13, atol=1e-13)
    # q_i q_{i+1}: exactly 2 x 2 table entries per row (the kernel's four-gather fast path) once a mode has
    # three or more states; a two-state mode has one entry per row
    assert all((wa, wb) == (min(2, dims[ma] - 1), min(2, dims[mb] - 1)) for (ma, mb, _, _, wa, wb, _) in t["desc"] if mb >= 0)
    assert len(t["desc"]) == len(dims) - 1                      # the n + 1/2 terms live in the diagonal table
    np.testing.assert_allclose(op._diag_host, H.diagonal(), rtol=1e-14, atol=1e-14)     # Jacobi diagonal
    assert t["nnz_equiv"] >= H.nnz                              # structural count (couplings cut at the basis edge stay)
    coo = H.tocoo()
    assert t["max_off"] == np.abs(coo.col.astype(np.int64) - coo.row).max()              # halo band of the sharded mode


@pytest.mark.parametrize("seed", range(8))
def test_random_sum_of_products_tables(seed):
    rng = np.random.default_rng(seed)
    dims = tuple(int(d) for d in rng.integers(1, 6, size=rng.integers(1, 5)))
    terms = _random_terms(dims, rng, int(rng.integers(1, 7)))
    if not terms:
        terms = [(1.5, {0: np.eye(dims[0])})]
    op, t = _build(dims, terms)
    H = assemble_csr(dims, terms)
    N = H.shape[0]
    x = rng.standard_normal(N)
    ref = H @ x
    np.testing.assert_allclose(_emulate(t, x), ref, rtol=1e-12, atol=1e-12 * max(1.0, np.abs(ref).max()))
    np.testing.assert_allclose(op._diag_host, H.diagonal(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(op.to_csr().toarray(), H.toarray())
    coo = H.tocoo()
    true_reach = int(np.abs(coo.col.astype(np.int64) - coo.row).max()) if H.nnz else 0
    assert t["max_off"] >= true_reach                           # the band the neighbours push covers every gather


def test_identical_factors_share_one_table():
    dims = (5, 5, 5)
    q = np.diag(np.sqrt(np.arange(1, 5) / 2.0), 1)
    q = q + q.T
    terms = [(0.1, {0: q, 1: q}), (0.2, {1: q, 2: q}), (0.3, {0: q, 2: q})]
    op, t = _build(dims, terms)
    assert len(t["tab_val"]) == 3 * 5 * 2                       # one ELL table per (mode, matrix), not per use
    x = np.random.default_rng(1).standard_normal(125)
    np.testing.assert_allclose(_emulate(t, x), assemble_csr(dims, terms) @ x, rtol=1e-13, atol=1e-13)


def test_limits_are_refused_on_the_host():
    with pytest.raises(NotImplementedError):                    # three non-identity factors in one term
        _build((3, 3, 3), [(1.0, {0: np.ones((3, 3)), 1: np.ones((3, 3)), 2: np.ones((3, 3))})])
    with pytest.raises(ValueError):                             # factor of the wrong size
        _build((3, 4), [(1.0, {0: np.ones((4, 4)), 1: np.ones((4, 4))})])
    with pytest.raises(ValueError):                             # more modes than the kernel decodes
        _build((2,) * 9, [(1.0, {0: np.eye(2)})])
    big = np.ones((200, 200))
    with pytest.raises(NotImplementedError):                    # ELL tables beyond the shared-memory budget
        _build((200, 200), [(1.0, {0: big, 1: big})])
