"""Matrix-free sum-of-products (Kronecker-sum) Hamiltonian as an `H` type for `CudaVector`
(SURVEY §8f.3):

    H = sum_s  c_s  (x)_d  h_{d,s}        on the product basis dims = (d_0, ..., d_{D-1}), last mode fastest

— the form the reference's physics Hamiltonians have before they are assembled into a matrix
(`operatornD.operatorSumOfProduct`, unittests/test_lanczosBlockTTNS.py:21-35).  Pass the object
wherever the reference takes `H` (`applyOp`, `solve`, `matrixRepresentation`, the drivers): the
fused shifted product is evaluated from the 1-D factors inside the kernel
(csrc/kernels_kron.cuh), nothing of the N x N matrix is stored or streamed.

    terms = [(coef, {mode: 1-D matrix (d_mode x d_mode), ...}), ...]      one or two factors per term
    H = KroneckerSumOperator(dims, terms)

`to_csr()` assembles the same operator with scipy (parity tests, small sizes);
`KroneckerSumOperator.coupled_oscillators(dims, ...)` is the matrix-free twin of
`hamiltonians.coupled_oscillators` (BASELINE config 3's Hamiltonian).
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib
from .operator import DeviceOperator

_MAX_TAB, _MAX_DTAB, _MAX_TERMS, _MAX_DIM = 1536, 512, 40, 8


def assemble_csr(dims, terms):
    """sum_s c_s (x)_d h_{d,s} assembled with scipy.sparse.kron (host only; parity checks at small N)."""
    dims = [int(d) for d in dims]
    N = int(np.prod(dims))
    H = sp.csr_matrix((N, N))
    for coef, factors in terms:
        mats = [sp.csr_matrix(np.asarray(factors[m], dtype=np.float64)) if m in factors else sp.identity(d, format="csr")
                for m, d in enumerate(dims)]
        K = mats[0]
        for M in mats[1:]:
            K = sp.kron(K, M, format="csr")
        H = H + float(coef) * K
    H = H.tocsr()
    H.sum_duplicates()
    H.eliminate_zeros()
    H.sort_indices()
    return H


def oscillator_terms(dims, coupling=0.1, seed=1):
    """(terms, omega) of H = sum_i w_i (n_i + 1/2) + coupling * sum_i q_i q_{i+1} in the truncated
    number basis — the factors behind hamiltonians.coupled_oscillators."""
    from .hamiltonians import oscillator_frequencies
    dims = [int(d) for d in dims]
    omega = oscillator_frequencies(len(dims), seed)
    terms, qs = [], []
    for i, d in enumerate(dims):
        terms.append((float(omega[i]), {i: np.diag(np.arange(d) + 0.5)}))
        q = np.zeros((d, d))
        for n in range(d - 1):                       # <n+1|q|n> = sqrt((n+1)/2)
            q[n + 1, n] = q[n, n + 1] = np.sqrt((n + 1) / 2.0)
        qs.append(q)
    for i in range(len(dims) - 1):
        terms.append((float(coupling), {i: qs[i], i + 1: qs[i + 1]}))
    return terms, omega


class KroneckerSumOperator(DeviceOperator):
    _is_device_operator = True

    def __init__(self, dims, terms, runtime=None):
        from .runtime import Runtime
        rt = runtime or Runtime.get()
        dims = [int(d) for d in dims]
        D = len(dims)
        N = int(np.prod(dims))
        super().__init__(rt, (N, N))
        if not (1 <= D <= _MAX_DIM) or max(dims) > 256 or N >= 2 ** 31:
            raise ValueError(f"product basis {dims}: at most {_MAX_DIM} modes of at most 256 states, N < 2^31")
        self.dims = dims
        self.terms = [(float(c), {int(m): np.asarray(h, dtype=np.float64) for m, h in f.items()}) for c, f in terms]
        strides = [int(np.prod(dims[i + 1:])) for i in range(D)]

        # -- host tables ------------------------------------------------------------------------
        dtab_off = np.zeros(_MAX_DIM, dtype=np.int32)
        dtab_off[:D] = np.concatenate(([0], np.cumsum(dims)[:-1]))
        dtab = np.zeros(int(np.sum(dims)), dtype=np.float64)
        tab_val, tab_col, cache, desc, coefs = [], [], {}, [], []
        nnz_equiv, max_off = N, 0

        def ell(mode, h):
            """ELL table (row-wise, width = max non-zeros per row) of one factor; deduplicated."""
            key = (mode, h.tobytes())
            if key in cache:
                return cache[key]
            d = dims[mode]
            if h.shape != (d, d):
                raise ValueError(f"factor of mode {mode} has shape {h.shape}, expected {(d, d)}")
            nzr = [np.nonzero(h[n])[0] for n in range(d)]
            w = max(1, max(len(c) for c in nzr))
            first = sum(len(v) for v in tab_val)
            vals, offs = np.zeros((d, w)), np.zeros((d, w), dtype=np.int64)
            for n in range(d):
                vals[n, :len(nzr[n])] = h[n, nzr[n]]
                offs[n, :len(nzr[n])] = (nzr[n] - n) * strides[mode]   # element offset of the gathered entry
            tab_val.append(vals.reshape(-1))
            tab_col.append(offs.reshape(-1).astype(np.int32))
            reach = max((int(np.max(np.abs(nzr[n] - n))) if len(nzr[n]) else 0) for n in range(d))
            cache[key] = (first, w, reach, int(sum(len(c) for c in nzr)))
            return cache[key]

        for coef, factors in self.terms:
            modes = sorted(factors)
            if not 1 <= len(modes) <= 2:
                raise NotImplementedError("terms with one or two non-identity factors are supported")
            if len(modes) == 1 and np.count_nonzero(factors[modes[0]] - np.diag(np.diag(factors[modes[0]]))) == 0:
                m = modes[0]
                dtab[dtab_off[m]:dtab_off[m] + dims[m]] += coef * np.diag(factors[m])   # merged diagonal
                continue
            ta = ell(modes[0], factors[modes[0]])
            if len(modes) == 1:
                desc.append([modes[0], -1, ta[0], 0, ta[1], 1, 0])
                max_off = max(max_off, ta[2] * strides[modes[0]])
                nnz_equiv += ta[3] * (N // dims[modes[0]])
            else:
                tb = ell(modes[1], factors[modes[1]])
                desc.append([modes[0], modes[1], ta[0], tb[0], ta[1], tb[1], 0])
                max_off = max(max_off, ta[2] * strides[modes[0]] + tb[2] * strides[modes[1]])
                nnz_equiv += ta[3] * tb[3] * (N // (dims[modes[0]] * dims[modes[1]]))
            coefs.append(coef)
        tab_val = np.concatenate(tab_val) if tab_val else np.zeros(1)
        tab_col = np.concatenate(tab_col) if tab_col else np.zeros(1, dtype=np.int32)
        if len(desc) > _MAX_TERMS or len(tab_val) > _MAX_TAB or len(dtab) > _MAX_DTAB:
            raise NotImplementedError(f"{len(desc)} product terms / {len(tab_val)} table entries exceed the kernel's limits "
                                      f"({_MAX_TERMS} / {_MAX_TAB}); assemble the operator as a sparse matrix instead")
        self.nnz = int(nnz_equiv)
        self.max_offset = int(max_off)

        # -- device side ------------------------------------------------------------------------
        r0, r1 = (0, N) if rt.world == 1 else rt.local_range(N)
        self.n_local, self.row0 = r1 - r0, r0
        d_val, d_col, d_dtab = rt.upload(tab_val), rt.upload(tab_col.astype(np.int32)), rt.upload(dtab)
        self._keep += [d_val, d_col, d_dtab]
        dims_arr = np.ascontiguousarray(dims, dtype=np.int32)
        desc_arr = np.ascontiguousarray(desc, dtype=np.int32).reshape(-1) if desc else np.zeros(7, dtype=np.int32)
        coef_arr = np.ascontiguousarray(coefs, dtype=np.float64) if coefs else np.zeros(1)
        _lib.check(rt.lib.cv_op_create_kron(rt.ctx, self.n_local, r0, D, dims_arr.ctypes.data, len(desc),
                                            desc_arr.ctypes.data, coef_arr.ctypes.data, d_val.data_ptr(), d_col.data_ptr(),
                                            len(tab_val), d_dtab.data_ptr(), dtab_off.ctypes.data, len(dtab),
                                            self.max_offset, self.nnz, C.byref(self.handle)))
        self.format = "kron"
        self.padded_nnz = 0
        # diagonal of H over this rank's rows (Jacobi preconditioner): sum over the terms of the products of
        # the factors' diagonal entries at the row's digits
        rows = np.arange(r0, r1, dtype=np.int64)
        digit = [((rows // strides[i]) % dims[i]) for i in range(D)]
        diag = np.zeros(r1 - r0)
        for coef, factors in self.terms:
            prod = np.full(r1 - r0, float(coef))
            for mode, h in factors.items():
                prod = prod * np.diag(h)[digit[mode]]
            diag += prod
        self._diag_host = diag
        if rt.world > 1:
            self._setup_band_halo(self.max_offset, self.max_offset)

    # the N x N matrix is never stored: bytes a launch must move are x and y only
    def algorithmic_bytes(self, cplx=False):
        return (32 if cplx else 16) * self.n_local

    def set_format(self, fmt):
        if fmt != "kron":
            raise NotImplementedError("a matrix-free operator has no stored format")

    # ---------------------------------------------------------------------------------------
    def to_csr(self):
        """The same operator assembled with scipy.sparse (host; parity tests at small N)."""
        return assemble_csr(self.dims, self.terms)

    @classmethod
    def coupled_oscillators(cls, dims, coupling=0.1, seed=1, runtime=None):
        """The matrix-free twin of hamiltonians.coupled_oscillators (same frequencies, same truncated
        number basis)."""
        terms, omega = oscillator_terms(dims, coupling, seed)
        op = cls(dims, terms, runtime=runtime)
        op.omega = omega
        return op
