"""Synthetic sparse Hermitian Hamiltonians of the shapes BASELINE.json names, built
on the host as scipy CSR (float64 values, sorted int32 column indices):

  C1  prescribed-spectrum dense matrix  A = Q^T diag(ev) Q   (examples/driver_numpyVector.py:27-39)
  C2  3-D 7-point Laplacian (Dirichlet) + random diagonal potential, N = n^3
  C3  coupled harmonic oscillators in the number (Hermite) product basis with chain couplings
      c q_i q_{i+1}, q = (a + a^dagger)/sqrt(2)   (unittests/test_lanczosBlockTTNS.py:21-35 is
      the reference's own two-mode instance of this family, stateFollowingHO its 1-D one)
"""
import numpy as np
import scipy.sparse as sp


def prescribed_spectrum(n=100, ev_max=300.0, seed=10):
    """examples/driver_numpyVector.py:27-39: returns (A, ev, Y0) with the legacy global RNG."""
    import scipy.linalg as la
    ev = np.linspace(1, ev_max, n)
    np.random.seed(seed)
    Q = la.qr(np.random.rand(n, n))[0]
    A = Q.T @ np.diag(ev) @ Q
    Y0 = np.random.random(n)
    return A, ev, Y0


def laplacian3d(n, seed=2, W=1.0):
    """H = T(x)I(x)I + I(x)T(x)I + I(x)I(x)T + diag(V),  T = tridiag(-1,2,-1), V = W*rng.random(N)."""
    T = sp.diags([-np.ones(n - 1), 2.0 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1], format="csr")
    I = sp.identity(n, format="csr")
    H = sp.kron(sp.kron(T, I), I) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(I, I), T)
    rng = np.random.default_rng(seed)
    V = W * rng.random(n ** 3)
    H = (H + sp.diags(V)).tocsr()
    H.sort_indices()
    H.indices = H.indices.astype(np.int32)
    H.indptr = H.indptr.astype(np.int64 if H.nnz >= 2 ** 31 else np.int32)
    return H


def oscillator_frequencies(D, seed=1):
    rng = np.random.default_rng(seed)
    return 1.0 + 0.37 * np.arange(D) / D + 0.05 * rng.random(D)


def coupled_oscillators(dims, coupling=0.1, seed=1, dtype_index=np.int32, rows=None):
    """H = sum_i w_i (n_i + 1/2) + coupling * sum_i q_i q_{i+1} in the product number basis.

    Last mode is the fastest index.  Assembled directly in CSR order: the column offsets
    {0, +-stride_i +- stride_{i+1}} are the same for every row, so visiting them in ascending
    order yields sorted rows without a COO sort (N = 2e7 needs ~25 passes over N-vectors).
    `rows=(r0, r1)` builds only that row block (global column indices), which is what a rank of
    the row-sharded mode needs; the result equals H[r0:r1] of the full matrix.
    """
    dims = [int(d) for d in dims]
    D = len(dims)
    N = int(np.prod(dims))
    r0, r1 = (0, N) if rows is None else (int(rows[0]), int(rows[1]))
    nloc = r1 - r0
    omega = oscillator_frequencies(D, seed)
    strides = [int(np.prod(dims[i + 1:])) for i in range(D)]
    idx = np.arange(r0, r1, dtype=np.int64)
    occ = [((idx // strides[i]) % dims[i]).astype(np.int16) for i in range(D)]
    diag = np.zeros(nloc)
    for i in range(D):
        diag += omega[i] * (occ[i] + 0.5)

    # off-diagonal terms: (offset, mask, value) for every pair and sign combination
    terms = [(0, None, diag)]
    for i in range(D - 1):
        j = i + 1
        for si in (-1, 1):
            for sj in (-1, 1):
                off = si * strides[i] + sj * strides[j]
                ni, nj = occ[i], occ[j]
                mask = np.ones(nloc, dtype=bool)
                mask &= (ni + si >= 0) & (ni + si < dims[i])
                mask &= (nj + sj >= 0) & (nj + sj < dims[j])
                # <n+1|q|n> = sqrt((n+1)/2), <n-1|q|n> = sqrt(n/2)
                fi = np.sqrt(((ni + 1) if si > 0 else ni) / 2.0)
                fj = np.sqrt(((nj + 1) if sj > 0 else nj) / 2.0)
                terms.append((off, mask, coupling * fi * fj))
    terms.sort(key=lambda t: t[0])
    counts = np.zeros(nloc, dtype=np.int64)
    for off, mask, _ in terms:
        counts += 1 if mask is None else mask
    indptr = np.zeros(nloc + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    nnz = int(indptr[-1])
    indices = np.empty(nnz, dtype=dtype_index)
    data = np.empty(nnz, dtype=np.float64)
    pos = indptr[:-1].copy()
    for off, mask, val in terms:
        if mask is None:
            indices[pos] = idx
            data[pos] = val
            pos += 1
        else:
            sel = np.nonzero(mask)[0]
            p = pos[sel]
            indices[p] = idx[sel] + off
            data[p] = val[sel]
            pos[sel] += 1
    H = sp.csr_matrix((data, indices, indptr), shape=(nloc, N))
    H.has_sorted_indices = True
    return H, omega


def oscillator_levels(omega, coupling, n_levels=64, max_quanta=12):
    """Analytic levels of the UNtruncated coupled-oscillator Hamiltonian (normal modes):
    E = sum_k Omega_k (n_k + 1/2),  Omega^2 = eig(W^(1/2) (W + C) W^(1/2)),  W = diag(omega)
    (probe-verified).  Returns the lowest `n_levels` energies, sorted."""
    import itertools
    D = len(omega)
    Cm = np.zeros((D, D))
    for i in range(D - 1):
        Cm[i, i + 1] = Cm[i + 1, i] = coupling
    Wh = np.diag(np.sqrt(omega))
    Om = np.sqrt(np.linalg.eigvalsh(Wh @ (np.diag(omega) + Cm) @ Wh))
    zero = 0.5 * Om.sum()
    levels = []
    for quanta in itertools.product(range(max_quanta + 1), repeat=D):
        if sum(quanta) <= max_quanta:
            levels.append(zero + float(np.dot(Om, quanta)))
    levels = np.sort(np.array(levels))
    return levels[:n_levels]


def orthonormal_block(N, nBlock, seed=3):
    """nBlock orthonormal random guesses: qr(rng.standard_normal((N, nBlock)))."""
    rng = np.random.default_rng(seed)
    Y = rng.standard_normal((N, nBlock))
    Q, _ = np.linalg.qr(Y)
    return [np.ascontiguousarray(Q[:, i]) for i in range(nBlock)]
