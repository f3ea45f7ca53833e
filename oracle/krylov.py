"""Plain-numpy restatement of the two SciPy solvers behind NumpyVector.solve — TEST
INFRASTRUCTURE (see oracle/__init__.py).

The arithmetic of the reference's inner solve lives in a third-party dependency, SciPy
(README.md:22 recommends 1.10.1; no lockfile; this image ships 1.18.1 whose sources were read):
  gcrotmk  scipy/sparse/linalg/_isolve/_gcrotmk.py:16-183 (_fgmres), :187-506
  minres   scipy/sparse/linalg/_isolve/minres.py:13-379
Published algorithms: Hicken & Zingg, SIAM J. Sci. Comput. 32, 172 (2010) (flexible GCROT(m,k));
Paige & Saunders, SIAM J. Numer. Anal. 12, 617 (1975) (MINRES).

`gcrotmk(..., orth="mgs")` follows SciPy step for step (sequential dot/axpy projections);
`orth="cgs2"` is the variant the CUDA library runs (two classical Gram-Schmidt passes, Givens
QR of the Hessenberg matrix).  tests/test_krylov_oracle.py pins the former against SciPy itself
and bounds the difference between the two.  Both return (x, info, n_matvec).
"""
import numpy as np


def _givens_insert(Q, R, hcur, j):
    """Append column `hcur` (length j+2) to H = Q R; one rotation on rows (j, j+1)."""
    Q2 = np.zeros((j + 2, j + 2), dtype=Q.dtype)
    Q2[:j + 1, :j + 1] = Q
    Q2[j + 1, j + 1] = 1
    R2 = np.zeros((j + 2, j + 1), dtype=R.dtype)
    R2[:j + 1, :j] = R
    u = Q2.conj().T @ hcur
    a, b = u[j], u[j + 1]
    na, nb = abs(a), abs(b)
    rho = np.hypot(na, nb)
    if rho == 0 or not np.isfinite(rho):
        c, s = 1.0, 0.0
    elif na == 0:
        c, s = 0.0, np.conj(b) / nb
    else:
        c, s = na / rho, (a / na) * np.conj(b) / rho
    R2[:j, j] = u[:j]
    R2[j, j] = c * a + s * b
    qa, qb = Q2[:, j].copy(), Q2[:, j + 1].copy()
    Q2[:, j] = qa * c + qb * np.conj(s)
    Q2[:, j + 1] = -qa * s + qb * c
    return Q2, R2


def gcrotmk(matvec, b, x0=None, rtol=1e-5, atol=0.0, maxiter=1000, m=20, k=None, orth="mgs", CU=None):
    """`CU`: list of (c, u) pairs kept from an earlier solve with the SAME operator (recycling,
    _gcrotmk.py:227-236).  SciPy re-orthogonalises the c's with a pivoted QR on entry
    (:317-371); the vectors left by a previous call are already orthonormal, so this restatement
    (and the CUDA library) skips the QR and goes straight to the projection step
    x += U C^H r, r -= C C^H r (:373-388).  The list is updated in place."""
    b = np.asarray(b)
    dtype = np.result_type(b.dtype, matvec(np.zeros_like(b)).dtype, np.float64)
    b = b.astype(dtype)
    dot = np.vdot
    nmv = 0
    if k is None:
        k = m
    if x0 is None:
        x = np.zeros_like(b)
        r = b.copy()
    else:
        x = np.array(x0, dtype=dtype)
        r = b - matvec(x)
        nmv += 1
    b_norm = np.linalg.norm(b)
    atol = max(float(atol), float(rtol) * float(b_norm))
    if b_norm == 0:
        return b, 0, nmv
    eps = np.finfo(np.float64).eps
    if CU is None:
        CU = []
    for c, u in CU:                      # _gcrotmk.py:381-388
        yc = dot(c, r)
        x = x + u * yc
        r = r - c * yc
    j_outer = -1
    for j_outer in range(maxiter):
        beta = np.linalg.norm(r)
        beta_tol = max(atol, rtol * b_norm)
        if beta <= beta_tol and (j_outer > 0 or CU):
            r = b - matvec(x)
            nmv += 1
            beta = np.linalg.norm(r)
        if beta <= beta_tol:
            j_outer = -1
            break
        ml = m + max(k - len(CU), 0)
        cs = [c for c, u in CU]
        # ---- FGMRES with projection against C
        vs = [r / beta]
        inner_atol = max(atol, rtol * b_norm) / beta
        B = np.zeros((len(cs), ml), dtype=dtype)
        Q = np.ones((1, 1), dtype=dtype)
        R = np.zeros((1, 0), dtype=dtype)
        breakdown = False
        for j in range(ml):
            w = matvec(vs[-1])
            nmv += 1
            w_norm = np.linalg.norm(w)
            hcur = np.zeros(j + 2, dtype=dtype)
            if orth == "mgs":
                for i, c in enumerate(cs):
                    alpha = dot(c, w)
                    B[i, j] = alpha
                    w = w - alpha * c
                for i, v in enumerate(vs):
                    alpha = dot(v, w)
                    hcur[i] = alpha
                    w = w - alpha * v
            else:
                basis = np.array(cs + vs)
                h1 = basis.conj() @ w
                w = w - h1 @ basis
                h2 = basis.conj() @ w
                w = w - h2 @ basis
                h = h1 + h2
                B[:, j] = h[:len(cs)]
                hcur[:j + 1] = h[len(cs):]
            hcur[j + 1] = np.linalg.norm(w)
            with np.errstate(over="ignore", divide="ignore"):
                alpha = 1 / hcur[-1]
            if np.isfinite(alpha):
                w = alpha * w
            if not (hcur[-1].real > eps * w_norm):
                breakdown = True
            vs.append(w)
            Q, R = _givens_insert(Q, R, hcur, j)
            res = abs(Q[0, -1])
            if res < inner_atol or breakdown:
                break
        if not np.isfinite(R[j, j]):
            break
        y = np.zeros(j + 1, dtype=dtype)
        rhs = Q[0, :j + 1].conj()
        for i in range(j, -1, -1):
            s = rhs[i] - R[i, i + 1:j + 1] @ y[i + 1:j + 1]
            y[i] = s / R[i, i] if abs(R[i, i]) > 0 else 0.0
        B = B[:, :j + 1]
        y = y * beta
        # ---- GCROT update
        ux = sum(vs[i] * y[i] for i in range(j + 1))
        by = B @ y
        for (c, u), byc in zip(CU, by):
            ux = ux - u * byc
        hy = Q @ (R @ y)
        cx = sum(vs[i] * hy[i] for i in range(j + 2))
        with np.errstate(divide="ignore"):
            alpha = 1 / np.linalg.norm(cx)
        if not np.isfinite(alpha):
            continue
        cx = alpha * cx
        ux = alpha * ux
        gamma = dot(cx, r)
        r = r - gamma * cx
        x = x + gamma * ux
        while len(CU) >= k and CU:
            del CU[0]
        CU.append((cx, ux))
    return x, j_outer + 1, nmv


def minres(matvec, b, x0=None, rtol=1e-5, maxiter=None):
    """MINRES with shift = 0, M = I (real symmetric operator)."""
    b = np.asarray(b, dtype=np.float64)
    n = len(b)
    nmv = 0
    if maxiter is None:
        maxiter = 5 * n
    eps = np.finfo(np.float64).eps
    if x0 is None:
        x = np.zeros(n)
        r1 = b.copy()
    else:
        x = np.array(x0, dtype=np.float64)
        r1 = b - matvec(x)
        nmv += 1
    y = r1
    beta1 = np.inner(r1, y)
    if beta1 == 0:
        return x, 0, nmv
    bnorm = np.linalg.norm(b)
    if bnorm == 0:
        return b, 0, nmv
    beta1 = np.sqrt(beta1)
    oldb, beta, dbar, epsln, phibar = 0.0, beta1, 0.0, 0.0, beta1
    tnorm2, gmax, gmin, cs, sn = 0.0, 0.0, np.finfo(np.float64).max, -1.0, 0.0
    w = np.zeros(n)
    w2 = np.zeros(n)
    r2 = r1
    istop, itn = 0, 0
    while itn < maxiter:
        itn += 1
        s = 1.0 / beta
        v = s * y
        y = matvec(v)
        nmv += 1
        if itn >= 2:
            y = y - (beta / oldb) * r1
        alfa = np.inner(v, y)
        y = y - (alfa / beta) * r2
        r1 = r2
        r2 = y
        oldb = beta
        beta = np.inner(r2, y)
        if beta < 0:
            raise ValueError("non-symmetric matrix")
        beta = np.sqrt(beta)
        tnorm2 += alfa ** 2 + oldb ** 2 + beta ** 2
        if itn == 1 and beta / beta1 <= 10 * eps:
            istop = -1
        oldeps = epsln
        delta = cs * dbar + sn * alfa
        gbar = sn * dbar - cs * alfa
        epsln = sn * beta
        dbar = -cs * beta
        root = np.hypot(gbar, dbar)
        gamma = max(np.hypot(gbar, beta), eps)
        cs = gbar / gamma
        sn = beta / gamma
        phi = cs * phibar
        phibar = sn * phibar
        denom = 1.0 / gamma
        w1 = w2
        w2 = w
        w = (v - oldeps * w1 - delta * w2) * denom
        x = x + phi * w
        gmax = max(gmax, gamma)
        gmin = min(gmin, gamma)
        Anorm = np.sqrt(tnorm2)
        ynorm = np.linalg.norm(x)
        epsx = Anorm * ynorm * eps
        rnorm = phibar
        test1 = np.inf if (ynorm == 0 or Anorm == 0) else rnorm / (Anorm * ynorm)
        test2 = np.inf if Anorm == 0 else root / Anorm
        Acond = gmax / gmin
        if istop == 0:
            if 1 + test2 <= 1:
                istop = 2
            if 1 + test1 <= 1:
                istop = 1
            if itn >= maxiter:
                istop = 6
            if Acond >= 0.1 / eps:
                istop = 4
            if epsx >= beta1:
                istop = 3
            if test2 <= rtol:
                istop = 2
            if test1 <= rtol:
                istop = 1
        if istop != 0:
            break
    return x, (maxiter if istop == 6 else 0), nmv
