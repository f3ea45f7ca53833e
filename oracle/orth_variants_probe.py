"""TEST INFRASTRUCTURE (CPU only; see oracle/__init__.py) — numerical probe of Gram-Schmidt variants inside
GCROT(m,k) on a hard interior shift: SciPy-style MGS, CGS2, single-pass CGS and single-pass CGS
with Gram correction h = (2I - B^H B) h1.  Usage:
    python oracle/orth_variants_probe.py [lattice size, default 20]
Prints matvec counts, true residuals and the worst loss of orthogonality of [C,V] per variant; the
right-hand side is the third vector of a Lanczos-like sequence (the kind of solve on which the
single-pass variant degraded at N = 1e6 on the GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eigensolvers_b200 import hamiltonians as hm
from eigensolvers_b200.hostmath import calculateTarget

def gcrot(mv, b, rtol, maxiter, m=20, k=20, orth="mgs", eta=0.1):
    n = len(b); x = np.zeros(n); r = b.copy(); bn = np.linalg.norm(b); nmv = 0
    CU = []; atol = rtol * bn; maxloss = 0.0
    for outer in range(maxiter):
        beta = np.linalg.norm(r)
        if beta <= atol and (outer > 0 or CU):
            r = b - mv(x); nmv += 1; beta = np.linalg.norm(r)
        if beta <= atol:
            return x, 0, nmv, maxloss
        ml = m + max(k - len(CU), 0)
        cs = [c for c, u in CU]
        vs = [r / beta]
        B = np.zeros((len(cs), ml)); Hm = np.zeros((ml + 1, ml))
        basis = np.array(cs + vs)            # rows
        G = basis @ basis.T if orth == "gram" else None
        jdone = 0
        for j in range(ml):
            w = mv(vs[-1]); nmv += 1
            basis = np.array(cs + vs)
            if orth == "mgs":
                h = np.zeros(len(basis))
                for i, q in enumerate(basis):
                    h[i] = q @ w; w = w - h[i] * q
                nrm = np.linalg.norm(w)
            elif orth == "cgs2":
                h1 = basis @ w; w = w - h1 @ basis; h2 = basis @ w; w = w - h2 @ basis; h = h1 + h2
                nrm = np.linalg.norm(w)
            elif orth == "cgs1":
                h = basis @ w; w = w - h @ basis; nrm = np.linalg.norm(w)
            elif orth == "gram":
                h1 = basis @ w
                h = h1 + (h1 - G @ h1)          # (2I - G) h1
                ww = w @ w
                w = w - h @ basis
                nrm = np.linalg.norm(w)         # explicit here (device: lagged check)
            w = w / nrm
            if orth == "gram":
                g = basis @ w                   # Gram column of the new vector (device: in the next sweep)
                G = np.block([[G, g[:, None]], [g[None, :], np.array([[w @ w]])]])
            vs.append(w)
            B[:, j] = h[:len(cs)]; Hm[:j + 1, j] = h[len(cs):]; Hm[j + 1, j] = nrm
            jdone = j + 1
            # residual estimate via least squares (small)
            e1 = np.zeros(jdone + 1); e1[0] = 1.0
            y, res, *_ = np.linalg.lstsq(Hm[:jdone + 1, :jdone], e1, rcond=None)
            resn = np.linalg.norm(e1 - Hm[:jdone + 1, :jdone] @ y)
            if resn < atol / beta:
                break
        allb = np.array(cs + vs)
        loss = np.abs(allb @ allb.T - np.eye(len(allb))).max(); maxloss = max(maxloss, loss)
        y = y * beta
        ux = sum(vs[i] * y[i] for i in range(jdone))
        by = B[:, :jdone] @ y
        for (c, u), byc in zip(CU, by):
            ux = ux - u * byc
        hy = Hm[:jdone + 1, :jdone] @ y
        cx = sum(vs[i] * hy[i] for i in range(jdone + 1))
        al = 1 / np.linalg.norm(cx); cx *= al; ux *= al
        gamma = cx @ r; r = r - gamma * cx; x = x + gamma * ux
        while len(CU) >= k and CU: del CU[0]
        CU.append((cx, ux))
    return x, 1, nmv, maxloss

nn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
H = hm.laplacian3d(nn, seed=2, W=1.0)
ev = np.linalg.eigvalsh(H.toarray()) if nn <= 16 else None
import scipy.sparse.linalg as spla
if ev is None:
    ev = np.sort(spla.eigsh(H, k=24, which="SA")[0])
sigma = calculateTarget(ev, 10)
print("N", H.shape[0], "sigma", sigma, "gaps", np.diff(ev[8:13]))
b = np.random.default_rng(3).standard_normal(H.shape[0]); b /= np.linalg.norm(b)
mv = lambda v: sigma * v - H @ v
x1, _, _, _ = gcrot(mv, b, 1e-4, 3000, orth="cgs2")
b2 = x1 - (x1 @ b) * b
b2 /= np.linalg.norm(b2)
x2, _, _, _ = gcrot(mv, b2, 1e-4, 3000, orth="cgs2")
b3 = x2 - (x2 @ b) * b; b3 -= (b3 @ b2) * b2; b3 /= np.linalg.norm(b3)
b = b3
for orth in ("mgs", "cgs2", "cgs1", "gram"):
    t = time.time()
    x, info, nmv, loss = gcrot(mv, b, 1e-4, 3000, orth=orth)
    print(f"{orth:5s} info {info} matvecs {nmv:6d} true_res {np.linalg.norm(b - mv(x)):.3e} max_orth_loss {loss:.2e} t {time.time()-t:.1f}s", flush=True)
