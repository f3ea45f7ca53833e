"""CPU restatement of the reference's NumpyVector (numpyVector.py:23-238) on numpy/scipy.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Dense ndarray storage, the same arithmetic in
the same order as the reference: temporaries per term in linear combinations, sequential
Gram-Schmidt with unconjugated products, SciPy's gcrotmk / minres / spsolve behind `solve`.
Differences: SciPy >= 1.14 spells the relative tolerance `rtol` (the reference passes `tol=`,
numpyVector.py:161,163; same criterion), and the class derives from the stand-alone interface
so that it runs without /root/reference.
"""
import warnings

import numpy as np
import scipy.sparse.linalg as spla
from scipy import linalg as la
from scipy.sparse import csc_matrix

from eigensolvers_b200.vector_api import AbstractVector, LINDEP_DEFAULT_VALUE


class NumpyVectorOracle(AbstractVector):
    matvec_count = 0  # class-wide counter of operator applications (bench bookkeeping)

    def __init__(self, array, options=dict()):  # numpyVector.py:25-36
        self.array = array
        self.size = array.size
        self.shape = array.shape
        self.options = dict()
        opt = options.get("linearSystemArgs", dict())
        opt["linearSolver"] = opt.get("linearSolver", "minres")
        opt["linearIter"] = opt.get("linearIter", 1000)
        opt["linear_tol"] = opt.get("linear_tol", 1e-4)
        opt["linear_atol"] = opt.get("linear_atol", 1e-4)
        self.options["linearSystemArgs"] = opt

    hasExactAddition = property(lambda self: True)        # :38-46
    dtype = property(lambda self: self.array.dtype)       # :48-50
    maxD = property(lambda self: 0)                       # :52-55

    def __mul__(self, other):                             # :57-58
        return NumpyVectorOracle(self.array * other, self.options)

    __rmul__ = __mul__                                    # :60-61

    def __truediv__(self, other):                         # :63-64
        return NumpyVectorOracle(self.array / other, self.options)

    def __imul__(self, other):                            # :66-67
        raise NotImplementedError

    def __itruediv__(self, other):                        # :69-70
        raise NotImplementedError

    def __len__(self):                                    # :73-74
        return len(self.array)

    def normalize(self):                                  # :76-78
        self.array /= la.norm(self.array)
        return self

    def norm(self):                                       # :80-81
        return la.norm(self.array)

    def real(self):                                       # :83-84
        return NumpyVectorOracle(np.real(self.array), self.options)

    def conjugate(self):                                  # :86-87
        return NumpyVectorOracle(self.array.conj(), self.options)

    def vdot(self, other, conjugate=True):                # :89-93
        if conjugate:
            return np.vdot(self.array, other.array)
        return np.dot(self.array.ravel(), other.array.ravel())

    def copy(self):                                       # :95-96
        return NumpyVectorOracle(self.array.copy(), self.options)

    def applyOp(self, other):                             # :98-100
        NumpyVectorOracle.matvec_count += 1
        return NumpyVectorOracle(other @ self.array, self.options)

    def compress(self):                                   # :102-103
        return self

    def linearCombination(vectors, coeffs):               # :105-119
        assert len(vectors) == len(coeffs)
        acc = np.zeros(len(vectors[0]), dtype=vectors[0].dtype)
        for n in range(len(vectors)):
            acc += coeffs[n] * vectors[n].array
        return NumpyVectorOracle(acc, vectors[0].options)

    def orthogonalize_against_set(x, qs, lindep=LINDEP_DEFAULT_VALUE):  # :121-145
        for q in qs:
            t1 = x.vdot(q, conjugate=False)
            t2 = q.vdot(q, conjugate=False)
            x = NumpyVectorOracle.linearCombination([x, q * (t1 / t2)], [1.0, -1.0])
        innerprod = x.vdot(x, conjugate=False)
        if innerprod > lindep:
            return x / np.sqrt(innerprod)
        return None

    @staticmethod
    def solve(H, b, sigma, x0=None, opType="her", reverseGF=False):  # :147-178
        n = H.shape[0]
        dtype = np.result_type(sigma, H.dtype, b.dtype)

        def shifted(x):
            NumpyVectorOracle.matvec_count += 1
            return (sigma * x - H @ x) if not reverseGF else (H @ x - sigma * x)
        linOp = spla.LinearOperator((n, n), matvec=shifted, dtype=dtype)
        opt = b.options["linearSystemArgs"]
        tol, atol, maxiter = opt["linear_tol"], opt["linear_atol"], opt["linearIter"]
        if opt["linearSolver"] == "gcrotmk":
            M = None
            if opt.get("preconditioner") == "jacobi":   # test counterpart of CudaVector's opt-in option: SciPy's M=
                den = (sigma - H.diagonal()) if not reverseGF else (H.diagonal() - sigma)
                M = spla.LinearOperator((n, n), matvec=lambda x: x / den, dtype=dtype)
            wk, conv = spla.gcrotmk(linOp, b.array, x0, M=M, rtol=tol, atol=atol, maxiter=maxiter)
        elif opt["linearSolver"] == "minres":
            wk, conv = spla.minres(linOp, b.array, x0, rtol=tol, maxiter=maxiter)
        elif opt["linearSolver"] == "pardiso":  # dense -> CSC spsolve, only for the Fortran comparison
            A1 = csc_matrix(sigma * np.eye(n) - H) if not reverseGF else csc_matrix(H - sigma * np.eye(n))
            wk = spla.spsolve(A1, csc_matrix(np.reshape(b.array, (n, 1))))
            conv = 0
        else:
            raise Exception("Got linear solver other than gcrotmk, minres and pardiso!")
        if conv != 0:
            warnings.simplefilter('error', UserWarning)
            warnings.warn("Warning:: Iterative solver is not converged ")
        return NumpyVectorOracle(wk, b.options)

    @staticmethod
    def sumOverRanks(vectors, like=None):
        """Test-only counterpart of CudaVector.sumOverRanks (node-distributed FEAST over gloo)."""
        import torch
        import torch.distributed as dist
        out = []
        for i, v in enumerate(vectors):
            arr = np.zeros(len(like[i])) if v is None else np.ascontiguousarray(v.array, dtype=np.float64)
            t = torch.from_numpy(arr)
            dist.all_reduce(t)
            out.append(NumpyVectorOracle(t.numpy(), like[i].options))
        return out

    def matrixRepresentation(operator, vectors):          # :180-190
        m = len(vectors)
        M = np.zeros((m, m), dtype=vectors[0].dtype)
        for j in range(m):
            ket = vectors[j].applyOp(operator)
            for i in range(j, m):
                M[i, j] = vectors[i].vdot(ket)
                M[j, i] = M[i, j].conj()
        return M

    def overlapMatrix(vectors):                           # :192-203
        m = len(vectors)
        S = np.zeros((m, m), dtype=vectors[0].dtype)
        for i in range(m):
            for j in range(i, m):
                S[i, j] = vectors[i].vdot(vectors[j], True)
                S[j, i] = S[i, j].conj()
        return S

    def extendMatrixRepresentation(operator, vectors, opMat):  # :205-221
        m = len(vectors)
        elems = np.empty((1, m), dtype=vectors[0].dtype)
        ket = vectors[-1].applyOp(operator)
        for i in range(m):
            elems[0, i] = vectors[i].vdot(ket)
        opMat = np.append(opMat, elems[:, :-1].conj(), axis=0)
        return np.append(opMat, elems.T, axis=1)

    def extendOverlapMatrix(vectors, overlap):            # :223-238
        m = len(vectors)
        elems = np.empty((1, m), dtype=vectors[0].dtype)
        for i in range(m):
            elems[0, i] = vectors[i].vdot(vectors[-1], True)
        overlap = np.append(overlap, elems[:, :-1].conj(), axis=0)
        return np.append(overlap, elems.T, axis=1)
