"""Launch each hot kernel a couple of times on realistic sizes (for `ncu --set full`):
fused SpMV on the N=2e6 oscillator Hamiltonian (460 MB of matrix, beyond L2) and the vector
kernels at N=2e7 (160 MB vectors)."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from eigensolvers_b200 import CudaVector, DeviceOperator, Runtime, _lib, hamiltonians as hm  # noqa: E402


def main():
    rt = Runtime.get()
    t = rt.torch
    H = hm.coupled_oscillators((20, 10, 10, 10, 10, 10))[0]
    n = H.shape[0]
    x = CudaVector(np.random.default_rng(0).standard_normal(n))
    y = rt.empty(n, 0)
    for fmt in ("sell", "dia"):
        op = DeviceOperator.from_host(H, fmt=fmt)
        for _ in range(2):
            _lib.check(rt.lib.cv_spmv(rt.ctx, op.handle, 0, 0, 0.0, 0.0, x._ptr, y.data_ptr(), rt.stream))
            _lib.check(rt.lib.cv_spmv_dots(rt.ctx, op.handle, 0, 1, 0.7, 0.0, x._ptr, y.data_ptr(), None, rt.stream))
    t.cuda.synchronize()
    n = 20_000_000
    rng = np.random.default_rng(1)
    vs = [CudaVector(rng.standard_normal(n)) for _ in range(24)]
    w = CudaVector(rng.standard_normal(n))
    out = _lib.dbl_array(64)
    for _ in range(2):
        vs[0].vdot(vs[1])
        vs[0].norm()
        for m in (12, 24):
            vp, k1 = _lib.ptr_array([v._ptr for v in vs[:m]])
            wp, k2 = _lib.ptr_array([w._ptr])
            _lib.check(rt.lib.cv_tsdot(rt.ctx, n, 0, 1, m, vp, 1, wp, out, rt.stream))
        CudaVector.linearCombination(vs, [1.0] * 24)
        CudaVector.orthogonalize_against_set(w, vs[:3])
    t.cuda.synchronize()
    print("launches", rt.launch_count())


if __name__ == "__main__":
    main()
