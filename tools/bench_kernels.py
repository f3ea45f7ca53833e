"""Micro-benchmarks of the individual kernels (CUDA events, L2 flushed between reps by using
working sets > 126 MB or an explicit flush).  Prints one JSON line per kernel; used while
optimising, not by the driver."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from eigensolvers_b200 import CudaVector, DeviceOperator, Runtime, _lib, hamiltonians as hm  # noqa: E402


def timeit(rt, fn, reps=20, flush=None):
    t = rt.torch
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.add_(1.0)
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        t.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts)), float(np.min(ts))


def main():
    rt = Runtime.get()
    t = rt.torch
    peak = 6458.4
    flush = t.zeros(64 * 1024 * 1024, dtype=t.float64, device=rt.device)  # 512 MB > L2
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    mats = {}
    if which in ("all", "lap"):
        mats["lap100"] = hm.laplacian3d(100)
    if which in ("all", "osc"):
        mats["osc2e6"] = hm.coupled_oscillators((20, 10, 10, 10, 10, 10))[0]
    for name, H in mats.items():
        n = H.shape[0]
        x = CudaVector(np.random.default_rng(0).standard_normal(n))
        y = rt.empty(n, 0)
        for fmt in ("csr", "sell", "dia"):
            op = DeviceOperator.from_host(H, fmt=fmt)
            ab = op.algorithmic_bytes()
            for label, fn in (
                ("spmv_plain", lambda: _lib.check(rt.lib.cv_spmv(rt.ctx, op.handle, 0, 0, 0.0, 0.0, x._ptr, y.data_ptr(), rt.stream))),
                ("spmv_shift", lambda: _lib.check(rt.lib.cv_spmv(rt.ctx, op.handle, 0, 1, 0.7, 0.0, x._ptr, y.data_ptr(), rt.stream))),
                ("spmv_plain_dots", lambda: _lib.check(rt.lib.cv_spmv_dots(rt.ctx, op.handle, 0, 0, 0.0, 0.0, x._ptr, y.data_ptr(), None, rt.stream))),
                ("spmv_shift_dots", lambda: _lib.check(rt.lib.cv_spmv_dots(rt.ctx, op.handle, 0, 1, 0.7, 0.0, x._ptr, y.data_ptr(), None, rt.stream))),
            ):
                for fl in ((None, flush) if fmt != "csr" else (flush,)):
                    med, mn = timeit(rt, fn, flush=fl)
                    print(json.dumps({"kernel": label, "matrix": name, "fmt": fmt, "n": n, "nnz": int(H.nnz),
                                      "padded": op.padded_nnz, "l2_flush": fl is not None, "ms": med * 1e3,
                                      "GBs": ab / med / 1e9, "frac_measured": ab / med / 1e9 / peak}))
            del op
    # BLAS-1 / tall-skinny at N = 2e7 (160 MB vectors, beyond L2)
    n = 20_000_000
    rng = np.random.default_rng(1)
    vs = [CudaVector(rng.standard_normal(n)) for _ in range(24)]
    w = CudaVector(rng.standard_normal(n))
    out = _lib.dbl_array(64)

    def rep(label, fn, nbytes):
        med, mn = timeit(rt, fn)
        print(json.dumps({"kernel": label, "n": n, "ms": med * 1e3, "GBs": nbytes / med / 1e9,
                          "frac_measured": nbytes / med / 1e9 / peak}))
    rep("dot", lambda: vs[0].vdot(vs[1]), 16 * n)
    rep("nrm2", lambda: vs[0].norm(), 8 * n)
    rep("scal", lambda: vs[0] * 1.5, 16 * n)
    for m in (4, 12, 24):
        vp, k1 = _lib.ptr_array([v._ptr for v in vs[:m]])
        wp, k2 = _lib.ptr_array([w._ptr])
        rep(f"tsdot_m{m}_b1", lambda: _lib.check(rt.lib.cv_tsdot(rt.ctx, n, 0, 1, m, vp, 1, wp, out, rt.stream)), (m + 1) * 8 * n)
        coef = np.ones((m, 1))
        rep(f"lincomb_m{m}_k1", lambda: CudaVector.linearCombination(vs[:m], list(coef[:, 0])), (m + 1) * 8 * n)
    rep("gs_m12", lambda: CudaVector.orthogonalize_against_set(w, vs[:12]), (2 * 12 + 3) * 8 * n)
    rep("lincomb_m24_k24", lambda: CudaVector.linearCombinationBlock(vs, np.ones((24, 24))), (24 + 24) * 8 * n)
    print(json.dumps({"launches": rt.launch_count()}))


if __name__ == "__main__":
    main()
