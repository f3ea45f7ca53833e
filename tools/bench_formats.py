"""General-sparsity evidence for the fused shifted SpMV: the C3 Hamiltonian (N = 2e7 by default) in its
natural order and under two seeded symmetric permutations P H P^T that destroy the diagonal
structure (DIA storage no longer applies, the general formats must carry it):

  natural   product-basis order: 25 distinct column offsets -> DIA / SELL-32x2 / CSR
  window    rows shuffled inside windows of 4096: ~1e5 distinct offsets, gathers stay within the
            original band (cache friendly)                       -> SELL-32x2 / CSR
  random    a full random permutation of all N indices: every gather is a random 32-byte sector of
            a 160 MB vector (the worst case of a general sparse matrix) -> SELL-32x2 / CSR

Per (order, format): CUDA-event time of cv_spmv_dots (fused shift + dots), GB/s on SURVEY 8d's
12*nnz + 20*N, on the bytes the stored format moves, and the fraction of the measured copy peak.

    python tools/bench_formats.py [c3|c3mid] [orders...]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eigensolvers_b200 import CudaVector, DeviceOperator, Runtime, _lib  # noqa: E402
from eigensolvers_b200.workloads import build_workload  # noqa: E402


def permuted(H, perm):
    """P H P^T for the symmetric relabelling old index perm[i] -> new index i."""
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm), dtype=perm.dtype)
    A = H[perm]                                    # row gather
    A.indices = inv[A.indices].astype(np.int32)    # column relabel (rows are left unsorted: no kernel needs order)
    A.has_sorted_indices = True                    # skip scipy's sort of 4e8 entries; see above
    return A


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    orders = sys.argv[2:] or ["natural", "window", "random"]
    rt = Runtime.get()
    t = rt.torch
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    w = build_workload(name)
    H, n = w["H"], w["N"]
    rng = np.random.default_rng(12)
    x = rng.standard_normal(n)
    flush = t.zeros(48 * 1024 * 1024, dtype=t.float64, device=rt.device)
    for order in orders:
        t0 = time.time()
        if order == "natural":
            A, xs = H, x
        else:
            if order == "window":
                win = 4096
                perm = np.concatenate([s + rng.permutation(min(win, n - s)) for s in range(0, n, win)]).astype(np.int64)
            else:
                perm = rng.permutation(n).astype(np.int64)
            A, xs = permuted(H, perm), x[perm]
        ref = None
        for fmt in (("dia", "sell", "csr") if order == "natural" else ("sell", "csr")):
            try:
                op = DeviceOperator.from_host(A, fmt=fmt)
            except Exception as e:
                print(json.dumps({"order": order, "fmt": fmt, "error": str(e)[:200]}), flush=True)
                continue
            X = CudaVector(xs)
            y = rt.empty(n, 0)
            out3 = (_lib.C.c_double * 3)()

            def run():
                _lib.check(rt.lib.cv_spmv_dots(rt.ctx, op.handle, 0, 1, w["sigma"], 0.0, X._ptr, y.data_ptr(), out3, rt.stream))
            for _ in range(3):
                run()
            got = y.cpu().numpy()
            if ref is None:
                ref = w["sigma"] * xs - A @ xs
            err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
            ts = []
            for _ in range(10):
                flush.add_(1.0)
                e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
                e0.record()
                run()
                e1.record()
                t.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            alg = 12 * A.nnz + 20 * n
            fmtb = {"dia": 8 * getattr(op, "padded_nnz", 0) + 16 * n, "sell": 12 * op.padded_nnz + 8 * (n // 32) + 16 * n,
                    "csr": 12 * A.nnz + 8 * (n + 1) + 16 * n}[op.format]
            print(json.dumps({"order": order, "fmt": op.format, "n": n, "nnz": int(A.nnz), "padded_nnz": int(op.padded_nnz),
                              "ms": round(ms, 4), "GBs_12nnz_20N": round(alg / ms / 1e6, 1), "frac_of_measured_peak": round(alg / ms / 1e6 / peak, 3),
                              "GBs_format_bytes": round(fmtb / ms / 1e6, 1), "max_rel_err_vs_scipy": err,
                              "setup_s": round(time.time() - t0, 1)}), flush=True)
            del op, X, y
        del A


if __name__ == "__main__":
    main()
