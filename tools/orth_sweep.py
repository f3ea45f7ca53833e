"""Sweep of the fused Arnoldi-step kernel (k_orth_step) over the basis size m and the dot-phase work
split (slab_mode 0 / 1), through the C-ABI entry cv_arnoldi_step: per-launch CUDA-event time, the
kernel's own phase stamps (dots / barrier+all-reduce / update) and GB/s on (2m+3)*8N.

    python tools/orth_sweep.py [N=20000000] [cplx=0] [snake]     # "snake": A/B the update-phase row order, L2 flushed between launches
"""
import ctypes as C
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from eigensolvers_b200 import Runtime, _lib  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
    cplx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rt = Runtime.get()
    t = rt.torch
    mmax = 60
    dt = t.complex128 if cplx else t.float64
    g = t.Generator(device=rt.device).manual_seed(1)
    basis = [(t.randn(n, dtype=t.float64, device=rt.device, generator=g) / np.sqrt(n)).to(dt) for _ in range(mmax)]
    w0 = t.randn(n, dtype=t.float64, device=rt.device, generator=g).to(dt)
    ww = float((w0.abs() ** 2).sum().item())
    eb = 16 if cplx else 8
    tr = (C.c_double * 16)()
    out = (C.c_double * (2 + 2 * mmax))()
    snake_sweep = len(sys.argv) > 3 and sys.argv[3] == "snake"
    flush = t.zeros(40 * 1024 * 1024, dtype=t.float64, device=rt.device)   # 320 MB: what the SpMV between two steps does to L2
    for m in ((21, 28, 36, 48) if snake_sweep else (1, 5, 12, 16, 17, 21, 24, 28, 32, 33, 36, 40, 48, 49, 60)):
        ptrs, _keep = _lib.ptr_array([b.data_ptr() for b in basis[:m]])
        for mode in (0, 1):
            if snake_sweep:   # mode = snake off / on, new slab split
                _lib.check(rt.lib.cv_ctx_set_option(rt.ctx, b"slab_mode", 1.0))
                _lib.check(rt.lib.cv_ctx_set_option(rt.ctx, b"snake", float(mode)))
            else:
                _lib.check(rt.lib.cv_ctx_set_option(rt.ctx, b"slab_mode", float(mode)))
            w = w0.clone()
            # correctness of h against torch on the first launch
            _lib.check(rt.lib.cv_arnoldi_step(rt.ctx, None, n, cplx, m, ptrs, w.data_ptr(), ww, 0.1, out, rt.stream))
            h = np.ctypeslib.as_array(out)[2:2 + m * (2 if cplx else 1)].copy()
            if cplx:
                h = h.view(np.complex128)
            href = np.array([(b.conj() * w0).sum().item() for b in basis[:m]])
            err = float(np.max(np.abs(h - href)) / max(np.max(np.abs(href)), 1e-300))
            reps = 8
            _lib.check(rt.lib.cv_ctx_trace_read(rt.ctx, tr, 1))
            evs = []
            for _ in range(reps):
                w.copy_(w0)
                if snake_sweep:
                    flush.add_(1.0)
                e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(rt.lib.cv_arnoldi_step(rt.ctx, None, n, cplx, m, ptrs, w.data_ptr(), ww, 0.1, out, rt.stream))
                e1.record()
                t.cuda.synchronize()
                evs.append(e0.elapsed_time(e1))
            _lib.check(rt.lib.cv_ctx_trace_read(rt.ctx, tr, 1))
            k = max(tr[5], 1.0)
            ms = float(np.median(evs))
            print(json.dumps({"m": m, ("snake" if snake_sweep else "slab_mode"): mode, "n": n, "cplx": cplx, "ms": round(ms, 4),
                              "GBs": round((2 * m + 3) * eb * n / ms / 1e6, 1),
                              "dots_us": round(tr[0] / k * 1e-3, 1), "barrier_us": round(tr[1] / k * 1e-3, 1),
                              "update_us": round(tr[2] / k * 1e-3, 1), "second_pass": out[0], "h_rel_err": err}), flush=True)


if __name__ == "__main__":
    main()
