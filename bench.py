#!/usr/bin/env python
"""bench.py — time-to-eConv per eigenpair of the inexact shift-and-invert Lanczos hot path on
B200, with the roofline of its dominant kernels and the reference's CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
                    [--driver auto|reference|mirror] [--no-cpu] [--extras]

One "step" = one complete `inexactLanczosDiagonalization` (c5: `feastDiagonalization`) run on the
workload (eigensolvers_b200/workloads.py):
  c3      (default) coupled-oscillator product-basis Hamiltonian, N = 2e7 (BASELINE configs[2],
          the configuration BASELINE.json's metric is quoted on; it fits one B200), single
          guess, sigma a quarter-gap above the 9th analytic level, L=8, eConv=1e-10,
          GCROT(20,20) rtol 1e-4.  Row-sharded over N GPUs (strong scaling: the problem is fixed).
  c2      block Lanczos (4 guesses) on the 100^3 Laplacian + random potential, N = 1e6
  c4      near-linearly-dependent block start, N = 5e7 oscillator Hamiltonian, L = 100, rtol 1e-1
  c5      FEAST nc=16 (8 retained nodes, one per GPU at --gpus 8), m0 = 6, N = 2e7
  *mid / *small   the same generators at reduced N (development)

The driver is the reference's OWN unchanged inexact_Lanczos.py / feast.py from baseline/_ref when
that installation is present (--driver auto), else the stand-alone mirror in eigensolvers_b200.

JSON keys follow the driver's contract; `value` is seconds per eigenpair with everything resident
in HBM, `e2e` the same through the public API from pinned HOST buffers (H and guesses copied
host->device and the eigenvectors device->host inside the timed region).  `roofline` describes the
kernel with the largest share of the step (the fused Arnoldi step), `roofline_spmv` the fused
shifted SpMV; both are algorithmic bytes (accumulated per launch inside libcudavec, SURVEY 8d
formulas) over CUDA-event time on the launching stream.

`--impl reference`: the reference's NumpyVector path on the host cores — one CONTINUOUS slice of
the real run (its first (W+K)*n operator applications with everything SciPy's GCROT does between
them), cut into W+K windows of n; `value` extrapolates the measured seconds per operator
application to the number of applications the unmodified reference needed for the full run
(tests/golden/<workload>_full.npz), `ms_per_step` is the measured time of one window.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
import traceback
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from eigensolvers_b200.workloads import WORKLOADS, build_workload, solver_options  # noqa: E402  (no libcudavec)

# operator applications of one full run, used ONLY when no golden reference run exists for the
# workload (GPU-measured; flagged in the `sample` text)
MATVECS_FALLBACK = {"c3": 4250, "c3mid": 6772, "c3small": 3738, "c3tiny": 900, "c2": 139000, "c2small": 3000,
                    "c4": 20000, "c4mid": 20000, "c4small": 5000, "c5": 40000, "c5mid": 40000, "c5small": 40000}


def reference_matvecs(name):
    """(count, source): operator applications the UNMODIFIED reference needed for this workload."""
    path = os.path.join(ROOT, "tests", "golden", f"{name}_full.npz")
    if os.path.exists(path):
        meta = json.loads(str(np.load(path)["meta"]))
        return int(meta["total_matvecs"]), f"measured on the unmodified reference (tests/golden/{name}_full.npz)"
    return MATVECS_FALLBACK[name], "GPU-measured count (no full reference run of this workload is committed)"


def config_dict(args, w, nnz_total):
    cfg = {"workload": f"{args.workload}: {w['label']}, N={w['N']}, nnz={nnz_total}, nBlock={w['nBlock']}, "
                       f"sigma={w['sigma']:.6f}, L={w.get('L')}, maxit={w['maxit']}, eConv={w['eConv']:g}, "
                       f"gcrotmk rtol={w['tol']:g}",
           "parallelism": f"row-shard x{args.gpus}" if w["kind"] != "osc_feast" else f"quadrature nodes over {args.gpus} GPU(s)",
           "l2": "working set >> 126 MB L2; L2 also flushed between steps"}
    if w.get("deviation"):
        cfg["deviation"] = w["deviation"]
    return cfg


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.path = os.path.join(tempfile.gettempdir(), f"clocks_{os.getpid()}.csv")
        self.proc = None

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as fh:
            for line in fh:
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for nm, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(mx)), reasons=sorted(reasons), samples=len(sm))
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------------------------- CPU arm
class _SliceDone(Exception):
    pass


class TimedOperator:
    """H with a time stamp at the start of every application; raises after `limit` of them."""

    def __init__(self, H, limit):
        self.H, self.shape, self.dtype = H, H.shape, H.dtype
        self.limit, self.stamps = limit, []

    def __matmul__(self, x):
        self.stamps.append(time.perf_counter())
        if len(self.stamps) > self.limit:
            raise _SliceDone()
        return self.H @ x


def _reference_backend():
    """The reference's own modules (baseline/_ref) when installed, else None."""
    try:
        from eigensolvers_b200 import refdrivers
        if refdrivers.available():
            return refdrivers.load(register_cuda=False, numpy_backend=True)
    except Exception:
        pass
    return None


_CPU_THREADS = [None]


def cpu_threads(w):
    """BLAS thread count for the CPU legs: all host cores or one, whichever runs a short probe of the
    real path faster (threaded BLAS-1 on vectors this long LOSES to one thread on oversubscribed or
    bandwidth-starved hosts; csr_matvec is serial either way)."""
    if _CPU_THREADS[0] is None:
        best = None
        for t in sorted({os.cpu_count() or 1, 1}, reverse=True):
            stamps, _, _ = cpu_slice(w, 6 + w["nBlock"], threads=t)
            dt = float(stamps[-1] - stamps[w["nBlock"]]) if len(stamps) > w["nBlock"] + 1 else 1e30
            if best is None or dt < best[0]:
                best = (dt, t)
        _CPU_THREADS[0] = best[1]
    return _CPU_THREADS[0]


def cpu_slice(w, n_total, threads=None):
    """Run the reference's CPU path on workload `w` until it has applied the operator n_total times;
    returns (stamps[n_total+1], kind, threads).  kind "reference": the unmodified reference driver +
    NumpyVector from baseline/_ref; "port": the same SciPy call NumpyVector.solve makes
    (numpyVector.py:152,161) when that installation is absent."""
    if threads is None:
        threads = cpu_threads(w)
    try:  # all host threads for the BLAS-1 part, also under torchrun (which exports OMP_NUM_THREADS=1)
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=threads)
    except Exception:
        limiter = None
    Ht = TimedOperator(w["H"], n_total)
    ns = _reference_backend()
    opts = solver_options(w)
    kind = "reference" if ns is not None else "port"
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if ns is not None and w["kind"] != "osc_feast":
                NV = ns.numpyVector.NumpyVector
                gs = [NV(g.copy(), opts) for g in w["guesses"]]
                ns.inexactLanczosDiagonalization(Ht, gs[0] if len(gs) == 1 else gs, w["sigma"], w["L"], w["maxit"],
                                                 w["eConv"], writeOut=False, saveTNSsEachIteration=False)
            elif ns is not None:
                NV = ns.numpyVector.NumpyVector
                gs = [NV(g.copy(), opts) for g in w["guesses"]]
                ns.feastDiagonalization(Ht, gs, w["nc"], "legendre", w["eMin"], w["eMax"], w["eConv"], w["maxit"],
                                        writeOut=False)
            else:
                import scipy.sparse.linalg as spla
                n = w["N"]
                sigma = w["sigma"]
                lin = spla.LinearOperator((n, n), matvec=lambda x: sigma * x - Ht @ x, dtype=np.float64)
                b = w["guesses"][0] / np.linalg.norm(w["guesses"][0])
                la = opts["linearSystemArgs"]
                spla.gcrotmk(lin, b, None, rtol=la["linear_tol"], atol=la["linear_atol"], maxiter=la["linearIter"])
    except _SliceDone:
        pass
    finally:
        warnings.resetwarnings()
        if limiter is not None:
            limiter.restore_original_limits()
    return np.asarray(Ht.stamps), kind, threads


def cpu_plan(n_windows, budget_s, per_mv_guess):
    """operator applications per window: at least two full GCROT(20,20) cycles (80) in total, more if
    the time budget allows, at most one cycle (40) per window."""
    total_min = 80
    per = max(int(np.ceil(total_min / n_windows)), min(40, int(budget_s / max(per_mv_guess, 1e-3) / n_windows)))
    return max(per, 4)


def run_reference_arm(args, w):
    n_win = args.warmup + args.steps
    # size the windows from a 6-application probe (first GCROT steps: cheaper than average, so doubled)
    probe, _, _ = cpu_slice(w, 6 + w["nBlock"])
    per_mv_guess = 2.0 * float(np.mean(np.diff(probe[w["nBlock"]:]))) if len(probe) > w["nBlock"] + 1 else 1.0
    n_per = cpu_plan(n_win, args.cpu_budget, per_mv_guess)
    stamps, kind, threads = cpu_slice(w, n_win * n_per + w["nBlock"] + 1)
    # the first nBlock applications belong to the driver's start-up (matrixRepresentation); windows start after them
    s0 = 0 if w["kind"] == "osc_feast" else w["nBlock"]
    got = (len(stamps) - s0) // n_per
    win = [stamps[s0 + (i + 1) * n_per] - stamps[s0 + i * n_per] for i in range(got) if s0 + (i + 1) * n_per < len(stamps)]
    timed = win[args.warmup:] if len(win) > args.warmup else win
    per_mv = float(np.mean(timed)) / n_per
    total, source = reference_matvecs(args.workload)
    value = per_mv * total / w["nBlock"]
    sample = (f"one continuous slice of the run: its first {len(win) * n_per} operator applications (scipy csr_matvec, serial) "
              f"with everything SciPy's GCROT(20,20) does between them (BLAS-1 on {threads} threads), cut into {len(win)} windows "
              f"of {n_per}; the last {len(timed)} windows are timed: {np.mean(timed):.2f} s each = {per_mv:.3f} s per application; "
              f"value = that x {total} applications of the full run [{source}] / nBlock; host has {os.cpu_count()} cores")
    line = {
        "impl": "reference", "metric": "time_to_eConv_per_eigenpair", "value": value, "unit": "s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(timed)) * 1e3, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, w, w["nnz"]),
        "cpu_baseline": {"value": value, "unit": "s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "windows_s": [round(float(x), 3) for x in win],
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- GPU arm
PROF_NAMES = ["spmv", "arnoldi_step", "tsupdate", "other_vector", "gram_schmidt_set", "lincomb", "spmv_csr_equivalent", "spare"]


def read_profile(rt):
    import ctypes as C
    from eigensolvers_b200 import _lib
    ms, cnt, by = (C.c_double * 8)(), (C.c_uint64 * 8)(), (C.c_double * 8)()
    _lib.check(rt.lib.cv_ctx_profile_read(rt.ctx, 8, ms, cnt, by))
    return [float(x) for x in ms], [int(x) for x in cnt], [float(x) for x in by]


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def ncu_traffic(workload, fmt, world, kernel):
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            table = json.load(fh)
        return table.get(f"{workload}/{fmt}/{world}/{kernel}")
    except Exception:
        return None


def roofline_entry(kernel, ms, cnt, by, peak, peak_src, note, traffic=None, extra=None):
    avg_ms = ms / max(cnt, 1)
    achieved = by / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    out = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
           "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0,
           "traffic": traffic["bytes"] if traffic else None, "traffic_source": traffic["source"] if traffic else None,
           "algorithmic_bytes_per_launch": by / max(cnt, 1), "launches_timed": cnt, "avg_launch_ms": avg_ms, "note": note}
    if extra:
        out.update(extra)
    return out


def run_ours(args, w):
    import torch
    import torch.distributed as dist
    from eigensolvers_b200 import CudaVector, DeviceOperator, Runtime, _lib, refdrivers
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    feast = w["kind"] == "osc_feast"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    rt = Runtime.get()
    if world > 1 and not feast:
        rt.init_distributed()          # row-sharded mode; FEAST replicates H and distributes nodes instead
    opts = solver_options(w)
    if args.preconditioner:   # development: the whole run (headline included) with the opt-in preconditioner
        opts["linearSystemArgs"]["preconditioner"] = args.preconditioner
    H = w["H"]
    use_ref = args.driver == "reference" or (args.driver == "auto" and refdrivers.available())
    if feast:
        if world == 1 and use_ref:
            drv, drv_name = refdrivers.feast_driver()
        else:
            from eigensolvers_b200.contour import feastDiagonalization as drv
            drv_name = "eigensolvers_b200.contour (reference feast.py control flow + node distribution over ranks)"
    elif use_ref:
        drv, drv_name = refdrivers.load().inexactLanczosDiagonalization, "reference inexact_Lanczos.py (unchanged, baseline/_ref)"
    else:
        from eigensolvers_b200.lanczos import inexactLanczosDiagonalization as drv
        drv_name = "eigensolvers_b200.lanczos (mirror)"

    # pinned host copies of the inputs (the e2e leg copies from these every step)
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy()
    import scipy.sparse as sp
    Hp = sp.csr_matrix((pinned(H.data), pinned(H.indices.astype(np.int32)), pinned(H.indptr.astype(np.int64))),
                       shape=H.shape, copy=False)
    Hp.has_sorted_indices = True
    sharded = world > 1 and not feast

    def make_operator():
        if sharded:
            return DeviceOperator.from_local_rows(Hp, w["N"], runtime=rt)
        return DeviceOperator.from_host(Hp, runtime=rt)
    guesses_p = [pinned(g) for g in w["guesses"]]
    flush = torch.zeros(48 * 1024 * 1024, dtype=torch.float64, device=rt.device)  # 384 MB > L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=rt.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def one_run(op, guess_dev, o=opts):
        vecs = [CudaVector._wrap(g.clone(), dict(o), w["N"]) for g in guess_dev]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if feast:
                kw = dict(distribute=args.distribute) if not (world == 1 and use_ref) else {}
                out = drv(op, vecs, w["nc"], "legendre", w["eMin"], w["eMax"], w["eConv"], w["maxit"], writeOut=False, **kw)
            else:
                v0 = vecs[0] if w["nBlock"] == 1 else vecs
                kw = dict(saveTNSsEachIteration=False) if use_ref else {}
                out = drv(op, v0, w["sigma"], w["L"], w["maxit"], w["eConv"], writeOut=False, **kw)
        warnings.resetwarnings()
        return out

    nnz_total = torch.tensor([float(H.nnz)], dtype=torch.float64, device=rt.device)
    if sharded:
        dist.all_reduce(nnz_total)
    nnz_total = int(nnz_total.item())

    # ---- resident leg: operator and guesses already in HBM
    op = make_operator()
    guess_dev = [CudaVector(g, dict(opts))._t for g in guesses_p]
    for _ in range(args.warmup):
        flush.add_(1.0)
        ev, Y, st = one_run(op, guess_dev)
    barrier()
    launches0 = rt.launch_count()
    mv0, solves0 = rt.stats["matvecs"], rt.stats["solves"]
    tr16 = (C.c_double * 16)()
    _lib.check(rt.lib.cv_ctx_trace_read(rt.ctx, tr16, 1))
    clocks = ClockSampler(rt.device_index)
    clocks.start()
    marks = [torch.cuda.Event(enable_timing=True)]
    marks[0].record()
    for _ in range(args.steps):
        flush.add_(1.0)
        ev, Y, st = one_run(op, guess_dev)
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record()
    barrier()
    clk = clocks.stop()
    ms = reduce_max(marks[0].elapsed_time(marks[-1]))
    step_ms = [round(marks[i].elapsed_time(marks[i + 1]), 1) for i in range(args.steps)]
    _lib.check(rt.lib.cv_ctx_trace_read(rt.ctx, tr16, 1))
    nk = max(tr16[5], 1.0)
    orth_trace = {k: round(tr16[i] / nk * 1e-3, 2) for i, k in enumerate(
        ["dots_us", "barrier_allreduce_us", "update_normalise_push_us", "second_pass_barrier_us"])}
    orth_trace.update(launches=int(tr16[5]), passes=int(tr16[6]), halo_flags_us=round(tr16[7] / nk * 1e-3, 2))
    launches = rt.launch_count() - launches0
    matvecs = (rt.stats["matvecs"] - mv0) // max(args.steps, 1)
    solves = (rt.stats["solves"] - solves0) // max(args.steps, 1)
    ms_per_step = ms / args.steps
    n_eig = w["nBlock"]
    value = ms_per_step * 1e-3 / n_eig
    converged = bool(st["isConverged"])
    if feast:   # feast.py leaves status["isConverged"] untouched; its stop rule is residual < eConv (feast.py:231)
        converged = st.get("residual") is not None and float(st["residual"]) < w["eConv"]
    ev_arr = np.asarray(ev, dtype=float)
    if feast:
        ev_out = [float(x) for x in np.sort(ev_arr[(ev_arr > w["eMin"]) & (ev_arr < w["eMax"])])]
    else:
        ev_out = [float(x) for x in np.sort(ev_arr[:w["nBlock"]])]

    feast_profile = getattr(drv, "last_profile", None) if feast else None

    # ---- profiled step: per-launch durations (CUDA events on the stream) and algorithmic bytes per class
    if args.no_profile:
        t_prof, pms, pcnt, pby = 1.0, [0.0] * 8, [0] * 8, [0.0] * 8
    else:
        _lib.check(rt.lib.cv_ctx_profile(rt.ctx, 1))
        read_profile(rt)
        t_prof = time.perf_counter()
        one_run(op, guess_dev)
        torch.cuda.synchronize()
        t_prof = time.perf_counter() - t_prof
        pms, pcnt, pby = read_profile(rt)
        _lib.check(rt.lib.cv_ctx_profile(rt.ctx, 0))
    peak, peak_src = measured_peak()
    share = {PROF_NAMES[i]: pms[i] / (t_prof * 1e3) for i in range(6)}
    share["host_and_launch_gaps"] = max(0.0, 1.0 - sum(share.values()))   # reference driver's host work (m x m eigh, ...)
    fmt = op.format
    cplx_note = " (complex128 vectors)" if feast else ""
    roof_orth = roofline_entry(
        "k_orth_step (fused Arnoldi step: h = [C,V]^H w, all-reduce, w <- (w - [C,V]h)/|w'|, halo push)" + cplx_note,
        pms[1], pcnt[1], pby[1], peak, peak_src,
        "algorithmic bytes = sum over launches of (2m+3)*8N (m = basis size of that step; +(2m+2)*8N when the second "
        "Gram-Schmidt pass ran); includes the grid barrier + cross-GPU all-reduce inside the kernel",
        ncu_traffic(args.workload, fmt, world, "k_orth_step"), {"step_share": share[PROF_NAMES[1]]})
    roof_spmv = roofline_entry(
        ("k_spmv_dia2<double> (two rows per thread, fused shift + dots)" if fmt == "dia" and not feast
         else f"k_spmv_{fmt} (fused shift + dots)" + cplx_note),
        pms[0], pcnt[0], pby[0], peak, peak_src,
        f"algorithmic bytes = what the stored format ({fmt}) must move: matrix stream + x + y"
        " (DIA: 8*D*ld, no index stream); csr_equivalent_GBs quotes SURVEY 8d's 12*nnz+20*N over the same time",
        ncu_traffic(args.workload, fmt, world, "k_spmv"),
        {"step_share": share[PROF_NAMES[0]], "format": fmt,
         "csr_equivalent_GBs": pby[6] / (pms[0] * 1e-3) / 1e9 if pms[0] > 0 else None,
         "csr_equivalent_frac": (pby[6] / (pms[0] * 1e-3) / 1e9 / peak) if pms[0] > 0 else None})
    dominant = roof_orth if pms[1] >= pms[0] else roof_spmv
    roof_gs = None
    if pcnt[4]:
        roof_gs = roofline_entry("k_mgs_step chain (orthogonalize_against_set, numpyVector.py:121-145)", pms[4], pcnt[4], pby[4],
                                 peak, peak_src, "algorithmic bytes = (4m+2)*8N per call (sequential MGS as the reference: "
                                 "x read+written and two basis vectors per step); launches_timed counts calls",
                                 None, {"step_share": share[PROF_NAMES[4]]})

    # ---- end-to-end leg: host buffers in, host eigenvectors out, every step
    h2d = nnz_total * 12 + (w["N"] + world) * 8 + len(guesses_p) * w["N"] * 8      # all ranks together
    if feast:
        h2d = world * (nnz_total * 12 + (w["N"] + 1) * 8 + len(guesses_p) * w["N"] * 8)   # replicated inputs
    n_out = len(ev_out) if feast else w["nBlock"]
    d2h = max(n_out, 1) * w["N"] * 8 * (world if feast else 1)                        # each rank reads its own rows
    feast_tasks = None
    if feast and world > 1 and args.feast_tasks:
        # informational: the same FEAST run with the (node, vector) solves spread by measured cost
        try:
            vecs_t = [CudaVector._wrap(g.clone(), dict(opts), w["N"]) for g in guess_dev]
            barrier()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                evt, Yt, stt = drv(op, vecs_t, w["nc"], "legendre", w["eMin"], w["eMax"], w["eConv"], w["maxit"],
                                   writeOut=False, distribute=args.feast_tasks)
            warnings.resetwarnings()
            q1.record()
            barrier()
            feast_tasks = {"seconds": reduce_max(q0.elapsed_time(q1)) * 1e-3, "iterations": int(stt["outerIter"]) + 1,
                           "profile": getattr(drv, "last_profile", None)}
            del vecs_t, Yt
        except Exception as e:
            feast_tasks = {"error": f"{type(e).__name__}: {e}"}
    del op
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    host_vecs = []
    ev2 = ev
    for _ in range(0 if args.no_e2e else args.steps):
        flush.add_(1.0)
        op2 = make_operator()                                                # H2D of the Hamiltonian
        gd = [CudaVector(g, dict(opts))._t for g in guesses_p]               # H2D of the guesses
        ev2, Y2, st2 = one_run(op2, gd)
        host_vecs = [Y2[i].local_array for i in range(min(max(n_out, 1), len(Y2)))]   # D2H of the eigenvectors
        del op2
    e3.record()
    barrier()
    e2e_value = None if args.no_e2e else reduce_max(e2.elapsed_time(e3)) / args.steps * 1e-3 / n_eig
    true_res = None
    if world == 1 and host_vecs:
        x = host_vecs[0]
        true_res = float(np.linalg.norm(H @ x - ev2[0] * x))

    # ---- informational: the same run on the MATRIX-FREE Kronecker-sum form of the same Hamiltonian;
    # the headline above takes H as the reference does (a scipy CSR matrix) and stores it (DIA)
    matrix_free = None
    if w["kind"] == "osc" and not args.no_extras:
        try:
            from eigensolvers_b200 import KroneckerSumOperator
            kop = KroneckerSumOperator.coupled_oscillators(w["dims"], coupling=0.1, seed=1, runtime=rt)
            one_run(kop, guess_dev)
            barrier()
            mvk = rt.stats["matvecs"]
            _lib.check(rt.lib.cv_ctx_profile(rt.ctx, 1))
            read_profile(rt)
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            evk, Yk, stk = one_run(kop, guess_dev)
            k1.record()
            barrier()
            kms, kcnt, kby = read_profile(rt)
            _lib.check(rt.lib.cv_ctx_profile(rt.ctx, 0))
            matrix_free = {"value": reduce_max(k0.elapsed_time(k1)) * 1e-3 / n_eig, "unit": "s", "format": kop.format,
                           "matvecs": int(rt.stats["matvecs"] - mvk), "converged": bool(stk["isConverged"]),
                           "eigenvalues": [float(v) for v in np.sort(np.asarray(evk, dtype=float)[:w["nBlock"]])],
                           "spmv_avg_launch_ms": kms[0] / max(kcnt[0], 1),
                           "spmv_bytes_moved_GBs": kby[0] / (kms[0] * 1e-3) / 1e9 if kms[0] > 0 else None,
                           "spmv_csr_equivalent_GBs": kby[6] / (kms[0] * 1e-3) / 1e9 if kms[0] > 0 else None,
                           "note": "H passed as KroneckerSumOperator (1-D factors only, nothing of the N x N matrix stored); "
                                   "not the headline"}
            del kop
        except Exception as e:
            matrix_free = {"error": f"{type(e).__name__}: {e}"}

    # ---- informational: the same run with JACOBI-PRECONDITIONED inner solves (opt-in linearSystemArgs["preconditioner"],
    # SciPy's M= argument).  The reference never passes M, so the headline stays unpreconditioned.
    preconditioned = None
    if w["kind"] == "osc" and not args.no_extras and not args.preconditioner:
        try:
            popts = {"linearSystemArgs": dict(opts["linearSystemArgs"], preconditioner="jacobi")}
            opp = make_operator()
            one_run(opp, guess_dev, popts)
            barrier()
            mvp = rt.stats["matvecs"]
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            evp, Yp, stp = one_run(opp, guess_dev, popts)
            p1.record()
            barrier()
            pv = reduce_max(p0.elapsed_time(p1)) * 1e-3 / n_eig
            preconditioned = {"value": pv, "unit": "s", "speedup_vs_headline": value / pv,
                              "matvecs": int(rt.stats["matvecs"] - mvp), "converged": bool(stp["isConverged"]),
                              "cumIter": int(stp["cumIter"]),
                              "eigenvalues": [float(v) for v in np.sort(np.asarray(evp, dtype=float)[:w["nBlock"]])],
                              "overlap_with_headline_eigenvector": (float(abs(Y[0].vdot(Yp[0]))) if converged and stp["isConverged"] else None),
                              "note": "GCROT with the right preconditioner M = diag(1/(sigma - H_ii)); same rtol on the true residual, "
                                      "same driver; an algorithmic option the reference does not have — not the headline"}
            del opp, Yp
        except Exception as e:
            preconditioned = {"error": f"{type(e).__name__}: {e}"}

    # ---- informational: block solves advanced in LOCK STEP (cv_solve_batch) by the mirror driver — the reference's
    # unchanged driver calls solve() once per block vector and normalises in between, so it cannot batch
    lockstep = None
    if not feast and w["nBlock"] > 1 and world == 1 and not args.no_extras:
        try:
            from eigensolvers_b200.lanczos import inexactLanczosDiagonalization as mirror
            opl = make_operator()

            def lock_run():
                vecs = [CudaVector._wrap(g.clone(), dict(opts), w["N"]) for g in guess_dev]
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    out = mirror(opl, vecs, w["sigma"], w["L"], w["maxit"], w["eConv"], writeOut=False, lockstep=True)
                warnings.resetwarnings()
                return out
            lock_run()
            barrier()
            mvl = rt.stats["matvecs"]
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record()
            evl, Yl, stl = lock_run()
            l1.record()
            barrier()
            lv = reduce_max(l0.elapsed_time(l1)) * 1e-3 / n_eig
            lockstep = {"value": lv, "unit": "s", "speedup_vs_headline": value / lv, "matvecs": int(rt.stats["matvecs"] - mvl),
                        "converged": bool(stl["isConverged"]), "cumIter": int(stl["cumIter"]),
                        "eigenvalues": [float(v) for v in np.sort(np.asarray(evl, dtype=float)[:w["nBlock"]])],
                        "driver": "eigensolvers_b200.lanczos (mirror) with CudaVector.solveBlock",
                        "note": "the nBlock solves of a Krylov step share one pass over H and one fused "
                                "orthogonalisation launch per Arnoldi step; not the headline"}
            del opl
        except Exception as e:
            lockstep = {"error": f"{type(e).__name__}: {e}"}

    # ---- CPU baseline (rank 0, N = 1 only): a bounded continuous slice of the reference's run
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            n_mv = 40 if not feast else 24
            stamps, kind, threads = cpu_slice(w, n_mv + 2)
            per_mv = float(stamps[-1] - stamps[1]) / (len(stamps) - 2)
            cpu = {"value": per_mv * matvecs / n_eig, "unit": "s", "cores": threads, "kind": kind,
                   "sample": f"the first {len(stamps) - 2} operator applications of the run on the host (scipy csr_matvec serial, "
                             f"BLAS-1 on {threads} threads), {stamps[-1] - stamps[1]:.1f} s; extrapolated to the {matvecs} applications "
                             f"this GPU run needed; host has {os.cpu_count()} cores"}
        except Exception as e:  # the baseline must never take the headline down
            cpu = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        line = {
            "metric": "time_to_eConv_per_eigenpair", "value": value, "unit": "s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "c128" if feast else "f64",
            "data": "synthetic", "config": config_dict(args, w, nnz_total),
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
            "e2e": {"value": e2e_value, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "roofline": dominant, "roofline_spmv": roof_spmv, "roofline_arnoldi_step": roof_orth,
            "roofline_gram_schmidt": roof_gs, "cpu_baseline": cpu, "matrix_free": matrix_free, "lockstep": lockstep, "preconditioned": preconditioned,
            "result": {"driver": drv_name, "options": {k: v for k, v in opts["linearSystemArgs"].items()}, "transport": rt.transport, "format": fmt, "converged": converged, "eigenvalues": ev_out,
                       "cumIter": int(st.get("cumIter", st.get("outerIter", 0))),
                       "n_vectors_returned": len(Y), "lindep_abort": bool(np.any(np.isnan(ev_arr))),
                       "status_at_exit": {k: int(st[k]) for k in ("outerIter", "innerIter", "iBlock") if k in st},
                       "matvecs_per_step": int(matvecs), "solves_per_step": int(solves), "true_residual": true_res,
                       "profiled_step_s": t_prof, "step_share": share, "each_step_ms_rank0": step_ms,
                       "arnoldi_step_kernel_phases": orth_trace,
                       "solves_switched_to_safe_reorth": int(rt.stats.get("safe_solves", 0)),
                       "max_orthogonality_loss_seen": float(rt.stats.get("orth_loss", 0.0))},
        }
        if feast:
            iters = int(st.get("outerIter", 0)) + 1
            line["result"].update(feast_iterations=iters, seconds_per_feast_iteration=ms_per_step * 1e-3 / iters,
                                  window=[w["eMin"], w["eMax"]], analytic_levels_in_window=[
                                      float(x) for x in w["analytic"] if w["eMin"] < x < w["eMax"]],
                                  feast=feast_profile, feast_tasks=feast_tasks)
        print(json.dumps(line), flush=True)

    # ---- informational extras, AFTER the headline is out (stderr): GCROT recycling (SciPy's CU=)
    if args.extras and not feast:
        try:
            op3 = make_operator()
            ropts = {"linearSystemArgs": dict(opts["linearSystemArgs"], recycle=True)}
            one_run(op3, guess_dev, ropts)
            barrier()
            mv1 = rt.stats["matvecs"]
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            ev3, Y3, st3 = one_run(op3, guess_dev, ropts)
            r1.record()
            barrier()
            rec = {"gcrot_recycling": {"value": reduce_max(r0.elapsed_time(r1)) * 1e-3 / n_eig, "unit": "s",
                                       "matvecs": int(rt.stats["matvecs"] - mv1), "converged": bool(st3["isConverged"]),
                                       "eigenvalues": [float(x) for x in np.sort(ev3[:w["nBlock"]])]}}
        except Exception as e:
            rec = {"gcrot_recycling": {"error": f"{type(e).__name__}: {e}"}}
        if rank == 0:
            sys.stderr.write(json.dumps(rec) + "\n")
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--driver", default="auto", choices=["auto", "reference", "mirror"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--extras", action="store_true", help="informational GCROT-recycling leg after the headline (stderr)")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational matrix-free leg")
    ap.add_argument("--no-e2e", action="store_true", help="development: skip the end-to-end leg (e2e.value = null)")
    ap.add_argument("--no-profile", action="store_true", help="development: skip the profiled step (rooflines empty)")
    ap.add_argument("--feast-tasks", default=None, choices=["tasks", "dynamic"],
                    help="c5, N > 1: also time this distribution of the (node, vector) solves (informational)")
    ap.add_argument("--distribute", default="nodes", choices=["nodes", "tasks", "dynamic"],
                    help="c5, N > 1: distribution of the timed run (BASELINE config 5: nodes = one node per GPU)")
    ap.add_argument("--preconditioner", default=None, choices=["jacobi"],
                    help="development: run the workload with linearSystemArgs['preconditioner'] set (stated in result.options)")
    ap.add_argument("--cpu-budget", type=float, default=240.0, help="seconds of CPU sampling for --impl reference")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    try:
        if args.impl == "reference":
            if rank != 0:
                return
            run_reference_arm(args, build_workload(args.workload))
            return
        world = int(os.environ.get("WORLD_SIZE", "1"))
        run_ours(args, build_workload(args.workload, rank, world))
    except BaseException as e:  # surface the traceback (torchrun's elastic summary swallows it otherwise)
        if isinstance(e, SystemExit) and e.code in (0, None):
            raise
        tb = traceback.format_exc()
        sys.stderr.write(f"[bench.py rank {rank}] FAILED: {type(e).__name__}: {e}\n{tb}\n")
        sys.stderr.write(json.dumps({"bench_error": f"{type(e).__name__}: {e}", "rank": rank,
                                     "traceback": tb.replace("\n", " | ")}) + "\n")
        sys.stderr.flush()
        raise


if __name__ == "__main__":
    try:
        from torch.distributed.elastic.multiprocessing.errors import record
        main = record(main)
    except Exception:
        pass
    main()
