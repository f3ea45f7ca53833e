"""Index logic of the fused Arnoldi-step kernels, checked on the CPU.

k_orth_step / k_orth_step_batch (csrc/kernels_orth.cuh, kernels_batch.cuh) split their dot phase
into slabs of basis vectors reduced by disjoint CTA ranges, and the lock-step kernel gives every
problem a CTA range for its update phase.  The functions that compute those splits are
__host__ __device__; libcudavec exports them as pure host routines (cv_orth_slab_plan,
cv_orth_batch_plan).  A vector that falls into no slab, or a slab without a CTA, would silently
drop a Hessenberg entry -- so the invariants are checked exhaustively over every basis size the
solvers can produce and over grid sizes from the minimum to beyond a full B200 wave."""
import ctypes as C
import itertools

import numpy as np
import pytest

from eigensolvers_b200 import _lib

MAX_PTRS = 128      # CV_MAX_PTRS
BATCH_PTRS = 64     # CV_BATCH_PTRS
MAX_BATCH = 4       # CV_MAX_BATCH


def slab_plan(m, grid, cplx, mode):
    lib = _lib.load()
    ny = C.c_int()
    start = np.zeros(MAX_PTRS // 8 + 2, dtype=np.int32)
    i0 = np.zeros(MAX_PTRS // 8 + 2, dtype=np.int32)
    _lib.check(lib.cv_orth_slab_plan(m, grid, cplx, mode, C.byref(ny), start.ctypes.data, i0.ctypes.data))
    return ny.value, start[:ny.value + 1].copy(), i0[:ny.value + 1].copy()


GRIDS = sorted(set(list(range(1, 40)) + [64, 147, 148, 149, 295, 296, 297, 443, 444, 445, 592, 1184, 4096]))


@pytest.mark.parametrize("cplx", [0, 1])
@pytest.mark.parametrize("mode", [0, 1])
def test_slab_plan_invariants(cplx, mode):
    MI = 8 if cplx else 16                                  # accumulators per thread (ORTH_MI)
    for m in range(1, MAX_PTRS + 1):
        ny_min = -(-m // MI)
        for grid in GRIDS:
            if grid < ny_min:
                continue
            ny, start, i0 = slab_plan(m, grid, cplx, mode)
            assert ny == ny_min
            sizes = np.diff(i0)
            assert i0[0] == 0 and i0[-1] == m and np.all(sizes >= 1) and np.all(sizes <= MI), (m, grid, i0)
            if mode == 1:
                assert sizes.max() - sizes.min() <= 1, (m, sizes)      # even slabs (33 -> 11 + 11 + 11)
            else:
                assert np.all(sizes[:-1] == MI)                         # full slabs + remainder
            ctas = np.diff(start)
            assert start[0] == 0 and start[-1] == grid and np.all(ctas >= 1), (m, grid, start)
            if grid >= 8 * ny:
                # shares proportional to the loads per row (mi + 1), up to rounding; the last slab takes the rest
                ideal = grid * (sizes + 1) / float(m + ny)
                assert np.all(np.abs(ctas[:-1] - ideal[:-1]) <= 1.0), (m, grid, ctas, ideal)
                assert abs(ctas[-1] - ideal[-1]) <= ny, (m, grid, ctas, ideal)


def test_slab_plan_refuses_a_grid_smaller_than_the_slab_count():
    lib = _lib.load()
    ny = C.c_int()
    buf = np.zeros(32, dtype=np.int32)
    assert lib.cv_orth_slab_plan(40, 2, 0, 1, C.byref(ny), buf.ctypes.data, buf.ctypes.data) != 0   # 3 slabs, 2 CTAs
    assert lib.cv_orth_slab_plan(0, 8, 0, 1, C.byref(ny), buf.ctypes.data, buf.ctypes.data) != 0
    assert lib.cv_orth_slab_plan(MAX_PTRS + 1, 64, 0, 1, C.byref(ny), buf.ctypes.data, buf.ctypes.data) != 0


def batch_plan(ms, mask, grid, cplx):
    lib = _lib.load()
    n = len(ms)
    marr = np.asarray(ms, dtype=np.int32)
    nslab = C.c_int()
    cap = MAX_BATCH * (BATCH_PTRS // 8)
    q, i0, mi = (np.zeros(cap, dtype=np.int32) for _ in range(3))
    start = np.zeros(cap + 1, dtype=np.int32)
    pstart = np.zeros(MAX_BATCH + 1, dtype=np.int32)
    _lib.check(lib.cv_orth_batch_plan(n, marr.ctypes.data, mask, grid, cplx, C.byref(nslab), q.ctypes.data, i0.ctypes.data,
                                      mi.ctypes.data, start.ctypes.data, pstart.ctypes.data))
    k = nslab.value
    return q[:k].copy(), i0[:k].copy(), mi[:k].copy(), start[:k + 1].copy(), pstart[:n + 1].copy()


@pytest.mark.parametrize("cplx", [0, 1])
def test_batch_plan_invariants(cplx):
    MI = 8 if cplx else 16
    rng = np.random.default_rng(5)
    cases = []
    for nprob in range(1, MAX_BATCH + 1):
        for ms in itertools.product((1, 7, 16, 17, 33, 62, 64), repeat=nprob):
            cases.append(ms)
    cases = [cases[i] for i in rng.permutation(len(cases))[:400]] + [(64, 64, 64, 64), (1, 1, 1, 1), (62,), (1,)]
    for ms in cases:
        nprob = len(ms)
        for mask in range(1, 1 << nprob):
            active = [p for p in range(nprob) if (mask >> p) & 1]
            nslabs = sum(-(-ms[p] // MI) for p in active)
            for grid in (max(nslabs, len(active)), max(nslabs, len(active)) + 1, 148, 444, 512):
                if grid < max(nslabs, len(active)):
                    continue
                q, i0, mi, start, pstart = batch_plan(ms, mask, grid, cplx)
                assert len(q) == nslabs
                # phase A: every vector of every active problem in exactly one slab, slabs in problem order
                for p in active:
                    sel = q == p
                    assert sel.sum() == -(-ms[p] // MI)
                    cover = np.concatenate([np.arange(a, a + b) for a, b in zip(i0[sel], mi[sel])])
                    np.testing.assert_array_equal(cover, np.arange(ms[p]))
                    assert mi[sel].max() - mi[sel].min() <= 1 and mi[sel].max() <= MI
                assert set(q.tolist()) == set(active)
                assert start[0] == 0 and start[-1] == grid and np.all(np.diff(start) >= 1), (ms, mask, grid, start)
                # phase B: active problems tile the grid with at least one CTA each, inactive ones get none
                width = np.diff(pstart)
                assert pstart[0] == 0 and pstart[-1] == grid, (ms, mask, grid, pstart)
                for p in range(nprob):
                    if p in active:
                        assert width[p] >= 1, (ms, mask, grid, pstart)
                    elif p < max(active):
                        assert width[p] == 0, (ms, mask, grid, pstart)
                # an inactive problem after the last active one ends at `grid` by construction (empty too
                # unless it is the final entry, which only closes the table)
                if grid >= 8 * nslabs:
                    ideal = grid * np.array([ms[p] + 2 for p in active]) / float(sum(ms[p] + 2 for p in active))
                    got = np.array([width[p] for p in active])
                    assert np.all(np.abs(got[:-1] - ideal[:-1]) <= 1.0), (ms, mask, grid, got, ideal)


def test_batch_plan_refuses_too_small_grids():
    lib = _lib.load()
    marr = np.asarray([62, 62], dtype=np.int32)
    nslab = C.c_int()
    buf = np.zeros(64, dtype=np.int32)
    args = (C.byref(nslab), buf.ctypes.data, buf.ctypes.data, buf.ctypes.data, buf.ctypes.data, buf.ctypes.data)
    assert lib.cv_orth_batch_plan(2, marr.ctypes.data, 3, 7, 0, *args) != 0          # 8 slabs, 7 CTAs
    assert lib.cv_orth_batch_plan(2, marr.ctypes.data, 0, 64, 0, *args) != 0         # nobody active
    assert lib.cv_orth_batch_plan(5, marr.ctypes.data, 1, 64, 0, *args) != 0         # more problems than a launch takes
