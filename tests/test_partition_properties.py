"""Randomised properties of the host side of the row-sharded mode, CPU only: the
native partition / halo-map / exchange-plan routines of libcudavec (csrc/comm.cu, no device call)
against the numpy oracle on arbitrary sparsity patterns, rank counts (including more ranks than
rows, i.e. ranks that own nothing) and band widths wider than a rank's block."""
import ctypes as C

import numpy as np
import scipy.sparse as sp
from hypothesis import HealthCheck, example, given, settings, strategies as st

from eigensolvers_b200 import _lib
from eigensolvers_b200 import partition as part
from oracle import partition_oracle as po

SETTINGS = dict(deadline=None, max_examples=60, suppress_health_check=[HealthCheck.too_slow], derandomize=True)


@st.composite
def sparse_matrices(draw):
    n = draw(st.integers(1, 120))
    density = draw(st.sampled_from([0.0, 0.01, 0.05, 0.3, 1.0]))
    seed = draw(st.integers(0, 2 ** 16))
    symmetric = draw(st.booleans())
    rng = np.random.default_rng(seed)
    k = int(density * n * n)
    rows, cols = rng.integers(0, n, k), rng.integers(0, n, k)
    A = sp.coo_matrix((rng.standard_normal(k), (rows, cols)), shape=(n, n)).tocsr()
    if symmetric:
        A = (A + A.T).tocsr()
    if draw(st.booleans()):
        A = (A + sp.identity(n)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def _dense_small(n):
    A = sp.csr_matrix(np.arange(1.0, n * n + 1).reshape(n, n))
    A.sort_indices()
    return A


@settings(**SETTINGS)
@given(sparse_matrices(), st.integers(1, 9))
@example(_dense_small(3), 8)      # more ranks than rows: five ranks own nothing
@example(_dense_small(1), 4)
def test_local_blocks_match_the_oracle(H, world):
    n = H.shape[0]
    off = part.row_offsets(n, world)
    np.testing.assert_array_equal(off, po.offsets(n, world))
    assert off[0] == 0 and off[-1] == n and np.all(np.diff(off) >= 0)
    assert np.diff(off).max() - np.diff(off).min() <= 1            # blocks differ by at most one row
    nnz_seen = 0
    for rank in range(world):
        got = part.local_block(H.indptr, H.indices, H.data, off, rank)
        ref = po.local_block(H, off, rank)
        for key in ("indptr", "indices", "data", "halo_cols", "halo_owner"):
            np.testing.assert_array_equal(got[key], ref[key], err_msg=f"world {world} rank {rank} {key}")
        assert got["n_local"] == ref["n_local"] and got["n_halo"] == ref["n_halo"]
        # halo columns: sorted, unique, outside the block, grouped by ascending owner
        hc, ho = got["halo_cols"], got["halo_owner"]
        assert np.all(np.diff(hc) > 0) and np.all(np.diff(ho) >= 0)
        assert np.all((hc < off[rank]) | (hc >= off[rank + 1]))
        assert np.all((off[ho] <= hc) & (hc < off[ho + 1]))
        nnz_seen += int(got["indptr"][-1])
    assert nnz_seen == H.nnz


@settings(**SETTINGS)
@given(sparse_matrices(), st.integers(2, 7), st.integers(0, 2 ** 16))
@example(_dense_small(3), 7, 1)
@example(_dense_small(2), 2, 1)
def test_emulated_sharded_spmv_equals_the_full_product(H, world, seed):
    """Every rank's pack list / receive offsets, emulated in one process: gathering what the plan says
    and multiplying the renumbered local block reproduces H @ x exactly (same summation order)."""
    n = H.shape[0]
    x = np.random.default_rng(seed).standard_normal(n)
    off = part.row_offsets(n, world)
    blocks = [part.local_block(H.indptr, H.indices, H.data, off, r) for r in range(world)]
    reqs = [part.requests_by_owner(b["halo_cols"], b["halo_owner"], off, world) for b in blocks]
    all_requests = [r[0] for r in reqs]
    sends = [part.send_lists(all_requests, p, world) for p in range(world)]
    y = np.empty(n)
    for r in range(world):
        b, recv_off = blocks[r], reqs[r][1]
        assert recv_off[0] == 0 and recv_off[-1] == b["n_halo"] and recv_off[r + 1] == recv_off[r]   # nothing from itself
        halo = np.full(b["n_halo"], np.nan)
        for p in range(world):
            send_idx, send_off = sends[p]
            seg = send_idx[send_off[r]:send_off[r + 1]]
            assert len(seg) == recv_off[p + 1] - recv_off[p]         # what p packs for r is what r expects from p
            assert np.all((seg >= 0) & (seg < off[p + 1] - off[p]))   # local row ids of the sender
            halo[recv_off[p]:recv_off[p + 1]] = x[off[p]:off[p + 1]][seg]
        assert not np.any(np.isnan(halo))                            # every halo slot was filled exactly once
        xl = np.concatenate([x[off[r]:off[r + 1]], halo])
        Hl = sp.csr_matrix((b["data"], b["indices"], b["indptr"]), shape=(b["n_local"], len(xl)))
        y[off[r]:off[r + 1]] = Hl @ xl
    np.testing.assert_array_equal(y, H @ x)


@settings(**SETTINGS)
@given(st.integers(1, 400), st.integers(1, 9), st.integers(0, 450), st.integers(0, 450))
@example(3, 8, 5, 5)          # more ranks than rows
@example(1, 2, 1, 1)
@example(64, 8, 64, 0)        # lower band spans every rank, no upper band
@example(100, 7, 0, 0)        # diagonal operator: nothing to exchange
def test_dia_band_halo_plan_matches_the_oracle(N, world, lo, hi):
    """cv_dia_halo_plan for arbitrary sizes: bands wider than a block (rows come from several ranks),
    ranks without rows, one-sided bands."""
    lib = _lib.load()
    off = part.row_offsets(N, world)
    cap = 4 * world + 4
    total_sent = total_recv = 0
    for rank in range(world):
        send5 = np.zeros(5 * cap, dtype=np.int64)
        recv4 = np.zeros(4 * cap, dtype=np.int64)
        ns, nr = C.c_int(), C.c_int()
        _lib.check(lib.cv_dia_halo_plan(off.ctypes.data, world, rank, lo, hi, cap, C.byref(ns), send5.ctypes.data,
                                        C.byref(nr), recv4.ctypes.data))
        send = sorted((tuple(int(v) for v in send5[5 * i:5 * i + 5]) for i in range(ns.value)), key=lambda t: (t[0], t[3]))
        recv = sorted((tuple(int(v) for v in recv4[4 * i:4 * i + 4]) for i in range(nr.value)), key=lambda t: (t[1], t[0]))
        ref_send, ref_recv = po.dia_halo_plan(off, rank, lo, hi)
        # zero-length ranges carry no information: compare the non-empty ones
        assert [s for s in send if s[2] > 0] == [s for s in ref_send if s[2] > 0], (rank, send, ref_send)
        assert [r for r in recv if r[3] > 0] == [r for r in ref_recv if r[3] > 0], (rank, recv, ref_recv)
        r0, r1 = int(off[rank]), int(off[rank + 1])
        for peer, first, count, band, slot in send:
            assert peer != rank and 0 <= first and first + count <= r1 - r0
        total_sent += sum(s[2] for s in send)
        total_recv += sum(r[3] for r in recv)
    assert total_sent == total_recv
