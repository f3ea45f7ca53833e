"""Row partition and halo maps: the native host routines (csrc/comm.cu) must equal
the scipy-slicing oracle bit for bit; the exchange plan is exercised with 2 gloo ranks on CPU,
where a numpy emulation of the sharded SpMV must reproduce H @ x exactly."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from eigensolvers_b200 import hamiltonians as hm
from eigensolvers_b200 import partition as part
from oracle import partition_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _matrices():
    yield "lap9", hm.laplacian3d(9)
    yield "osc", hm.coupled_oscillators((5, 4, 4, 3))[0]
    rng = np.random.default_rng(0)
    A = sp.random(301, 301, density=0.02, random_state=rng, format="csr")
    yield "rand", (A + A.T).tocsr()
    yield "diag", sp.identity(17, format="csr")
    yield "empty_rows", sp.csr_matrix(([1.0, 2.0], ([0, 9], [9, 0])), shape=(10, 10))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_and_halo_bit_exact(world):
    for name, H in _matrices():
        H = H.tocsr()
        H.sort_indices()
        n = H.shape[0]
        off = part.row_offsets(n, world)
        np.testing.assert_array_equal(off, po.offsets(n, world))
        for rank in range(world):
            got = part.local_block(H.indptr, H.indices, H.data, off, rank)
            ref = po.local_block(H, off, rank)
            for key in ("indptr", "indices", "data", "halo_cols", "halo_owner"):
                np.testing.assert_array_equal(got[key], ref[key], err_msg=f"{name} w{world} r{rank} {key}")
                assert got[key].dtype == ref[key].dtype, (name, key)
            assert got["n_local"] == ref["n_local"] and got["n_halo"] == ref["n_halo"]


def test_exchange_plan_single_process_emulation():
    """All ranks emulated in one process: sharded SpMV == H @ x (bit-exact, same summation order)."""
    for name, H in _matrices():
        H = H.tocsr()
        H.sort_indices()
        n = H.shape[0]
        x = np.random.default_rng(1).standard_normal(n)
        for world in (2, 3, 5):
            off = part.row_offsets(n, world)
            blocks = [part.local_block(H.indptr, H.indices, H.data, off, r) for r in range(world)]
            reqs = [part.requests_by_owner(b["halo_cols"], b["halo_owner"], off, world) for b in blocks]
            all_requests = [r[0] for r in reqs]
            y = np.empty(n)
            for r in range(world):
                b = blocks[r]
                recv_off = reqs[r][1]
                halo = np.empty(b["n_halo"])
                for p in range(world):
                    send_idx, send_off = part.send_lists(all_requests, p, world)
                    seg = send_idx[send_off[r]:send_off[r + 1]]
                    halo[recv_off[p]:recv_off[p + 1]] = x[off[p]:off[p + 1]][seg]
                xl = np.concatenate([x[off[r]:off[r + 1]], halo])
                Hl = sp.csr_matrix((b["data"], b["indices"], b["indptr"]), shape=(b["n_local"], len(xl)))
                y[off[r]:off[r + 1]] = Hl @ xl
            np.testing.assert_allclose(y, H @ x, rtol=1e-15, atol=1e-15, err_msg=f"{name} world {world}")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, n_side, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch
        H = hm.laplacian3d(n_side).tocsr()
        n = H.shape[0]
        x = np.random.default_rng(1).standard_normal(n)
        off = part.row_offsets(n, world)
        b = part.local_block(H.indptr, H.indices, H.data, off, rank)
        send_idx, send_off, recv_off = part.exchange_plan(b["halo_cols"], b["halo_owner"], off, rank, world)
        xl = x[off[rank]:off[rank + 1]]
        # halo exchange over gloo: the same pack -> send/recv -> halo-buffer schedule the GPU path runs
        halo = torch.empty(b["n_halo"], dtype=torch.float64)
        sendbuf = torch.from_numpy(np.ascontiguousarray(xl[send_idx]))
        ops = []
        for p in range(world):
            if p == rank:
                continue
            if send_off[p + 1] > send_off[p]:
                ops.append(dist.P2POp(dist.isend, sendbuf[send_off[p]:send_off[p + 1]], p))
            if recv_off[p + 1] > recv_off[p]:
                ops.append(dist.P2POp(dist.irecv, halo[recv_off[p]:recv_off[p + 1]], p))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        xe = np.concatenate([xl, halo.numpy()])
        Hl = sp.csr_matrix((b["data"], b["indices"], b["indptr"]), shape=(b["n_local"], len(xe)))
        y_loc = Hl @ xe
        # batched scalar all-reduce, as after every reduction kernel
        dots = torch.tensor([float(xl @ y_loc), float(y_loc @ y_loc)], dtype=torch.float64)
        dist.all_reduce(dots)
        y_ref = H @ x
        ok = np.allclose(y_loc, y_ref[off[rank]:off[rank + 1]], rtol=1e-15, atol=1e-15)
        ok &= np.allclose(dots.numpy(), [x @ y_ref, y_ref @ y_ref], rtol=1e-12)
        q.put((rank, bool(ok), int(b["n_halo"])))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_two_gloo_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, 8, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] for r in results), results
    assert all(r[2] == 64 for r in results)  # one 8x8 plane of halo on each side of the cut


def _feast_worker(rank, world, port, q, distribute="nodes"):
    import warnings
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=2)          # two workers share the host: do not oversubscribe BLAS
        from eigensolvers_b200.contour import feastDiagonalization
        from oracle.numpy_vector import NumpyVectorOracle as NV
        g = np.load(os.path.join(ROOT, "tests", "golden", "feast_t1.npz"))
        o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-2}}
        Y = [NV(g["Y1"][:, i].copy(), o) for i in range(6)]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ev, vecs, st = feastDiagonalization(g["A"], Y, 8, "legendre", 160.0, 166.0, 1e-10, 20, writeOut=False,
                                                distribute=distribute)
        inside = np.sort([e for e in ev if 160.0 <= e <= 166.0])
        ref = np.sort([e for e in g["ev"] if 160.0 <= e <= 166.0])
        q.put((rank, len(inside) == len(ref) and bool(np.allclose(inside, ref, rtol=0, atol=1e-6)), [float(x) for x in inside]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("distribute", ["nodes", "tasks", "dynamic"])
def test_feast_nodes_distributed_over_two_gloo_ranks(distribute):
    """FEAST with the quadrature nodes (or the individual (node, vector) solves, balanced by measured
    cost) split over 2 ranks (H replicated, one all-reduce of the m0 accumulated vectors per
    iteration) finds the same eigenvalues as the reference's serial run."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_feast_worker, args=(r, 2, port, q, distribute)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in results), results
    assert results[0][2] == results[1][2]        # both ranks hold identical Ritz values


def test_dia_offset_table_host_logic():
    """DeviceOperator._dia_table (host side of the DIA format decision): the offsets col-row of the
    banded generators, rejection of unstructured / too-padded matrices, row-block invariance."""
    import types
    from eigensolvers_b200 import hamiltonians as hm
    from eigensolvers_b200.operator import DeviceOperator

    def table(A, row0=0, rows=None, force=False):
        A = A.tocsr()
        A.sort_indices()
        if rows is not None:
            A = A[rows[0]:rows[1]]
            row0 = rows[0]
        op = DeviceOperator.__new__(DeviceOperator)
        op.rt = types.SimpleNamespace(world=1)
        op.nnz = int(A.nnz)
        return op._dia_table(A.indptr.astype(np.int64), A.indices.astype(np.int32), row0, force=force)

    lap = hm.laplacian3d(12)
    assert table(lap) == [-144, -12, -1, 0, 1, 12, 144]
    H = hm.coupled_oscillators((12, 10, 10))[0]
    offs = table(H)
    strides = (100, 10, 1)
    expect = {0} | {a * strides[i] + b * strides[i + 1] for i in range(2) for a in (-1, 1) for b in (-1, 1)}
    assert offs == sorted(expect)
    # a row block of the same matrix sees (a subset of) the same offsets, relative to GLOBAL rows
    blk = table(H, rows=(300, 900), force=True)
    assert set(blk) <= set(offs) and 0 in blk
    # unstructured sparsity: more than 64 distinct offsets -> no DIA
    rng = np.random.default_rng(0)
    R = sp.random(3000, 3000, density=0.004, random_state=rng, format="csr")
    assert table(R + R.T) is None
    # heavy basis-edge truncation: padding would cost more than CSR's index stream
    assert table(hm.coupled_oscillators((6, 5, 4, 4, 3))[0]) is None
    # tiny matrices are never converted
    assert table(hm.laplacian3d(5)) is None


@pytest.mark.parametrize("N,world,lo,hi", [(1000, 2, 110, 110), (1000, 8, 110, 110), (1000, 8, 300, 40),
                                            (97, 3, 5, 0), (64, 8, 20, 20), (20_000_000, 8, 1_100_000, 1_100_000)])
def test_dia_halo_plan_bit_exact(N, world, lo, hi):
    """The DIA band-halo exchange plan (cv_dia_halo_plan, shared with cv_op_set_dia_halo) against the
    numpy statement: which contiguous row ranges every rank sends / receives, where they land."""
    import ctypes as C
    from eigensolvers_b200 import _lib
    lib = _lib.load()
    off = part.row_offsets(N, world)
    np.testing.assert_array_equal(off, po.offsets(N, world))
    cap = 4 * world
    for rank in range(world):
        send5 = np.zeros(5 * cap, dtype=np.int64)
        recv4 = np.zeros(4 * cap, dtype=np.int64)
        ns, nr = C.c_int(), C.c_int()
        _lib.check(lib.cv_dia_halo_plan(off.ctypes.data, world, rank, lo, hi, cap, C.byref(ns), send5.ctypes.data,
                                        C.byref(nr), recv4.ctypes.data))
        send = sorted((tuple(int(v) for v in send5[5 * i:5 * i + 5]) for i in range(ns.value)), key=lambda t: (t[0], t[3]))
        recv = sorted((tuple(int(v) for v in recv4[4 * i:4 * i + 4]) for i in range(nr.value)), key=lambda t: (t[1], t[0]))
        ref_send, ref_recv = po.dia_halo_plan(off, rank, lo, hi)
        assert send == ref_send, (rank, send, ref_send)
        assert recv == ref_recv, (rank, recv, ref_recv)
        # every row of both bands inside [0, N) is covered exactly once
        r0, r1 = int(off[rank]), int(off[rank + 1])
        assert sum(c for p, b, s, c in recv if b == 0) == r0 - max(r0 - lo, 0)
        assert sum(c for p, b, s, c in recv if b == 1) == min(r1 + hi, N) - r1


def test_feast_task_assignment_is_balanced_and_deterministic():
    """contour._assign_tasks: every (node, vector) solve is owned by exactly one rank, all ranks
    compute the same map, and measured costs balance the load (longest processing time first)."""
    from eigensolvers_b200.contour import _assign_tasks
    nodes = [(0.1 * k, complex(1.0, 0.5 / (k + 1))) for k in range(8)]
    m0, world = 6, 8
    tasks = {(k, i) for k in range(8) for i in range(m0)}
    for cost in (None, {(k, i): 1.0 + k for (k, i) in tasks}):
        owned = [_assign_tasks("tasks", r, world, nodes, m0, cost) for r in range(world)]
        assert set().union(*owned) == tasks and sum(len(o) for o in owned) == len(tasks)
        c = cost or {(k, i): 1.0 / abs(nodes[k][1].imag) for (k, i) in tasks}
        loads = [sum(c[t] for t in o) for o in owned]
        by_node = [sum(c[t] for t in _assign_tasks("nodes", r, world, nodes, m0, cost)) for r in range(world)]
        assert max(loads) <= 1.1 * sum(loads) / world        # LPT: within 10 % of the mean here
        assert max(loads) < max(by_node)                     # whole nodes per rank are worse
    assert _assign_tasks("tasks", 0, 1, nodes, m0, None) == tasks
