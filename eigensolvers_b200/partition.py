"""Row partition and halo-exchange plan of the row-sharded mode.

Integer work only.  The partition and the halo maps themselves are computed by the native host
routines ``cv_partition_rows`` / ``cv_halo_build`` (csrc/comm.cu); this module wraps them for
host arrays and derives the send lists by exchanging the request lists once over
``torch.distributed`` (gloo on CPU in the tests, NCCL on the GPUs).  The reference has no
partitioner; the oracle is scipy row slicing + ``np.unique`` (oracle/partition_oracle.py).
"""
import ctypes as C

import numpy as np

from . import _lib


def row_offsets(n, world):
    """offsets[p] = floor(p*n/world): rank p owns rows [offsets[p], offsets[p+1])."""
    lib = _lib.load()
    off = np.empty(world + 1, dtype=np.int64)
    _lib.check(lib.cv_partition_rows(int(n), int(world), off.ctypes.data_as(C.POINTER(C.c_int64))))
    return off


def local_block(indptr, indices, data, offsets, rank):
    """Local CSR of `rank` with columns renumbered [owned | halo], plus the halo description.

    Returns dict(indptr, indices, data, halo_cols, halo_owner, n_local, n_halo, row0).
    """
    lib = _lib.load()
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    world = len(offsets) - 1
    r0, r1 = int(offsets[rank]), int(offsets[rank + 1])
    n_halo = C.c_int64()
    _lib.check(lib.cv_halo_count(indptr.ctypes.data, indices.ctypes.data, r0, r1, C.byref(n_halo)))
    nh = n_halo.value
    nnz_loc = int(indptr[r1] - indptr[r0])
    halo_cols = np.empty(nh, dtype=np.int32)
    halo_owner = np.empty(nh, dtype=np.int32)
    loc_indptr = np.empty(r1 - r0 + 1, dtype=np.int64)
    loc_indices = np.empty(nnz_loc, dtype=np.int32)
    _lib.check(lib.cv_halo_build(indptr.ctypes.data, indices.ctypes.data, r0, r1, offsets.ctypes.data,
                                 world, nh, halo_cols.ctypes.data, halo_owner.ctypes.data,
                                 loc_indptr.ctypes.data, loc_indices.ctypes.data))
    loc_data = None if data is None else np.ascontiguousarray(data[indptr[r0]:indptr[r1]])
    return dict(indptr=loc_indptr, indices=loc_indices, data=loc_data, halo_cols=halo_cols,
                halo_owner=halo_owner, n_local=r1 - r0, n_halo=nh, row0=r0)


def requests_by_owner(halo_cols, halo_owner, offsets, world):
    """For each owner p: the LOCAL row ids (at p) this rank needs, and the receive offsets.

    The halo is sorted by global column, hence grouped by owner in ascending rank order, so the
    entries received from p land in halo slots [recv_off[p], recv_off[p+1]).
    """
    counts = np.bincount(halo_owner, minlength=world).astype(np.int64)
    recv_off = np.zeros(world + 1, dtype=np.int64)
    np.cumsum(counts, out=recv_off[1:])
    reqs = [(halo_cols[recv_off[p]:recv_off[p + 1]].astype(np.int64) - int(offsets[p])).astype(np.int32)
            for p in range(world)]
    return reqs, recv_off


def send_lists(all_requests, rank, world):
    """all_requests[q][p] = rows of p that q needs.  Returns this rank's (send_idx, send_off)."""
    parts = [np.asarray(all_requests[q][rank], dtype=np.int32) for q in range(world)]
    send_off = np.zeros(world + 1, dtype=np.int64)
    np.cumsum([len(p) for p in parts], out=send_off[1:])
    send_idx = np.concatenate(parts) if send_off[-1] else np.empty(0, dtype=np.int32)
    return send_idx.astype(np.int32), send_off


def exchange_plan(halo_cols, halo_owner, offsets, rank, world):
    """Collective: every rank learns which of its rows each peer needs."""
    import torch.distributed as dist
    reqs, recv_off = requests_by_owner(halo_cols, halo_owner, offsets, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, reqs)
    send_idx, send_off = send_lists(gathered, rank, world)
    return send_idx, send_off, recv_off
