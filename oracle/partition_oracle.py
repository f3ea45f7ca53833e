"""Oracle for the row partition and halo maps: scipy row slicing + np.unique +
np.searchsorted.  TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference has no
partitioner; this is the independent statement the native routines must match bit for bit."""
import numpy as np
import scipy.sparse as sp


def offsets(n, world):
    return np.array([(p * n) // world for p in range(world + 1)], dtype=np.int64)


def local_block(H, off, rank):
    H = sp.csr_matrix(H)
    H.sort_indices()
    r0, r1 = int(off[rank]), int(off[rank + 1])
    Hp = H[r0:r1]
    cols = Hp.indices.astype(np.int64)
    outside = (cols < r0) | (cols >= r1)
    halo = np.unique(cols[outside])
    owner = np.searchsorted(off, halo, side="right") - 1
    local = np.where(outside, (r1 - r0) + np.searchsorted(halo, cols), cols - r0)
    return dict(indptr=Hp.indptr.astype(np.int64), indices=local.astype(np.int32), data=Hp.data.copy(),
                halo_cols=halo.astype(np.int32), halo_owner=owner.astype(np.int32),
                n_local=r1 - r0, n_halo=len(halo), row0=r0)


def dia_halo_plan(off, rank, lo_len, hi_len):
    """Band halo of the DIA format, stated independently with numpy: the rows below / above the
    block that `rank` needs, grouped by owner (np.searchsorted), and — by symmetry of the same
    statement applied to every other rank — the rows it has to send.
    Returns (send, recv): lists of tuples (peer, first local row, count, band, slot) and
    (peer, band, first slot, count)."""
    off = np.asarray(off, dtype=np.int64)
    P, N = len(off) - 1, int(off[-1])

    def needs(q):
        r0, r1 = int(off[q]), int(off[q + 1])
        out = []
        for band, rows, base in ((0, np.arange(max(r0 - lo_len, 0), r0), r0 - lo_len),
                                 (1, np.arange(r1, min(r1 + hi_len, N)), r1)):
            if len(rows) == 0:
                continue
            owner = np.searchsorted(off, rows, side="right") - 1
            for p in np.unique(owner):
                sel = rows[owner == p]
                if p == q:
                    continue
                assert np.array_equal(sel, np.arange(sel[0], sel[0] + len(sel)))
                out.append((int(p), band, int(sel[0]), len(sel), int(sel[0] - base)))
        return out          # (owner, band, first global row, count, slot)

    recv = sorted(((p, band, slot, cnt) for p, band, g0, cnt, slot in needs(rank)), key=lambda t: (t[1], t[0]))
    send = []
    for q in range(P):
        if q == rank:
            continue
        for p, band, g0, cnt, slot in needs(q):
            if p == rank:
                send.append((q, g0 - int(off[rank]), cnt, band, slot))
    send.sort(key=lambda t: (t[0], t[3]))
    return send, recv
