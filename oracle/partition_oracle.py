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
