"""Device-resident Hamiltonian: CSR uploaded once, optionally re-laid out as sliced ELL
(32-row slices, column-major inside a slice) by a device kernel.  Replaces the host objects the
reference applies with ``H @ x`` (numpyVector.py:100,152,154): scipy.sparse matrices and dense
ndarrays (stored as CSR with every entry).

Layout in HBM (single GPU, N rows, nnz non-zeros, P padded non-zeros):
    indptr  int64[N+1]   indices int32[nnz]   data float64[nnz]          (CSR)
    slice_ptr int64[N/32+1]   sell_col int32[P]   sell_val float64[P]    (SELL-32x2: 32-row slices, even widths,
                                                                          entry (lane, 2p+e) at base + (p*32+lane)*2 + e)
In row-sharded mode each rank holds rows [r_p, r_{p+1}) with columns renumbered to
[0, n_loc) (owned) ++ [n_loc, n_loc + n_halo) (halo, sorted by global column = grouped by owner).
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib


class DeviceOperator:
    """Handle to a cv_op plus the torch tensors that own its device arrays."""

    dtype = np.dtype(np.float64)

    def __init__(self, runtime, shape):
        self.rt = runtime
        self.shape = tuple(int(s) for s in shape)  # GLOBAL shape
        self.handle = C.c_void_p()
        self._keep = []  # tensors owning device memory
        self._peer_owned = []  # this rank's CUDA-IPC exported halo buffers (peer-memory transport)
        self.nnz = 0
        self.padded_nnz = 0
        self.n_local = self.shape[0]
        self.n_halo = 0
        self.format = "csr"
        self.halo_cols = None
        self.send_idx = None
        self._diag_host = None    # this rank's diagonal entries H_ii (host float64), for the Jacobi preconditioner
        self._dinv = {}           # (sigma, reverse, cplx) -> device tensor 1/(sigma - H_ii)

    def __del__(self):
        try:
            if self.handle:
                self.rt.lib.cv_op_destroy(self.handle)
                self.handle = C.c_void_p()
            for own in self._peer_owned:
                self.rt.peer_release(own)
            self._peer_owned = []
        except Exception:
            pass

    # ---------------------------------------------------------------------------------------
    @staticmethod
    def _as_csr(H):
        if sp.issparse(H):
            A = H.tocsr()
        elif isinstance(H, np.ndarray):
            if H.ndim != 2:
                raise ValueError("operator must be a 2-D matrix")
            A = sp.csr_matrix(H)
        else:
            raise TypeError(
                f"CudaVector cannot apply an operator of type {type(H).__name__}: pass a "
                "scipy.sparse matrix, a dense ndarray or a DeviceOperator (an opaque "
                "LinearOperator has no device representation)")
        if np.iscomplexobj(A) and not np.any(A.data.imag):
            A = A.real                      # complex dtype, real values: the fast real-valued formats apply
        if A.shape[1] >= 2 ** 31:
            raise ValueError("column indices must fit int32")
        if not A.has_sorted_indices:
            A = A.sorted_indices()
        return A

    @classmethod
    def from_host(cls, H, runtime=None, fmt="auto"):
        """Upload (this rank's row block of) ``H``.  ``fmt``: 'auto' | 'csr' | 'sell' | 'dia'."""
        from .runtime import Runtime
        rt = runtime or Runtime.get()
        A = cls._as_csr(H)
        op = cls(rt, A.shape)
        indptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(A.indices, dtype=np.int32)
        zvalued = np.iscomplexobj(A)
        if zvalued:
            # complex (Hermitian) H, numpyVector.py:98-100 takes any `other @ array`: the real and the
            # imaginary parts travel as two value streams over ONE sparsity pattern, in CSR storage
            if fmt not in ("auto", "csr"):
                raise NotImplementedError(f"a complex-valued operator is stored as CSR, not {fmt!r}")
            fmt = "csr"
            op.dtype = np.dtype(np.complex128)
            op._data_im_host = np.ascontiguousarray(A.data.imag, dtype=np.float64)
        data = np.ascontiguousarray(A.data.real if zvalued else A.data, dtype=np.float64)
        diag = A.diagonal()
        if np.iscomplexobj(diag) and not np.any(diag.imag):
            diag = diag.real                # Hermitian: real diagonal
        if rt.world == 1:
            op._diag_host = np.asarray(diag, dtype=np.complex128 if np.iscomplexobj(diag) else np.float64)
            return op._finish(indptr, indices, data, fmt, indices, 0)
        r0, r1 = rt.local_range(A.shape[0])
        op._diag_host = np.asarray(diag[r0:r1], dtype=np.complex128 if np.iscomplexobj(diag) else np.float64)
        lo, hi = int(indptr[r0]), int(indptr[r1])
        if zvalued:
            op._data_im_host = op._data_im_host[lo:hi]
        return op._finish_sharded(np.ascontiguousarray(indptr[r0:r1 + 1] - lo), indices[lo:hi], data[lo:hi], r0, r1, fmt)

    def _finish_sharded(self, indptr, gcols, data, r0, r1, fmt):
        """Row-sharded construction from this rank's block (indptr rebased to 0, GLOBAL column ids).
        Banded operators go straight to DIA storage: its halo is the contiguous band below/above
        the block, so the general plan (sorted unique halo columns, renumbered local CSR, gather
        lists exchanged between ranks — the expensive host part) is never built for them."""
        rt = self.rt
        self.n_local, self.row0 = r1 - r0, r0
        self.nnz = int(indptr[-1])
        if fmt in ("auto", "dia") and self.shape[0] == self.shape[1]:
            table = self._dia_table(indptr, gcols, r0, force=(fmt == "dia"))
            if table is not None and self._build_dia_direct(indptr, gcols, data, r0, table):
                return self
            if fmt == "dia":
                raise ValueError("operator is not banded enough for DIA storage")
        loc_indptr, loc_indices, loc_data = self._localize(indptr, gcols, data, r0, r1, full=False)
        return self._finish(loc_indptr, loc_indices, loc_data, "sell" if fmt == "sell" else ("csr" if fmt == "csr" else "auto_nodia"),
                            gcols, r0)

    def _finish(self, indptr, indices, data, fmt, gcols=None, row0=0):
        """Upload the (local) CSR arrays, create the cv_op, register the halo plan, then pick the
        storage format: DIA (banded structure) > SELL-32 (short regular rows) > CSR."""
        op, rt = self, self.rt
        n_rows = len(indptr) - 1
        n_cols = op.shape[1] if rt.world == 1 else op.n_local + op.n_halo
        op.nnz = int(indptr[-1])
        t = rt.torch
        d_indptr = rt.upload(indptr)
        d_indices = rt.upload(indices) if op.nnz else t.empty(0, dtype=t.int32, device=rt.device)
        d_data = rt.upload(data) if op.nnz else t.empty(0, dtype=t.float64, device=rt.device)
        op._keep += [d_indptr, d_indices, d_data]
        _lib.check(rt.lib.cv_op_create_csr(rt.ctx, n_rows, n_cols, op.nnz, d_indptr.data_ptr(),
                                           d_indices.data_ptr(), d_data.data_ptr(),
                                           C.byref(op.handle)))
        im = getattr(op, "_data_im_host", None)
        if im is not None:                 # same entry order as `data` (the halo renumbering keeps it)
            assert len(im) == op.nnz
            d_im = rt.upload(im) if op.nnz else t.empty(0, dtype=t.float64, device=rt.device)
            op._keep.append(d_im)
            _lib.check(rt.lib.cv_op_set_imag(rt.ctx, op.handle, d_im.data_ptr()))
            op._data_im_host = None
        if rt.world > 1:
            op._register_halo()
        if fmt in ("auto", "dia") and rt.world == 1 and op.shape[0] == op.shape[1]:
            table = op._dia_table(indptr, indices, 0, force=(fmt == "dia"))
            if table is not None and op._attach_dia(table, None, 0):
                return op
            if fmt == "dia":
                raise ValueError("operator is not banded enough for DIA storage")
        if fmt in ("auto", "auto_nodia", "sell") and n_rows > 0 and op.nnz > 0:
            op._build_sell(force=(fmt == "sell"))
        return op

    # ---------------------------------------------------------------------------------------
    MAX_DIAG = 64

    def _dia_table(self, indptr, gcols, row0, force=False):
        """Offsets col-row of a diagonal layout, or None.  The table comes from a sample of rows on
        the host (<= 64 distinct offsets: stencil and product-basis Hamiltonians); the device fill
        kernel later verifies EVERY entry.  Collective: all ranks take the same decision (the table
        is the union over ranks)."""
        rt = self.rt
        n_rows = len(indptr) - 1
        if n_rows == 0 and rt.world == 1:
            return None
        offs = np.empty(0, np.int64)
        if n_rows:
            sample = np.unique(np.concatenate([
                np.arange(0, min(n_rows, 2048)), np.arange(max(n_rows - 2048, 0), n_rows),
                np.linspace(0, n_rows - 1, num=min(n_rows, 4096), dtype=np.int64)]))
            starts, stops = indptr[sample], indptr[sample + 1]
            lens = (stops - starts).astype(np.int64)
            if lens.sum():
                pos = np.repeat(starts - np.concatenate(([0], np.cumsum(lens)[:-1])), lens) + np.arange(lens.sum())
                rows = np.repeat(sample + row0, lens)
                offs = np.unique(np.asarray(gcols)[pos].astype(np.int64) - rows)
                if len(offs) > self.MAX_DIAG:
                    offs = offs[:self.MAX_DIAG + 1]
        info = (offs.tolist(), int(self.nnz), int(n_rows))
        if rt.world > 1:
            import torch.distributed as dist
            box = [None] * rt.world
            dist.all_gather_object(box, info)
        else:
            box = [info]
        table = sorted(set().union(*[set(b[0]) for b in box]))
        nnz_all, rows_all = sum(b[1] for b in box), sum(b[2] for b in box)
        D = len(table)
        if D == 0 or D > self.MAX_DIAG:
            return None
        # 8 bytes per stored slot against 12 per CSR entry; tiny matrices are not worth a format
        if not force and (8.0 * D * rows_all > 0.95 * 12.0 * nnz_all or rows_all < 1024):
            return None
        if max(abs(table[0]), abs(table[-1])) >= 2 ** 31 - 1:
            return None
        return table

    def _attach_dia(self, table, d_gcols, row0):
        """Scatter the CSR values into the diagonal layout on the device.  Returns False (all ranks
        alike) if some entry is off the table; the operator then keeps its previous format."""
        rt, t = self.rt, self.rt.torch
        n_rows = self.n_local
        D = len(table)
        ld = (n_rows + 31) // 32 * 32
        d_val = t.zeros(max(D * ld, 1), dtype=t.float64, device=rt.device)
        offs_arr = np.ascontiguousarray(table, dtype=np.int32)
        ok = C.c_int(0)
        _lib.check(rt.lib.cv_op_attach_dia(rt.ctx, self.handle, D, offs_arr.ctypes.data,
                                           None if d_gcols is None else d_gcols.data_ptr(), int(row0),
                                           d_val.data_ptr(), ld, C.byref(ok), rt.stream))
        good = bool(ok.value)
        if rt.world > 1:
            import torch.distributed as dist
            flags = [None] * rt.world
            dist.all_gather_object(flags, good)
            good = all(flags)
        if not good:
            _lib.check(rt.lib.cv_op_set_format(self.handle, _lib.CV_FMT_CSR))
            return False
        self._keep.append(d_val)
        self.dia_offsets = offs_arr
        self.padded_nnz = D * n_rows
        self.format = "dia"
        if rt.world > 1:
            self._setup_band_halo(max(0, -int(table[0])), max(0, int(table[-1])))
        return True

    def _setup_band_halo(self, lo_len, hi_len):
        """Row-sharded banded operator (DIA storage or the matrix-free Kronecker form): the x entries
        below / above the owned block arrive in two contiguous band buffers; with the peer transport
        the neighbours push their boundary rows straight into them over NVLink."""
        rt, t = self.rt, self.rt.torch
        halo_lo = t.zeros(2 * max(lo_len, 1), dtype=t.float64, device=rt.device)
        halo_hi = t.zeros(2 * max(hi_len, 1), dtype=t.float64, device=rt.device)
        self._keep += [halo_lo, halo_hi]
        off = rt.offsets_for(self.shape[0])
        _lib.check(rt.lib.cv_op_set_dia_halo(rt.ctx, self.handle, off.ctypes.data, halo_lo.data_ptr(),
                                             halo_hi.data_ptr()))
        r0, r1 = int(off[rt.rank]), int(off[rt.rank + 1])
        self.n_halo = min(lo_len, r0) + min(hi_len, self.shape[0] - r1)   # band rows held by other ranks
        if rt.transport == "peer":
            ptrs, _ = rt.peer_shared_alloc(rt.lib.cv_op_dia_halo_bytes(self.handle))
            if ptrs is not None:
                self._peer_owned.append(ptrs[rt.rank])
                arr = (C.c_void_p * rt.world)(*ptrs)
                _lib.check(rt.lib.cv_op_set_dia_halo_peers(rt.ctx, self.handle,
                                                           C.cast(arr, C.POINTER(C.c_void_p))))

    def _build_dia_direct(self, indptr, gcols, data, r0, table):
        """Sharded DIA without the general halo plan: the CSR arrays on the device keep GLOBAL
        column ids (only the fill kernel reads them), so CSR/SELL are not available on this handle."""
        rt, t = self.rt, self.rt.torch
        n_rows = len(indptr) - 1
        d_indptr = rt.upload(indptr)
        d_gcols = rt.upload(np.ascontiguousarray(gcols, dtype=np.int32)) if self.nnz else t.empty(0, dtype=t.int32, device=rt.device)
        d_data = rt.upload(data) if self.nnz else t.empty(0, dtype=t.float64, device=rt.device)
        _lib.check(rt.lib.cv_op_create_csr(rt.ctx, n_rows, n_rows, self.nnz, d_indptr.data_ptr(),
                                           d_gcols.data_ptr(), d_data.data_ptr(), C.byref(self.handle)))
        if not self._attach_dia(table, d_gcols, r0):
            rt.lib.cv_op_destroy(self.handle)
            self.handle = C.c_void_p()
            return False
        self._dia_only = True
        self._keep += [d_indptr, d_gcols, d_data]   # the handle borrows them
        return True

    # ---------------------------------------------------------------------------------------
    def _build_sell(self, force=False):
        rt, t = self.rt, self.rt.torch
        n_rows = self.n_local
        n_slices = (n_rows + 31) // 32
        widths = t.empty(n_slices, dtype=t.int32, device=rt.device)
        _lib.check(rt.lib.cv_op_sell_widths(rt.ctx, self.handle, widths.data_ptr(), rt.stream))
        w = widths.cpu().numpy().astype(np.int64)
        w = (w + 1) // 2 * 2          # SELL-32x2: columns are stored in pairs (128-bit value loads)
        slice_ptr = np.zeros(n_slices + 1, dtype=np.int64)
        np.cumsum(w * 32, out=slice_ptr[1:])
        padded = int(slice_ptr[-1])
        # SELL pays off for short, regular rows; keep CSR when padding would cost > 25 % traffic
        # or the matrix is tiny (dense test matrices)
        if not force and (padded > 1.25 * self.nnz or n_rows < 1024):
            return
        d_ptr = rt.upload(slice_ptr)
        d_col = t.empty(padded, dtype=t.int32, device=rt.device)
        d_val = t.empty(padded, dtype=t.float64, device=rt.device)
        _lib.check(rt.lib.cv_op_attach_sell(rt.ctx, self.handle, d_ptr.data_ptr(), padded,
                                            d_col.data_ptr(), d_val.data_ptr(), rt.stream))
        self._keep += [d_ptr, d_col, d_val]
        self.padded_nnz = padded
        self.format = "sell"

    def set_format(self, fmt):
        if getattr(self, "_dia_only", False) and fmt != "dia":
            raise NotImplementedError("this sharded operator was built directly in DIA storage; construct it "
                                      "with fmt='csr' or fmt='sell' to get the general halo plan")
        code = {"csr": _lib.CV_FMT_CSR, "sell": _lib.CV_FMT_SELL, "dia": _lib.CV_FMT_DIA}[fmt]
        if fmt == "sell" and self.padded_nnz == 0:
            self._build_sell(force=True)
        _lib.check(self.rt.lib.cv_op_set_format(self.handle, code))
        self.format = fmt

    # -- row-sharded mode --------------------------------------------------------------------
    def _localize(self, indptr, indices, data, r0, r1, full):
        """Renumber the columns of rows [r0, r1) to [owned | halo] (cv_halo_build, comm.cu).
        `full`: indptr/indices/data describe the whole matrix; otherwise only the row block
        (indptr rebased to 0, global column indices)."""
        rt = self.rt
        off = rt.offsets_for(self.shape[0])
        nloc = r1 - r0
        # the native routine indexes indptr[row0..row1]; for a pre-sliced block shift the base
        p_indptr = indptr.ctypes.data - (0 if full else 8 * r0)
        lo, hi = (int(indptr[r0]), int(indptr[r1])) if full else (0, int(indptr[-1]))
        n_halo = C.c_int64()
        _lib.check(rt.lib.cv_halo_count(p_indptr, indices.ctypes.data, r0, r1, C.byref(n_halo)))
        nh = n_halo.value
        halo_cols = np.empty(nh, dtype=np.int32)
        halo_owner = np.empty(nh, dtype=np.int32)
        loc_indptr = np.empty(nloc + 1, dtype=np.int64)
        loc_indices = np.empty(hi - lo, dtype=np.int32)
        _lib.check(rt.lib.cv_halo_build(p_indptr, indices.ctypes.data, r0, r1, off.ctypes.data,
                                        rt.world, nh, halo_cols.ctypes.data, halo_owner.ctypes.data,
                                        loc_indptr.ctypes.data, loc_indices.ctypes.data))
        self.n_local, self.n_halo = nloc, nh
        self.halo_cols, self.halo_owner = halo_cols, halo_owner
        self.row0 = r0
        return loc_indptr, loc_indices, np.ascontiguousarray(data[lo:hi])

    @classmethod
    def from_local_rows(cls, H_rows, n_global, runtime=None, fmt="auto"):
        """Row-sharded construction without ever holding the full matrix: `H_rows` is this rank's
        row block H[r_p:r_{p+1}] (scipy CSR, GLOBAL column indices), n_global the matrix order."""
        from .runtime import Runtime
        rt = runtime or Runtime.get()
        A = cls._as_csr(H_rows)
        if np.iscomplexobj(A):
            raise NotImplementedError("complex-valued operators are built from the whole matrix (from_host)")
        op = cls(rt, (n_global, n_global))
        r0, r1 = rt.local_range(n_global)
        if A.shape[0] != r1 - r0:
            raise ValueError(f"rank {rt.rank} owns rows [{r0},{r1}) but got {A.shape[0]} rows")
        indptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(A.indices, dtype=np.int32)
        data = np.ascontiguousarray(A.data, dtype=np.float64)
        op._diag_host = np.asarray(A.diagonal(k=r0), dtype=np.float64)   # rows are local, columns global
        if rt.world == 1:
            return op._finish(indptr, indices, data, fmt, indices, r0)
        return op._finish_sharded(indptr, indices, data, r0, r1, fmt)

    def _register_halo(self):
        """Exchange the halo request lists once (host, torch.distributed) and register the plan."""
        from .partition import exchange_plan
        rt, t = self.rt, self.rt.torch
        send_idx, send_off, recv_off = exchange_plan(self.halo_cols, self.halo_owner,
                                                     rt.offsets_for(self.shape[0]), rt.rank, rt.world)
        self.send_idx = send_idx
        d_send = rt.upload(send_idx) if len(send_idx) else t.empty(0, dtype=t.int32, device=rt.device)
        cap = max(len(send_idx), self.n_halo, 1)
        sendbuf = t.empty(2 * cap, dtype=t.float64, device=rt.device)
        halobuf = t.zeros(2 * cap, dtype=t.float64, device=rt.device)
        self._keep += [d_send, sendbuf, halobuf]
        self._send_off = np.ascontiguousarray(send_off, dtype=np.int64)
        self._recv_off = np.ascontiguousarray(recv_off, dtype=np.int64)
        _lib.check(rt.lib.cv_op_set_halo(rt.ctx, self.handle, self.n_halo, d_send.data_ptr(),
                                         self._send_off.ctypes.data, self._recv_off.ctypes.data,
                                         sendbuf.data_ptr(), halobuf.data_ptr()))
        if rt.transport == "peer":
            # halo buffer in exportable memory, two parities; peers gather-push into it
            stride = (16 * max(self.n_halo, 1) + 255) // 256 * 256
            ptrs, infos = rt.peer_shared_alloc(2 * stride, (stride, [int(x) for x in self._recv_off]))
            if ptrs is not None:
                self._peer_owned.append(ptrs[rt.rank])
                arr = (C.c_void_p * rt.world)(*ptrs)
                strides = np.ascontiguousarray([i[0] for i in infos], dtype=np.int64)
                dst_off = np.ascontiguousarray([i[1][rt.rank] for i in infos], dtype=np.int64)
                _lib.check(rt.lib.cv_op_set_halo_peers(rt.ctx, self.handle, C.cast(arr, C.POINTER(C.c_void_p)),
                                                       strides.ctypes.data, dst_off.ctypes.data))

    # -- Jacobi preconditioner ---------------------------------------------------------------
    def inverse_shifted_diagonal(self, sigma, reverse=False, cplx=False):
        """Device vector 1/(sigma - H_ii) (reverse: 1/(H_ii - sigma)) over this rank's rows, cached per
        shift: the diagonal right preconditioner of the shifted solves.  Entries whose denominator
        vanishes (|.| < 1e-300) are left unpreconditioned (1)."""
        key = (complex(sigma), bool(reverse), bool(cplx))
        hit = self._dinv.get(key)
        if hit is not None:
            return hit
        if self._diag_host is None:
            raise NotImplementedError("this operator does not know its diagonal")
        den = (complex(sigma) - self._diag_host) if not reverse else (self._diag_host - complex(sigma))
        if not cplx:
            den = den.real
        safe = np.abs(den) > 1e-300
        dinv = np.where(safe, 1.0 / np.where(safe, den, 1.0), 1.0)
        t = self.rt.upload(np.ascontiguousarray(dinv, dtype=np.complex128 if cplx else np.float64))
        if len(self._dinv) > 8:
            self._dinv.clear()
        self._dinv[key] = t
        return t

    # -- roofline bookkeeping ----------------------------------------------------
    def algorithmic_bytes(self, cplx=False):
        """Compulsory traffic of one shifted SpMV: 12*nnz + 20*N (fp64), 12*nnz + 36*N (complex)."""
        return 12 * self.nnz + (36 if cplx else 20) * self.n_local
