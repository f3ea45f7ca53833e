"""Process-wide device runtime: the cv_ctx handle, its scratch memory (a torch tensor — torch is
used only as the device-memory holder and for streams / torch.distributed plumbing), the
operator cache and the solver workspaces.

One process drives one GPU.  In row-sharded mode (``init_distributed``) every rank runs the same
host code on its row block; all scalar results are summed over ranks inside libcudavec, so every
rank takes identical control-flow decisions.
"""
import ctypes as C
import os
import weakref

import numpy as np

from . import _lib


class Runtime:
    _instance = None

    def __init__(self):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("eigensolvers_b200 needs a CUDA device (B200, sm_100a); "
                               "there is no CPU fallback")
        self.torch = torch
        self.lib = _lib.load()
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.device_index = local_rank if local_rank < torch.cuda.device_count() else 0
        torch.cuda.set_device(self.device_index)
        self.device = torch.device("cuda", self.device_index)
        nbytes = self.lib.cv_ctx_scratch_bytes()
        self.scratch = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        handle = C.c_void_p()
        _lib.check(self.lib.cv_ctx_create(self.device_index, self.scratch.data_ptr(), nbytes,
                                          C.byref(handle)))
        self.ctx = handle
        self.rank, self.world = 0, 1
        self.offsets = None  # row partition in distributed mode (world+1 int64)
        self._workspaces = {}
        self._op_cache = {}
        self._tmp = {}
        self._peer_allocs = []
        self._peer_graveyard = []
        self.stats = {"solves": 0, "matvecs": 0, "syncs": 0, "outer": 0}

    # -- singletons ---------------------------------------------------------------------------
    @classmethod
    def get(cls):
        if cls._instance is None:
            cls._instance = Runtime()
        return cls._instance

    @classmethod
    def reset(cls):
        inst = cls._instance
        if inst is not None:
            inst._op_cache.clear()
            inst._workspaces.clear()
            for own, _ in list(inst._peer_allocs):
                inst.peer_release(own)
            try:
                import torch.distributed as dist
                if inst.world > 1 and dist.is_initialized():
                    inst.torch.cuda.synchronize(inst.device)
                    dist.barrier()
            except Exception:
                pass
            inst._drain_graveyard()
            inst.lib.cv_ctx_destroy(inst.ctx)
        cls._instance = None

    # -- helpers --------------------------------------------------------------------------------
    @property
    def stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def launch_count(self):
        n = C.c_uint64()
        _lib.check(self.lib.cv_ctx_launch_count(self.ctx, C.byref(n)))
        return n.value

    def empty(self, n, cplx):
        t = self.torch
        return t.empty(int(n), dtype=t.complex128 if cplx else t.float64, device=self.device)

    def upload(self, host_array, dtype=None):
        """Host ndarray -> device tensor through pinned staging (counted by bench.py's e2e)."""
        t = self.torch
        a = np.ascontiguousarray(host_array, dtype=dtype)
        src = t.from_numpy(a)
        if a.nbytes >= (1 << 20):
            src = src.pin_memory()
        return src.to(self.device, non_blocking=False)

    def workspace(self, nbytes):
        """Solver workspace, grown on demand and reused across solves."""
        t = self.torch
        cur = self._workspaces.get("solve")
        if cur is None or cur.numel() < nbytes:
            self._workspaces.pop("solve", None)
            cur = t.empty(int(nbytes), dtype=t.uint8, device=self.device)
            self._workspaces["solve"] = cur
        return cur

    def tmp_vector(self, key, n, cplx):
        cur = self._tmp.get((key, cplx))
        if cur is None or cur.numel() != n:
            cur = self.empty(n, cplx)
            self._tmp[(key, cplx)] = cur
        return cur

    # -- row-sharded mode ----------------------------------------------------------------------
    def init_distributed(self):
        """Enter row-sharded mode over an already initialised torch.distributed group (any backend:
        the group is only used for rendezvous and object exchange).  Ranks on distinct GPUs get an
        NCCL communicator (fall-back transport) plus the peer-memory windows; ranks that SHARE a GPU
        (the 2-process parity tests on one device: NCCL refuses duplicate GPUs, CUDA IPC does not)
        run on the peer-memory transport alone."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return self
        if self.world > 1:
            return self
        rank, world = dist.get_rank(), dist.get_world_size()
        ids = [None] * world
        props = self.torch.cuda.get_device_properties(self.device)
        gpu_id = str(getattr(props, "uuid", None) or getattr(props, "pci_bus_id", self.device_index))
        dist.all_gather_object(ids, (os.uname().nodename, gpu_id))
        shared_gpu = len(set(ids)) < world
        if shared_gpu:
            _lib.check(self.lib.cv_comm_init_peer_only(self.ctx, rank, world))
        else:
            uid = (C.c_char * 128)()
            if rank == 0:
                _lib.check(self.lib.cv_comm_unique_id(uid))
            box = [bytes(uid)]
            dist.broadcast_object_list(box, src=0)
            uid = (C.c_char * 128).from_buffer_copy(box[0])
            _lib.check(self.lib.cv_comm_init(self.ctx, uid, rank, world))
        self.rank, self.world = rank, world
        self._attach_peer_windows(required=shared_gpu)
        return self

    # -- peer-memory transport (csrc/peer.cu) ---------------------------------------------------
    @property
    def transport(self):
        """'single' | 'nccl' | 'peer' — what the scalar all-reduce and the halo exchange use."""
        code = C.c_int()
        _lib.check(self.lib.cv_comm_transport(self.ctx, C.byref(code)))
        return ("single", "nccl", "peer")[code.value]

    def peer_shared_alloc(self, nbytes, info=None):
        """Collective: every rank allocates `nbytes` (its own value) of CUDA-IPC exportable device
        memory and maps the other ranks' allocations.  Returns (ptrs[world], infos[world]) — ptrs[p]
        is rank p's allocation as addressable from THIS process — or (None, infos) if some rank
        could not export/map (the caller then stays on NCCL)."""
        import torch.distributed as dist
        own, handle = C.c_void_p(), (C.c_char * 64)()
        rc = self.lib.cv_peer_alloc(self.ctx, int(nbytes), C.byref(own), handle)
        box = [None] * self.world
        dist.all_gather_object(box, (bytes(handle) if rc == 0 else None, info))
        self._drain_graveyard()   # every rank is here: nobody touches the retired buffers any more
        infos = [b[1] for b in box]
        ptrs, opened, ok = [None] * self.world, [], all(b[0] is not None for b in box)
        if ok:
            for p in range(self.world):
                if p == self.rank:
                    ptrs[p] = own.value
                    continue
                q = C.c_void_p()
                if self.lib.cv_peer_open(self.ctx, (C.c_char * 64).from_buffer_copy(box[p][0]), C.byref(q)) != 0:
                    ok = False
                    break
                ptrs[p] = q.value
                opened.append(q.value)
        flags = [None] * self.world
        dist.all_gather_object(flags, ok)
        if not all(flags):
            for q in opened:
                self.lib.cv_peer_close(self.ctx, C.c_void_p(q))
            if rc == 0:
                self.lib.cv_peer_free(self.ctx, own)
            return None, infos
        self._peer_allocs.append((own.value, opened))
        return ptrs, infos

    def peer_release(self, own_ptr):
        """Retire one peer_shared_alloc (identified by this rank's own pointer).  A peer's last
        Arnoldi step may still be storing halo rows into this buffer (the push of a vector nobody
        will multiply any more), so the memory is only parked here; it is unmapped and freed at the
        next collective allocation, when every rank's host has provably passed its final
        synchronisation (`_drain_graveyard`)."""
        for i, (own, opened) in enumerate(self._peer_allocs):
            if own == own_ptr:
                self._peer_graveyard.append(self._peer_allocs.pop(i))
                return

    def _drain_graveyard(self):
        if not self._peer_graveyard:
            return
        self.torch.cuda.synchronize(self.device)
        for own, opened in self._peer_graveyard:
            for q in opened:
                self.lib.cv_peer_close(self.ctx, C.c_void_p(q))
            self.lib.cv_peer_free(self.ctx, C.c_void_p(own))
        self._peer_graveyard = []

    def _attach_peer_windows(self, required=False):
        if not required and (os.environ.get("EIGB200_TRANSPORT", "peer").lower() != "peer" or self.world > 8):
            return
        ptrs, _ = self.peer_shared_alloc(self.lib.cv_peer_window_bytes())
        if ptrs is None:
            if required:
                raise RuntimeError("ranks share a GPU, so NCCL is unavailable, and CUDA IPC peer mapping failed: "
                                   + self.lib.cv_last_error().decode())
            import warnings
            warnings.warn("CUDA IPC peer mapping unavailable: collectives stay on NCCL "
                          f"({self.lib.cv_last_error().decode()})")
            return
        arr = (C.c_void_p * self.world)(*ptrs)
        _lib.check(self.lib.cv_comm_attach_peers(self.ctx, C.cast(arr, C.POINTER(C.c_void_p))))

    def offsets_for(self, n):
        from .partition import row_offsets
        return row_offsets(n, self.world)

    def local_range(self, n):
        off = self.offsets_for(n)
        return int(off[self.rank]), int(off[self.rank + 1])

    # -- operator cache -------------------------------------------------------------------------
    @staticmethod
    def _fingerprint(H):
        """Cheap content fingerprint of a host matrix: storage addresses, sizes and a strided sample
        of the values.  The reference evaluates `H @ x` on every call (numpyVector.py:100,152), so an
        in-place edit of H (H.data *= ..., H -= shift*I, ndarray writes) must not be served from a
        stale device copy; the sample catches whole-array edits, `invalidate_operator` is the
        explicit route for anything finer."""
        try:
            import scipy.sparse as sp
            if sp.issparse(H):
                d = H.data
                parts = [H.shape, int(H.nnz), d.ctypes.data,
                         getattr(getattr(H, "indices", None), "ctypes", None) and H.indices.ctypes.data]
            else:
                d = np.asarray(H).reshape(-1)
                parts = [np.shape(H), d.ctypes.data]
            if d.size:
                step = max(1, d.size // 4096)
                parts += [d[::step].sum().item(), d[0].item(), d[-1].item()]   # python float / complex
            return tuple(parts)
        except Exception:
            return None

    def invalidate_operator(self, H=None):
        """Drop the cached device copy of `H` (all cached operators when H is None) — call after
        editing a matrix in place."""
        if H is None:
            self._op_cache.clear()
        else:
            self._op_cache.pop(id(H), None)

    def operator_for(self, H):
        """Device operator for a host matrix, cached by object identity + content fingerprint."""
        from .operator import DeviceOperator
        if isinstance(H, DeviceOperator):
            return H
        if getattr(H, "_is_device_operator", False):   # e.g. KroneckerSumOperator
            return H
        key = id(H)
        fp = self._fingerprint(H)
        hit = self._op_cache.get(key)
        if hit is not None and hit[0]() is H and hit[2] == fp:
            return hit[1]
        op = DeviceOperator.from_host(H, runtime=self)
        try:
            ref = weakref.ref(H, lambda _r, k=key, cache=self._op_cache: cache.pop(k, None))
        except TypeError:  # object without weakref support: keep it alive with the cache entry
            ref = (lambda obj: (lambda: obj))(H)
        self._op_cache[key] = (ref, op, fp)
        return op
