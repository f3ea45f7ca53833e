"""ctypes binding of libcudavec.so (include/cudavec.h).  Thin by design: argument marshalling
and error translation only.  There is no fallback: if the library is missing this module raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcudavec.so")

CV_OK = 0
CV_ABI_VERSION = 2
CV_SPMV_PLAIN, CV_SPMV_SHIFT, CV_SPMV_RSHIFT = 0, 1, 2
CV_FMT_CSR, CV_FMT_SELL, CV_FMT_DIA, CV_FMT_KRON = 0, 1, 2, 3
CV_SOLVER_GCROTMK, CV_SOLVER_MINRES = 0, 1
CV_ERR_NAMES = {1: "CUDA", 2: "ARG", 3: "UNSUPPORTED", 4: "NUMERIC", 5: "COMM"}


class CudaVecError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libcudavec error {code} ({CV_ERR_NAMES.get(code, '?')}): {msg}")
        self.code = code


class SolveStats(C.Structure):
    _fields_ = [("info", C.c_int), ("n_matvec", C.c_int), ("n_outer", C.c_int), ("n_sync", C.c_int),
                ("n_reorth", C.c_int), ("resid", C.c_double), ("b_norm", C.c_double),
                ("orth_loss", C.c_double), ("n_safe", C.c_int), ("n_recycled", C.c_int)]


_vp, _i, _i64, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t
_pd, _pi, _pi64, _pvp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol declared in include/cudavec.h is listed here and
# tests/test_abi.py checks the two lists agree.
SIGNATURES = {
    "cv_abi_version": (_i, []),
    "cv_last_error": (C.c_char_p, []),
    "cv_ctx_scratch_bytes": (_sz, []),
    "cv_ctx_create": (_i, [_i, _vp, _sz, _pvp]),
    "cv_ctx_destroy": (_i, [_vp]),
    "cv_ctx_launch_count": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "cv_ctx_sm_count": (_i, [_vp, _pi]),
    "cv_ctx_set_reorth_eta": (_i, [_vp, _d]),
    "cv_ctx_set_recycle": (_i, [_vp, _i]),
    "cv_ctx_set_option": (_i, [_vp, C.c_char_p, _d]),
    "cv_solve_precond": (_i, [_vp, _vp, _i, _i, _i, _d, _d, _vp, _vp, _vp, _d, _d, _i, _i, _i, _vp, _vp, _sz, _vp, _vp]),
    "cv_solve_batch": (_i, [_vp, _vp, _i, _i, _i, _pd, _pd, _pvp, _pvp, _pvp, _d, _d, _i, _i, _i, _vp, _sz, _vp, _vp]),
    "cv_arnoldi_step": (_i, [_vp, _vp, _i64, _i, _i, _pvp, _vp, _d, _d, _pd, _vp]),
    "cv_ctx_trace_read": (_i, [_vp, _pd, _i]),
    "cv_ctx_profile": (_i, [_vp, _i]),
    "cv_ctx_profile_read": (_i, [_vp, _i, _pd, C.POINTER(C.c_uint64), _pd]),
    "cv_comm_unique_id": (_i, [_vp]),
    "cv_comm_init": (_i, [_vp, _vp, _i, _i]),
    "cv_comm_init_peer_only": (_i, [_vp, _i, _i]),
    "cv_comm_finalize": (_i, [_vp]),
    "cv_comm_allreduce": (_i, [_vp, _vp, _i, _vp]),
    "cv_peer_window_bytes": (_sz, []),
    "cv_peer_alloc": (_i, [_vp, _sz, _pvp, _vp]),
    "cv_peer_open": (_i, [_vp, _vp, _pvp]),
    "cv_peer_close": (_i, [_vp, _vp]),
    "cv_peer_free": (_i, [_vp, _vp]),
    "cv_comm_attach_peers": (_i, [_vp, _pvp]),
    "cv_comm_transport": (_i, [_vp, _pi]),
    "cv_op_set_halo_peers": (_i, [_vp, _vp, _pvp, _vp, _vp]),
    "cv_op_dia_halo_bytes": (_sz, [_vp]),
    "cv_op_set_dia_halo_peers": (_i, [_vp, _vp, _pvp]),
    "cv_partition_rows": (_i, [_i64, _i, _pi64]),
    "cv_halo_count": (_i, [_vp, _vp, _i64, _i64, _pi64]),
    "cv_halo_build": (_i, [_vp, _vp, _i64, _i64, _vp, _i, _i64, _vp, _vp, _vp, _vp]),
    "cv_op_set_halo": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "cv_op_create_csr": (_i, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _pvp]),
    "cv_op_destroy": (_i, [_vp]),
    "cv_op_create_kron": (_i, [_vp, _i64, _i64, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i64, _i64, _pvp]),
    "cv_op_sell_widths": (_i, [_vp, _vp, _vp, _vp]),
    "cv_op_attach_sell": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "cv_op_attach_dia": (_i, [_vp, _vp, _i, _vp, _vp, _i64, _vp, _i64, _pi, _vp]),
    "cv_op_set_dia_halo": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "cv_dia_halo_plan": (_i, [_vp, _i, _i, _i64, _i64, _i, _pi, _vp, _pi, _vp]),
    "cv_op_set_imag": (_i, [_vp, _vp, _vp]),
    "cv_orth_slab_plan": (_i, [_i, _i, _i, _i, _pi, _vp, _vp]),
    "cv_orth_batch_plan": (_i, [_i, _vp, C.c_uint, _i, _i, _pi, _vp, _vp, _vp, _vp, _vp]),
    "cv_op_set_format": (_i, [_vp, _i]),
    "cv_op_info": (_i, [_vp, _pi64, _pi64, _pi64, _pi]),
    "cv_spmv": (_i, [_vp, _vp, _i, _i, _d, _d, _vp, _vp, _vp]),
    "cv_spmv_dots": (_i, [_vp, _vp, _i, _i, _d, _d, _vp, _vp, _pd, _vp]),
    "cv_copy": (_i, [_vp, _i64, _i, _vp, _vp, _vp]),
    "cv_scal": (_i, [_vp, _i64, _i, _i, _d, _d, _vp, _vp, _vp]),
    "cv_real": (_i, [_vp, _i64, _vp, _vp, _vp]),
    "cv_conj": (_i, [_vp, _i64, _vp, _vp, _vp]),
    "cv_dot": (_i, [_vp, _i64, _i, _i, _vp, _vp, _pd, _vp]),
    "cv_nrm2": (_i, [_vp, _i64, _i, _vp, _pd, _vp]),
    "cv_normalize": (_i, [_vp, _i64, _i, _vp, _pd, _vp]),
    "cv_lincomb": (_i, [_vp, _i64, _i, _i, _i, _pvp, _i, _pd, _pvp, _vp]),
    "cv_tsdot": (_i, [_vp, _i64, _i, _i, _i, _pvp, _i, _pvp, _pd, _vp]),
    "cv_gs_against_set": (_i, [_vp, _i64, _i, _vp, _i, _pvp, _d, _vp, _pi, _pd, _vp]),
    "cv_extend_columns": (_i, [_vp, _vp, _i64, _i, _i, _pvp, _vp, _pd, _pd, _vp]),
    "cv_solve_workspace_bytes": (_sz, [_i64, _i, _i, _i, _i]),
    "cv_solve": (_i, [_vp, _vp, _i, _i, _i, _d, _d, _vp, _vp, _vp, _d, _d, _i, _i, _i, _vp, _sz,
                      C.POINTER(SolveStats), _vp]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built — no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m eigensolvers_b200.build` "
            "(nvcc, sm_100a).  eigensolvers_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.cv_abi_version() != CV_ABI_VERSION:
        raise ImportError(f"libcudavec ABI {lib.cv_abi_version()} != {CV_ABI_VERSION}; rebuild the library")
    _lib = lib
    return lib


def check(rc):
    if rc != CV_OK:
        raise CudaVecError(rc, load().cv_last_error().decode("utf-8", "replace"))


def ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    return C.cast(arr, _pvp), arr  # keep `arr` alive while the call runs


def dbl_array(n):
    return (C.c_double * n)()
