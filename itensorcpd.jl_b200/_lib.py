"""ctypes binding of libitcpd_b200.so (include/itcpd_b200.h).  No CPU fallback: if the shared
library is missing or no sm_100 GPU is present every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libitcpd_b200.so")

OK, ERR_CUDA, ERR_ARG, ERR_NO_DEVICE, ERR_NAN, ERR_COMM, ERR_UNSUPPORTED = range(7)
SOLVE_CHOLESKY, SOLVE_QRCP = 0, 1
MTTKRP_TREE, MTTKRP_DIRECT = 0, 1


class ItcpdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libitcpd_b200 error {code}: {msg}")
        self.code = code


_lib = None

c_i64p = C.POINTER(C.c_int64)
c_dp = C.c_void_p  # double* passed as raw addresses (numpy .ctypes.data / torch .data_ptr())

_SIGS = {
    "itcpd_version": (C.c_int, []),
    "itcpd_last_error": (C.c_char_p, []),
    "itcpd_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "itcpd_destroy": (C.c_int, [C.c_void_p]),
    "itcpd_device_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "itcpd_synchronize": (C.c_int, [C.c_void_p]),
    "itcpd_launch_count": (C.c_int64, [C.c_void_p]),
    "itcpd_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "itcpd_set_tensor": (C.c_int, [C.c_void_p, C.c_int, c_i64p, c_dp]),
    "itcpd_set_shape": (C.c_int, [C.c_void_p, C.c_int, c_i64p]),
    "itcpd_generate_tensor": (C.c_int, [C.c_void_p, C.c_int, c_i64p, C.c_uint64, C.c_int64]),
    "itcpd_generate_lowrank_tensor": (C.c_int, [C.c_void_p, C.c_int, c_i64p, C.c_int, C.c_uint64, C.c_double]),
    "itcpd_get_tensor": (C.c_int, [C.c_void_p, c_dp]),
    "itcpd_tensor_norm": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "itcpd_set_rank": (C.c_int, [C.c_void_p, C.c_int]),
    "itcpd_set_factor": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "itcpd_get_factor": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "itcpd_set_lambda": (C.c_int, [C.c_void_p, c_dp]),
    "itcpd_get_lambda": (C.c_int, [C.c_void_p, c_dp]),
    "itcpd_get_gram": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "itcpd_random_cpd": (C.c_int, [C.c_void_p, C.c_uint64]),
    "itcpd_compute_grams": (C.c_int, [C.c_void_p]),
    "itcpd_gram_hadamard": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "itcpd_mttkrp": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "itcpd_solve": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "itcpd_last_solve_status": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "itcpd_normalize": (C.c_int, [C.c_void_p, C.c_int]),
    "itcpd_post_solve": (C.c_int, [C.c_void_p, C.c_int]),
    "itcpd_fit_terms": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "itcpd_cpd_snapshot": (C.c_int, [C.c_void_p]),
    "itcpd_cpd_diff_terms": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "itcpd_sweep": (C.c_int, [C.c_void_p, C.c_int, C.c_double, c_dp, c_dp]),
    "itcpd_sweep_async": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "itcpd_sweep_results": (C.c_int, [C.c_void_p, C.c_int, c_dp, c_dp, C.POINTER(C.c_int)]),
    "itcpd_als_from_host": (C.c_int, [C.c_void_p, C.c_int, c_i64p, c_dp, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_double,
                                      C.POINTER(C.c_void_p), c_dp, c_dp, c_dp]),
    "itcpd_reconstruct": (C.c_int, [C.c_void_p, c_dp]),
    "itcpd_residual_norm": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "itcpd_leverage_scores": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "itcpd_sample_factor_matrices": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_void_p]),
    "itcpd_pivot_hadamard": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, c_dp]),
    "itcpd_gather_fibers": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, c_dp]),
    "itcpd_column_to_multi_coords": (C.c_int, [C.c_int64, C.c_void_p, C.c_int, c_i64p, C.c_void_p]),
    "itcpd_multi_coords_to_column": (C.c_int, [C.c_int64, C.c_void_p, C.c_int, c_i64p, C.c_void_p]),
    "itcpd_sparse_sign": (None, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "itcpd_sparsestack": (None, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "itcpd_sparse_sign_fast_stream": (C.c_int, []),
    "itcpd_sketch_unfolding": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, c_dp]),
    "itcpd_sketch_unfolding_csc": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, c_dp]),
    "itcpd_sampled_update": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_double, C.c_int]),
    "itcpd_sampled_sweep_async": (C.c_int, [C.c_void_p, C.c_int, c_i64p, C.c_uint64, C.c_double, C.c_int]),
    "itcpd_qrcp_unfolding": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, c_dp]),
    "itcpd_qrcp_matrix": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_dp, C.c_int64, C.c_void_p, c_dp]),
    "itcpd_seqrcs": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, c_dp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "itcpd_seqrcs_modes": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "itcpd_seqrcs_krp": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, c_dp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "itcpd_set_projector": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "itcpd_projected_update": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int]),
    "itcpd_drop_tensor": (C.c_int, [C.c_void_p]),
    "itcpd_comm_unique_id": (C.c_int, [C.c_void_p]),
    "itcpd_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "itcpd_comm_destroy": (C.c_int, [C.c_void_p]),
    "itcpd_peer_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "itcpd_peer_import": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "itcpd_peer_disable": (C.c_int, [C.c_void_p]),
    "itcpd_allgather_factor": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_dp]),
    "itcpd_gemm_timing": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "itcpd_phase_timing": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int64)]),
    "itcpd_probe_dmma_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "itcpd_probe_dfma_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "itcpd_event_record": (C.c_int, [C.c_void_p, C.c_int]),
    "itcpd_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "itcpd_host_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "itcpd_host_free": (C.c_int, [C.c_void_p]),
    "itcpd_flush_l2": (C.c_int, [C.c_void_p, C.c_int64]),
}

DECLARED_SYMBOLS = tuple(_SIGS)


def load():
    """Load the shared library (built by csrc/build.sh / __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ItcpdError(ERR_NO_DEVICE, f"{LIB_PATH} not built: run itensorcpd.jl_b200/csrc/build.sh (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # raises AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status != OK:
        raise ItcpdError(status, load().itcpd_last_error().decode(errors="replace"))
