"""ctypes binding of libeqh_b200.so (the C ABI declared in include/eqh_b200.h).

There is no CPU fallback: if the library is missing or a call fails, a RuntimeError (or
NotImplementedError for EQH_ERR_UNSUPPORTED) is raised.  ctypes releases the GIL for the
duration of every call, and each call only enqueues kernels on the caller's current stream,
so the binding is safe from the autograd engine's device thread as well.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libeqh_b200.so")

EQH_ERR_UNSUPPORTED = -2

_P = c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "eqh_version": (c_int, []),
    "eqh_error_string": (c_char_p, [c_int]),
    "eqh_partial_rows": (c_int, []),
    "eqh_incidence_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "eqh_incidence_build_i64": (c_int, [_P, _P, c_int64, c_int64, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "eqh_incidence_build_i32": (c_int, [_P, _P, c_int64, c_int64, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "eqh_pool_index": (c_int, [_P, c_int64, _P, c_int64, _P, _P, _P]),
    "eqh_gather_reduce": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, _P, _P]),
    "eqh_gather_rows": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P]),
    "eqh_fused_msg_fwd": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, _P, _P, c_float, c_int, _P, _P]),
    "eqh_fused_msg_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int64, _P, _P, c_float, _P, _P, _P, _P, _P]),
    "eqh_rowln_fwd": (c_int, [_P, _P, _P, c_float, c_int, c_int64, c_int64, _P, _P]),
    "eqh_rowln_bwd": (c_int, [_P, _P, c_float, c_int, _P, c_int64, c_int64, _P, _P, _P]),
    "eqh_colsum_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "eqh_colsum_tall": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, c_int, _P, _P]),
    "eqh_colsum": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_float, c_int, _P]),
    "eqh_colsum_batch_workspace_bytes": (c_size_t, [_P, c_int]),
    "eqh_colsum_batch_counters": (c_int64, [_P, c_int]),
    "eqh_colsum_batch": (c_int, [_P, c_int, _P, _P, _P]),
    "eqh_knn16": (c_int, [_P, c_int64, _P, _P, _P]),
    "eqh_knn16_grid_workspace_bytes": (c_size_t, [c_int64]),
    "eqh_knn16_grid": (c_int, [_P, c_int64, _P, _P, _P, c_size_t, _P]),
    "eqh_knn16_grid_ordered": (c_int, [_P, c_int64, _P, _P, _P, _P, c_size_t, _P]),
    "eqh_knn": (c_int, [_P, c_int64, c_int, _P, _P, c_int, _P, _P, _P]),
    "eqh_radius": (c_int, [_P, c_int64, c_float, _P, _P, c_int, c_int, _P, _P, _P, _P]),
    "eqh_egnn_pack": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P]),
    "eqh_egnn_edge_fwd_smem_bytes": (c_size_t, [c_int64]),
    "eqh_egnn_edge_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, _P, _P, _P, _P]),
    "eqh_egnn_edge_fwd_tc": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, _P, _P, _P, _P]),
    "eqh_egnn_edge_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, _P]),
    "eqh_egnn_edge_fwd_tc_ordered": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, _P]),
    "eqh_egnn_edge_bwd_ordered": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, _P, _P]),
    "eqh_egnn_edge_bwd_tc_rows": (c_int, [c_int64]),
    "eqh_egnn_edge_bwd_tc_slices": (c_int, [c_int64, c_int64]),
    "eqh_egnn_edge_bwd_tc": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, _P]),
    "eqh_head_mse_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P]),
    "eqh_head_mse_bwd": (c_int, [_P, _P, _P, _P, c_int64, c_int64, _P, _P, _P, c_int, _P]),
    "eqh_atom_encoder_slices": (c_int, [c_int64]),
    "eqh_atom_encoder_fwd": (c_int, [_P, _P, _P, c_int, c_int64, c_int64, _P, _P, _P]),
    "eqh_atom_encoder_bwd": (c_int, [_P, _P, _P, c_int, c_int64, c_int64, _P, _P]),
    "eqh_collate_incidence": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P]),
    "eqh_collate": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, _P, _P]),
    "eqh_gemm_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "eqh_gemm_plan": (c_int, [c_int64, c_int64, c_int64, c_int, c_int, _P, _P, _P]),
    "eqh_gemm_tf32x3": (c_int, [_P, c_int, c_int64, _P, c_int, c_int64, _P, c_int64, c_int64, c_int64, c_int64, _P, _P, c_int,
                                _P, c_float, c_float, c_int, _P, _P, _P]),
    "eqh_adam_step": (c_int, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_float, c_int64,
                              c_float, _P]),
    "eqh_rowln_fwd_ld": (c_int, [_P, c_int64, _P, _P, c_float, c_int, c_int64, c_int64, _P, c_int64, _P]),
    "eqh_rowln_bwd_ld": (c_int, [_P, c_int64, _P, c_float, c_int, _P, c_int64, c_int64, c_int64, _P, c_int64, _P, _P, _P]),
    "eqh_gather_reduce_ld": (c_int, [_P, c_int64, _P, _P, _P, c_int64, c_int64, c_int, _P, c_int64, _P]),
    "eqh_egnn_edge_fwd_tc_ld": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, _P, c_int64, c_int64, _P, _P, _P, c_int64, _P, _P]),
    "eqh_egnn_edge_bwd_tc_ld": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64, _P, c_int64, _P,
                                        _P, _P, _P]),
    "eqh_egnn_unpack_grads": (c_int, [_P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P]),
    "eqh_gemm_tf32x3_ex": (c_int, [_P, c_int, c_int64, _P, c_int, c_int64, _P, c_int64, c_int64, c_int64, c_int64, _P, _P,
                                   c_int, _P, c_float, c_float, c_int, _P, _P, _P, _P, c_int64, _P]),
    "eqh_gemm_grouped_plan": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
    "eqh_gemm_grouped": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "eqh_egnn_coors_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P, _P, _P]),
    "eqh_egnn_coors_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "eqh_egnn_coors_bwd_pos": (c_int, [_P, _P, _P, _P, c_int64, _P, _P]),
    "eqh_batchnorm_partial_floats": (c_int, [c_int64]),
    "eqh_batchnorm_fwd": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, c_float, c_float, _P, _P, _P, _P, _P, _P, _P, _P,
                                  _P]),
    "eqh_batchnorm_bwd": (c_int, [_P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P]),
    "eqh_gemm_tf32x3_ex2": (c_int, [_P, c_int, c_int64, _P, c_int, c_int64, _P, c_int64, c_int64, c_int64, c_int64, _P, _P,
                                    c_int, _P, c_float, c_float, c_int, _P, _P, _P, _P, c_int64, _P, _P]),
    "eqh_gemm_tf32x3_ex3": (c_int, [_P, c_int, c_int64, _P, c_int, c_int64, _P, c_int64, c_int64, c_int64, c_int64, _P, _P,
                                    c_int, _P, c_float, c_float, c_int, _P, _P, _P, _P, c_int64, _P, _P, c_int64, _P]),
    "eqh_egnn_edge_slot_keys": (c_int, [c_int64]),
    "eqh_egnn_edge_slots_mark": (c_int, [_P, _P, c_int64, c_int64, _P, _P]),
    "eqh_egnn_edge_slots_assign": (c_int, [_P, _P, c_int64, c_int64, _P, _P, _P, _P]),
    "eqh_egnn_edge_bwd_tc_slots": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64, _P, c_int64, _P,
                                           _P, _P, _P, _P]),
    "eqh_cell_order": (c_int, [_P, c_int64, _P, _P, c_size_t, _P]),
    "eqh_rank1_update": (c_int, [_P, c_int64, _P, _P, c_int64, c_int64, c_float, _P]),
    "eqh_adam_step_dev": (c_int, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_float, _P, c_int, _P,
                                  c_float, _P]),
}
EXPORTS = tuple(_SIGNATURES)


class ColsumJob(ctypes.Structure):
    """struct eqh_colsum_job (include/eqh_b200.h)"""
    _fields_ = [("x", c_void_p), ("roww", c_void_p), ("out", c_void_p), ("rows", c_int64), ("cols", c_int64),
                ("ld", c_int64), ("scale", c_float), ("accumulate", ctypes.c_int32)]


COLSUM_MAX_JOBS = 32


class GemmProblem(ctypes.Structure):
    """struct eqh_gemm_problem (include/eqh_b200.h)"""
    _fields_ = [("A", c_void_p), ("lda", c_int64), ("B", c_void_p), ("ldb", c_int64), ("D", c_void_p), ("ldd", c_int64),
                ("M", c_int64), ("N", c_int64), ("K", c_int64), ("bias", c_void_p), ("rowscale", c_void_p),
                ("residual", c_void_p), ("out_scale", c_float), ("res_scale", c_float)]


GEMM_MAX_GROUP = 8


class Store(ctypes.Structure):
    """struct eqh_store (include/eqh_b200.h)"""
    _fields_ = [(n, c_void_p) for n in ("x", "pos", "y", "v_loc", "e_loc", "edge_attr", "e_order", "atom_off", "he_off",
                                        "inc_off")]


class StoreCsr(ctypes.Structure):
    """struct eqh_store_csr (include/eqh_b200.h)"""
    _fields_ = [(n, c_void_p) for n in ("rowptr_e_loc", "perm_e_loc", "rowptr_v_loc", "perm_v_loc")]


class IncidenceOut(ctypes.Structure):
    """struct eqh_incidence_out (include/eqh_b200.h)"""
    _fields_ = [(n, c_void_p) for n in ("rowptr_e", "perm_e", "vtx_by_e", "rowptr_v", "perm_v", "edge_by_v", "vertex32",
                                        "edges32")]


class Batch(ctypes.Structure):
    """struct eqh_batch (include/eqh_b200.h)"""
    _fields_ = [(n, c_void_p) for n in ("x", "pos", "edge_index0", "edge_index1", "edge_attr", "n_e", "e_order", "batch",
                                        "ptr", "y")]

_lib = None


def load() -> ctypes.CDLL:
    """Load the CUDA library (once).  Fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m equihgnn_b200.build` "
            "(equihgnn_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = load().eqh_error_string(rc).decode()
    if rc == EQH_ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    raise RuntimeError(f"{what} failed with code {rc}: {msg}")
