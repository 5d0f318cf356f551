"""ctypes binding of ``include/cgnn.h`` (the C ABI of the sm_100a kernels).

``load()`` opens the in-tree ``lib/libcgnn.so`` built by ``csrc/Makefile`` and fails loudly if it
is missing: this package has no CPU or eager-PyTorch fallback for the hot path.
``bind(path)`` declares the prototypes on an arbitrary shared object with the same ABI; the
package itself only ever calls it with the in-tree CUDA library.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcgnn.so")
ABI_VERSION = 11

c_f32p = C.c_void_p   # device pointers travel as integers (tensor.data_ptr())
c_ptr = C.c_void_p


class CsrT(C.Structure):
    """``cgnn_csr_t`` / ``cgnn_csr_out_t`` (identical layout)."""
    _fields_ = [(n, C.c_void_p) for n in (
        "in_rowptr", "in_col", "in_w", "in_wn",
        "out_rowptr", "out_col", "out_w", "out_wn",
        "deg", "dinv", "wsum", "graph_meta",
        "agg_in", "agg_out", "row_graph")] + [("agg_kind", C.c_int32)]


class ActT(C.Structure):
    """``cgnn_act_t``"""
    _fields_ = [
        ("scale", C.c_void_p), ("shift", C.c_void_p), ("relu", C.c_int32), ("p_drop", C.c_float),
        ("seed", C.c_uint64), ("site", C.c_uint32), ("row_base", C.c_int64), ("salt", C.c_void_p),
    ]


class BnBwdT(C.Structure):
    """``cgnn_bn_bwd_t``"""
    _fields_ = [
        ("scale", C.c_void_p), ("mean", C.c_void_p), ("rstd", C.c_void_p), ("s1", C.c_void_p),
        ("s2", C.c_void_p), ("count", C.c_double), ("train", C.c_int32), ("sums64", C.c_void_p),
    ]


class EvalLayerT(C.Structure):
    """``cgnn_eval_layer_t``"""
    _fields_ = [("W", C.c_void_p), ("bias", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("eps", C.c_float)]


class StoreT(C.Structure):
    """``cgnn_store_t``"""
    _fields_ = [
        ("x", C.c_void_p), ("src", C.c_void_p), ("dst", C.c_void_p), ("w", C.c_void_p),
        ("node_ptr", C.c_void_p), ("edge_ptr", C.c_void_p), ("label", C.c_void_p),
        ("num_features", C.c_int32), ("edge_pairs", C.c_int32),
    ]


_i32, _i64, _f32, _u64, _sz, _p = C.c_int32, C.c_int64, C.c_float, C.c_uint64, C.c_size_t, C.c_void_p
_P = C.POINTER

# name -> (restype, argtypes); one entry per function declared in include/cgnn.h
PROTOTYPES = {
    "cgnn_status_string": (C.c_char_p, [C.c_int]),
    "cgnn_abi_version": (C.c_int, []),
    "cgnn_last_cuda_error": (C.c_int, []),
    "cgnn_workspace_bytes": (_sz, []),
    "cgnn_kernel_launches": (C.c_uint64, []),
    "cgnn_set_option": (C.c_int, [_i32, _i32]),
    "cgnn_collate_csr": (C.c_int, [_P(StoreT), _p, _i64, _i64, _i64, _i32, _i32, _p, _p, _p, _p, _p, _p, _p,
                                   _P(CsrT), _p]),
    "cgnn_fetch_ids": (C.c_int, [_p, _i64, _p, _p]),
    "cgnn_ingest_threshold": (C.c_int, [_p, _i64, _i32, C.c_float, _p, _p, _p]),
    "cgnn_ingest_emit": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "cgnn_csr_from_coo": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _i32, _p, _P(CsrT), _p]),
    "cgnn_agg_words": (_sz, [_i64, _i64, _i64]),
    "cgnn_build_agg": (C.c_int, [_P(CsrT), _i32, _i64, _i64, _i64, _i32, _p, _p, _p, _p]),
    "cgnn_gcn_layer_fwd": (C.c_int, [_p, _P(ActT), _p, _p, _P(CsrT), _p, _i64, _i64, _i32, _i32, _i32, _i32, _p, _p,
                                     _p, _sz, _p]),
    "cgnn_gcn_layer_fwd_pool": (C.c_int, [_p, _P(ActT), _p, _p, _P(CsrT), _p, _i64, _i64, _i32, _i32, _i32, _i32, _P(ActT), _p, _p]),
    "cgnn_peer_exchange": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, C.c_uint32, _i32, _p, _i32, _i32, _p, _p, _p]),
    "cgnn_sage_layer_fwd": (C.c_int, [_p, _P(ActT), _p, _p, _P(CsrT), _p, _i64, _i64, _i32, _i32, _i32, _i32, _p, _p,
                                      _p, _p, _sz, _p]),
    "cgnn_eval_fused_fwd": (C.c_int, [_i32, _p, _i32, _P(EvalLayerT), _i32, _i32, _p, _p, _p, _p, _i32, _i32, _P(CsrT), _p,
                                      _i64, _i64, _i32, _i32, _p, _p, _p, _sz, _p]),
    "cgnn_project_tf32x3": (C.c_int, [_p, _p, _i64, _i32, _i32, _p, _p]),
    "cgnn_bn_merge_stats": (C.c_int, [_p, _i32, _i32, _p, _p]),
    "cgnn_bn_finalize": (C.c_int, [_p, _p, _p, _i32, _f32, _f32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "cgnn_bn_eval_affine": (C.c_int, [_p, _p, _p, _p, _i32, _f32, _p, _p, _p, _p, _p]),
    "cgnn_pool_fwd": (C.c_int, [_p, _P(ActT), _p, _i64, _i64, _i32, _p, _p]),
    "cgnn_head_fwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _f32, _u64, _i64, _p, _p, _p, _p]),
    "cgnn_step_tick": (C.c_int, [_p, _p]),
    "cgnn_adam_step": (C.c_int, [_p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _f32, _p, _p]),
    "cgnn_ce_fwd": (C.c_int, [_p, _p, _i64, _i32, _f32, _p, _p, _p, _p]),
    "cgnn_ce_bwd": (C.c_int, [_p, _p, _i64, _i32, _f32, _p, _p, _p]),
    "cgnn_head_bwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _f32, _p, _p, _p, _p, _p, _p, _sz,
                                _p]),
    "cgnn_bn_bwd_sums": (C.c_int, [_p, _P(ActT), _p, _p, _p, _p, _p, _i64, _i64, _i32, _p, _p, _p, _sz, _p]),
    "cgnn_gcn_layer_bwd": (C.c_int, [_p, _p, _p, _P(ActT), _P(BnBwdT), _p, _P(ActT), _p, _P(CsrT), _p, _i64, _i64,
                                     _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "cgnn_sage_layer_bwd": (C.c_int, [_p, _p, _p, _P(ActT), _P(BnBwdT), _p, _p, _P(ActT), _p, _P(CsrT), _p, _i64,
                                      _i64, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
}


class CgnnError(RuntimeError):
    """A C-ABI call returned a non-zero ``cgnn_status`` (``.status``)."""

    def __init__(self, message: str, status: int = -1):
        super().__init__(message)
        self.status = status


ERR_NEED_CSR = 5   # a lean batch met a code path that reads the CSR arrays
ERR_UNSUPPORTED = 6   # cgnn_eval_fused_fwd does not cover the shape


def bind(path: str) -> C.CDLL:
    """dlopen ``path`` and declare every prototype of include/cgnn.h on it."""
    lib = C.CDLL(path)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = the library does not export the ABI
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.cgnn_abi_version()
    if got != ABI_VERSION:
        raise CgnnError(f"{path}: ABI version {got}, expected {ABI_VERSION}")
    return lib


_LIB: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """The product library.  Raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise CgnnError(
                f"{LIB_PATH} is missing: build it with `make -C connectome-gnn-suite_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "connectome_gnn (B200 build) has no CPU / eager-PyTorch fallback.")
        _LIB = bind(LIB_PATH)
    return _LIB


def check(lib: C.CDLL, status: int, what: str) -> None:
    if status != 0:
        msg = lib.cgnn_status_string(status).decode()
        extra = f" (cudaError {lib.cgnn_last_cuda_error()})" if status == 4 else ""
        raise CgnnError(f"{what}: {msg}{extra}", status)
