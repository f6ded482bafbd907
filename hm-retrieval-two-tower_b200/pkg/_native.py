"""
ctypes binding of libtt.so (include/tt.h).  There is NO fallback: if the library is not built, or no
CUDA device is present, every compute entry point raises.  torch is used only to own device memory
and streams; its tensors are passed to the C ABI as raw device pointers.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG_DIR)
LIB_PATH = os.environ.get("TT_LIB_PATH") or os.path.join(_ROOT, "lib", "libtt.so")   # (TT_LIB_PATH: development override, e.g. a build variant)

TT_IMPL_AUTO, TT_IMPL_SIMT, TT_IMPL_TC = 0, 1, 2
TT_MAX_FEATURES, TT_MAX_SRC, TT_MAX_JOBS, TT_MAX_KS = 16, 16, 32, 8
TT_PEER_SLOTS = 4

c_void_p, c_int, c_int32, c_int64, c_float, c_size_t = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t)


class TTFeature(ctypes.Structure):
    _fields_ = [("table", c_void_p), ("src", c_void_p), ("rows", c_int32), ("e", c_int32), ("col", c_int32),
                ("shards", c_int32)]


class TTStageCol(ctypes.Structure):
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("kind", c_int32), ("reserved", c_int32)]


TT_MAX_STAGE_COLS = 32


class TTSparseJob(ctypes.Structure):
    _fields_ = [("table", c_void_p), ("slot0", c_void_p), ("slot1", c_void_p), ("rows", c_int32), ("e", c_int32),
                ("nsrc", c_int32), ("n_per_src", c_int32), ("shard_rank", c_int32), ("shard_world", c_int32),
                ("ids", c_void_p * TT_MAX_SRC),
                ("grad", c_void_p * TT_MAX_SRC), ("grad_ld", c_int32 * TT_MAX_SRC)]


# name -> (restype, argtypes); every name here must be exported by libtt.so (tests/test_abi.py checks it
# against include/tt.h)
SIGNATURES = {
    "tt_version": (c_int, []),
    "tt_last_error": (ctypes.c_char_p, []),
    "tt_device_supports_tc": (c_int, []),
    "tt_launch_count": (c_int64, []),
    "tt_tc_available": (c_int, [c_int, c_int]),
    "tt_debug_tc": (c_int, [c_void_p, c_int]),
    "tt_debug_flash": (c_int, [c_void_p, c_int, c_int]),
    "tt_peer_alloc": (c_int, [c_size_t, ctypes.POINTER(c_void_p), c_void_p]),
    "tt_peer_open": (c_int, [c_void_p, ctypes.POINTER(c_void_p)]),
    "tt_peer_close": (c_int, [c_void_p]),
    "tt_peer_free": (c_int, [c_void_p]),
    "tt_peer_barrier": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p]),
    "tt_peer_sum_f32": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "tt_peer_gather_f32": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "tt_debug_index_cap": (c_int, [c_int]),
    "tt_debug_index_stages": (c_int, [c_void_p]),
    "tt_debug_index_layout": (c_int, [c_int, c_int64, c_int, c_int, c_int, ctypes.POINTER(c_int64)]),
    # host-side helpers (csrc/tt_host.cu)
    "tt_vocab_create": (c_void_p, [c_void_p, c_void_p, c_int64]),
    "tt_vocab_destroy": (None, [c_void_p]),
    "tt_vocab_size": (c_int64, [c_void_p]),
    "tt_vocab_lookup": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int]),
    "tt_vocab_lookup_fixed": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int]),
    "tt_crc32c": (ctypes.c_uint32, [c_void_p, c_size_t]),
    "tt_crc32c_portable": (ctypes.c_uint32, [c_void_p, c_size_t]),
    "tt_crc32c_masked": (ctypes.c_uint32, [c_void_p, c_size_t]),
    "tt_tfrecord_scan": (c_int64, [c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_int64]),
    "tt_tfrecord_frame": (c_int, [c_void_p, ctypes.c_uint64, c_void_p]),
    "tt_gather_cells": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int]),
    "tt_example_parse": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, ctypes.POINTER(ctypes.c_char_p), c_void_p, c_int, c_void_p,
                                 c_void_p, c_void_p, c_int]),
    "tt_gather_concat": (c_int, [ctypes.POINTER(TTFeature), c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "tt_dense_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                             c_void_p]),
    "tt_input_dense_fwd": (c_int, [ctypes.POINTER(TTFeature), c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                   c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "tt_dense_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "tt_dense_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p,
                             c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "tt_softmax_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "tt_inbatch_softmax_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "tt_inbatch_softmax_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                       c_void_p, c_int, c_void_p, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "tt_inbatch_softmax_step": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                        c_int, c_void_p, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "tt_inbatch_softmax_bwd_one": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                           c_void_p, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "tt_logits": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int,
                          c_void_p]),
    "tt_logq_apply": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "tt_log_f32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "tt_dense_adagrad": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_void_p]),
    "tt_dense_adam": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                              c_void_p]),
    "tt_fill_f32": (c_int, [c_void_p, c_float, c_int64, c_void_p]),
    "tt_sparse_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "tt_sparse_sort": (c_int, [ctypes.POINTER(TTSparseJob), c_int, c_void_p, c_size_t, c_void_p]),
    "tt_sparse_sort_passes": (c_int, [ctypes.POINTER(TTSparseJob), c_int, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "tt_sparse_adagrad": (c_int, [ctypes.POINTER(TTSparseJob), c_int, c_float, c_float, c_void_p, c_size_t, c_void_p]),
    "tt_sparse_adam": (c_int, [ctypes.POINTER(TTSparseJob), c_int, c_float, c_float, c_float, c_float, c_void_p, c_size_t,
                               c_void_p]),
    "tt_debug_sparse_plan": (c_int, [ctypes.POINTER(TTSparseJob), c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "tt_index_workspace_bytes": (c_size_t, [c_int, c_int64, c_int, c_int, c_int, c_int]),
    "tt_index_topk": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int64,
                              c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "tt_index_prepared_bytes": (c_size_t, [c_int64, c_int]),
    "tt_index_prepare": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "tt_round_tf32": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_void_p]),
    "tt_topk_merge": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "tt_take_i32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "tt_stamp": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p]),
    "tt_stage_columns": (c_int, [ctypes.POINTER(TTStageCol), c_int, c_int64, c_void_p]),
    "tt_fill_uniform": (c_int, [c_void_p, c_int64, c_int, c_int64, c_int64, ctypes.c_uint64, c_float, c_float, c_void_p]),
    "tt_recall_hits": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
}

_lib: Optional[ctypes.CDLL] = None


class TTError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libtt.so and bind every symbol of tt.h.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TTError(
                f"{LIB_PATH} is missing: build it with `python hm-retrieval-two-tower_b200/build.py` "
                "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here means the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().tt_last_error()
        raise TTError(f"{what or 'libtt call'} failed (rc={rc}): {msg.decode() if msg else ''}")


def require_cuda():
    """torch handle with a usable CUDA device, or a loud failure (the product has no CPU path)."""
    import torch

    if not torch.cuda.is_available():
        raise TTError("a CUDA device (B200, sm_100a) is required: this package has no CPU fallback")
    return torch


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()
