"""ctypes binding of include/recemb_b200.h (librecemb_b200.so).

This is the only place that touches the C ABI.  There is no CPU fallback: if
the library is missing, or a tensor is not on a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

import torch

LIB_PATH = Path(__file__).resolve().parent / "lib" / "librecemb_b200.so"

# enums (mirror include/recemb_b200.h)
F32, BF16 = 0, 1
HASH_IDENTITY, HASH_FLOORMOD, HASH_ROTL_FLOORMOD, HASH_QR_QUOTIENT, HASH_QR_REMAINDER, HASH_DIV_FLOORMOD = range(6)
EPI_NONE, EPI_L2NORM, EPI_RSQRT_K = range(3)
POOL_SUM, POOL_MEAN = 0, 1
UPD_DENSE_GRAD, UPD_SGD, UPD_ADAGRAD, UPD_ROWWISE_ADAGRAD, UPD_ADAM, UPD_ADAMW = range(6)

UPDATE_BY_NAME = {
    "dense_grad": UPD_DENSE_GRAD,
    "sgd": UPD_SGD,
    "adagrad": UPD_ADAGRAD,
    "rowwise_adagrad": UPD_ROWWISE_ADAGRAD,
    "adam": UPD_ADAM,
    "adamw": UPD_ADAMW,
}


class OptimParams(C.Structure):
    _fields_ = [
        ("lr", C.c_float),
        ("eps", C.c_float),
        ("weight_decay", C.c_float),
        ("beta1", C.c_float),
        ("beta2", C.c_float),
        ("bias_correction1", C.c_float),
        ("bias_correction2", C.c_float),
        ("grad_div", C.c_float),
    ]


class Layout(C.Structure):
    _fields_ = [
        ("ids_per_table", C.c_int64),
        ("num_tables", C.c_int32),
        ("shard_world", C.c_int32),
        ("shard_rank", C.c_int32),
        ("flip_len", C.c_int32),
        ("seq_len", C.c_int32),
        ("window_side", C.c_int32),
        ("window_keep", C.c_void_p),
        ("out_features", C.c_int32),
        ("out_feature_offset", C.c_int32),
        ("partition", C.c_int32),
    ]


class GatherTerm(C.Structure):
    """recemb_gather_term: one (table, ids, row transform) summand of recemb_multi_gather_add_fwd."""
    _fields_ = [
        ("table", C.c_void_p),
        ("num_rows", C.c_int64),
        ("ids", C.c_void_p),
        ("hash_mode", C.c_int),
        ("hash_arg", C.c_int64),
    ]


PEER_HANDLE_BYTES = 64
MAX_PEERS = 16


class PeerGroupStruct(C.Structure):
    """recemb_peer_group: every rank's arena / table shard as mapped in this process."""
    _fields_ = [
        ("world", C.c_int32),
        ("rank", C.c_int32),
        ("arena", C.c_void_p * MAX_PEERS),
        ("table", C.c_void_p * MAX_PEERS),
    ]


class PeerArena(C.Structure):
    """recemb_peer_arena: byte offsets inside one rank's exchange arena."""
    _fields_ = [(name, C.c_int64) for name in
                ("bytes", "off_flags", "off_epoch", "off_status", "off_counts", "off_inbox", "off_grads",
                 "cap", "bags_total", "off_parts", "off_gate")]


def make_layout(ids_per_table: int = 0, num_tables: int = 0, shard_world: int = 1, shard_rank: int = 0,
                flip_len: int = 0, window=None, out_features: int = 0, out_feature_offset: int = 0):
    """None when nothing is batched / sharded / flipped / windowed (the C side treats NULL as one plain table).
    window: a sequence.SequenceWindow (device-resident number of kept columns)."""
    if not ids_per_table and shard_world <= 1 and not flip_len and window is None:
        return None
    lay = Layout(ids_per_table=ids_per_table, num_tables=num_tables, shard_world=shard_world,
                 shard_rank=shard_rank, flip_len=flip_len, out_features=out_features,
                 out_feature_offset=out_feature_offset)
    if window is not None:
        if flip_len not in (0, window.seq_len):
            raise NativeError("a windowed lookup flips whole sequences: flip_len must be 0 or the window's seq_len")
        lay.seq_len, lay.window_side, lay.window_keep = window.seq_len, window.side, window.keep_ptr()
    return lay


class NativeLibraryMissing(RuntimeError):
    pass


class NativeError(RuntimeError):
    pass


_P, _I64, _I32, _INT, _SZ = C.c_void_p, C.c_int64, C.c_int32, C.c_int, C.c_size_t

# name -> (restype, argtypes); every symbol include/recemb_b200.h declares
SIGNATURES = {
    "recemb_abi_version": (_INT, []),
    "recemb_last_error": (C.c_char_p, []),
    "recemb_launch_count": (C.c_uint64, []),
    "recemb_row_index": (_INT, [_P, _I64, _INT, _I64, _I64, _P, _INT, _P]),
    "recemb_layout_total_rows": (_I64, [_I64, C.POINTER(Layout), _I64]),
    "recemb_gather_fwd": (_INT, [_P, _I64, _P, _I64, _I32, _INT, _P, _I64, C.POINTER(Layout), _INT, _INT, _I64,
                                 _INT, _INT, _I64, _P, _P, _INT, _P]),
    "recemb_kshift_fwd": (_INT, [_P, _I64, _I32, _INT, _P, _I64, _I32, _INT, _I32, _P, _P, _INT, _P]),
    "recemb_kshift_fwd_layout": (_INT, [_P, _I64, _I32, _INT, _P, _I64, _I32, _INT, C.POINTER(Layout), _P, _P, _INT,
                                        _P]),
    "recemb_multi_gather_add_fwd": (_INT, [_P, C.POINTER(GatherTerm), _I32, _I64, _I32, _INT, _P, _P, _P, _INT, _P]),
    "recemb_sequence_window_workspace_bytes": (_SZ, [_I32]),
    "recemb_sequence_window": (_INT, [_P, _INT, _I64, _I32, _I64, _I32, _INT, _P, _SZ, _P, _INT, _P]),
    "recemb_pool_fwd": (_INT, [_P, _I64, _I32, _INT, _P, _I64, _I32, _P, _I32, _P, _INT, _I64, _INT,
                               _INT, _I64, C.POINTER(Layout), _P, _INT, _P]),
    "recemb_bwd_plan_bytes": (_SZ, [_I64, _I64]),
    "recemb_bwd_plan": (_INT, [_P, _I64, C.POINTER(Layout), _I32, _INT, _I64, _I64, _INT, _I64, _I64, _I32,
                               _P, _I32, _P, _SZ, _INT, _P]),
    "recemb_plan_count": (_INT, [_P, _SZ, _I64, _I64, _INT, _P]),
    "recemb_plan_views": (_INT, [_P, _SZ, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P),
                                 C.POINTER(_I64)]),
    "recemb_bwd_apply_workspace_bytes": (_SZ, [_I64, _I32]),
    "recemb_bwd_apply": (_INT, [_P, _SZ, _I64, _P, _INT, _I64, _I32, _I32, _P, _P, _INT, _P, _INT, _I64,
                                _P, _P, C.POINTER(OptimParams), _P, _SZ, _INT, _P]),
    "recemb_bwd_apply_guarded": (_INT, [_P, _SZ, _I64, _P, _INT, _I64, _I32, _I32, _P, _P, _INT, _P, _INT, _I64,
                                        _P, _P, C.POINTER(OptimParams), _P, _SZ, _P, _INT, _P]),
    "recemb_time_next_apply": (_INT, [_P, _P]),
    "recemb_epilogue_bwd": (_INT, [_P, _P, _INT, _P, _I64, _I32, _INT, _I32, _P, _INT, _P]),
    "recemb_shard_bucket_workspace_bytes": (_SZ, [_I64, _I32]),
    "recemb_shard_bucket": (_INT, [_P, _I64, C.POINTER(Layout), _INT, _I64, _I64, _INT, _I64, _I32, _P, _I32,
                                   _I64, _P, _P, _P, _SZ, _INT, _P]),
    "recemb_pool_entries": (_INT, [_P, _I32, _INT, _P, _I64, _P, _INT, _P]),
    "recemb_bwd_plan_entries": (_INT, [_P, _I64, _I64, _P, _SZ, _INT, _P]),
    "recemb_sum_partials": (_INT, [_P, _I32, _I64, _I32, _INT, _P, _P, _INT, _P]),
    "recemb_peer_export": (_INT, [_P, _P, C.POINTER(_I64), C.POINTER(_I64), _INT]),
    "recemb_peer_open": (_INT, [_P, C.POINTER(_P), _INT]),
    "recemb_peer_close": (_INT, [_P, _INT]),
    "recemb_peer_arena_layout": (_INT, [_I32, _I64, _I64, _I32, _INT, C.POINTER(PeerArena)]),
    "recemb_peer_barrier": (_INT, [C.POINTER(PeerGroupStruct), C.POINTER(PeerArena), _INT, _INT, _P]),
    "recemb_peer_signal": (_INT, [C.POINTER(PeerGroupStruct), C.POINTER(PeerArena), _INT, _INT, _P]),
    "recemb_peer_wait": (_INT, [C.POINTER(PeerGroupStruct), C.POINTER(PeerArena), _INT, _INT, _P]),
    "recemb_peer_pool_fwd": (_INT, [C.POINTER(PeerGroupStruct), _I64, _I32, _INT, _P, _I64, _I32, _P, _I32, _P,
                                    _INT, _I64, _INT, _INT, _I64, C.POINTER(Layout), _P, _INT, _P]),
    "recemb_peer_bucket_push_rows": (_INT, [C.POINTER(PeerGroupStruct), C.POINTER(PeerArena), _P, _I64,
                                            C.POINTER(Layout), _INT, _I64, _I64, _INT, _I64, _P, _P, _SZ, _INT, _P]),
    "recemb_peer_rows_scatter_push": (_INT, [C.POINTER(PeerGroupStruct), C.POINTER(PeerArena), _P, _I64, _I32, _INT,
                                             _P, _INT, _P]),
    "recemb_peer_bwd_apply_fused": (_INT, [C.POINTER(PeerGroupStruct), C.POINTER(PeerArena), _P, _SZ, _P, _I32, _I64,
                                           _I32, _INT, _INT, _P, _I64, _I64, _P, C.POINTER(OptimParams), _P, _SZ,
                                           _I32, _INT, _P]),
    "recemb_peer_pool_push_tablewise": (_INT, [C.POINTER(PeerGroupStruct), C.POINTER(PeerArena), _I32, _INT, _I64, _INT,
                                               _P]),
    "recemb_peer_bwd_apply_fused_tablewise": (_INT, [C.POINTER(PeerGroupStruct), C.POINTER(PeerArena), _P, _SZ, _P, _I32,
                                                     _I64, _I32, _INT, _INT, _P, _I64, _I64, _P,
                                                     C.POINTER(OptimParams), _P, _SZ, _I32, _INT, _P]),
    "recemb_peer_pool_push": (_INT, [C.POINTER(PeerGroupStruct), C.POINTER(PeerArena), _I32, _INT, _INT, _P]),
    "recemb_peer_bucket_push": (_INT, [C.POINTER(PeerGroupStruct), C.POINTER(PeerArena), _P, _I64,
                                       C.POINTER(Layout), _INT, _I64, _I64, _INT, _I64, _I32, _P, _I32, _P, _SZ,
                                       _INT, _P]),
    "recemb_peer_allgather_push": (_INT, [C.POINTER(PeerGroupStruct), _P, _I64, _I64, _INT, _P]),
    "recemb_peer_plan": (_INT, [C.POINTER(PeerGroupStruct), C.POINTER(PeerArena), _I64, _P, _SZ, _INT, _P]),
    "recemb_dot_interaction_fwd": (_INT, [_P, _I64, _I32, _I32, _P, _INT, _P]),
    "recemb_dot_interaction_bwd": (_INT, [_P, _P, _I64, _I32, _I32, _P, _INT, _P]),
    "recemb_xxh64_ids": (_INT, [_P, _P, _I64, C.c_uint64, _INT, _P, _INT, _P]),
    "recemb_pad_histories": (_INT, [_P, _P, _P, _I64, _I32, _I64, _P, _INT, _P]),
    "recemb_logq_fwd": (_INT, [_P, _I32, _P, _I64, _P, _I64, _P, _INT, _P]),
    "recemb_logq_update": (_INT, [_P, _P, _I32, _P, _I64, _P, _I64, _P, C.c_double, _I64, _P, _INT, _P]),
    "recemb_flat_step_host": (_INT, [_P, _I64, _I64, _P, _P, _I64, _I32, _INT, _P, _P, _INT, _P, _P,
                                     C.POINTER(OptimParams), _P, _SZ, _P, _SZ, _P, _P, _P, _INT, _P]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """dlopen the in-tree library and type every entry point.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise NativeLibraryMissing(
                f"{LIB_PATH} is missing: build it with `python -m recommendations_b200.build_native` "
                "(there is no CPU / PyTorch fallback for the embedding hot path)")
        lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_LOCAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.recemb_abi_version() != 1:
            raise NativeError("librecemb_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().recemb_last_error().decode(errors="replace")
        raise NativeError(f"{what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(load().recemb_launch_count())


# ------------------------------------------------------------------ helpers ----
def dtype_code(t: torch.dtype) -> int:
    if t == torch.float32:
        return F32
    if t == torch.bfloat16:
        return BF16
    raise NativeError(f"unsupported table dtype {t}: the kernels take float32 or bfloat16")


def require_cuda(*tensors: torch.Tensor) -> int:
    """All tensors must live on one CUDA device and be contiguous; returns the device index."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise NativeError(
                "recommendations_b200 runs on CUDA (sm_100a) only: got a tensor on "
                f"{t.device}; there is no CPU fallback")
        if not t.is_contiguous():
            raise NativeError("non-contiguous tensor passed to the native layer")
        if dev is None:
            dev = t.device.index
        elif dev != t.device.index:
            raise NativeError("tensors on different CUDA devices")
    return 0 if dev is None else int(dev)


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr(device: int) -> int:
    return torch.cuda.current_stream(device).cuda_stream
