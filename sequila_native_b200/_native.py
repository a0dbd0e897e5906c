"""ctypes binding of libsequila_cuda.so — the same C ABI (include/sequila_cuda.h) a Rust
`sequila-cuda-sys` crate would bind.  The library must exist: there is no
CPU fallback and no oracle import anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SQ_LIB_PATH") or os.path.join(_HERE, "libsequila_cuda.so")  # SQ_LIB_PATH: A/B experiments only

SQ_OK, SQ_EINVAL, SQ_ECUDA, SQ_ENOMEM, SQ_ESTATE, SQ_ECAPACITY, SQ_ECAST, SQ_EPARSE, SQ_EBUSY = range(9)
TILE_COUNT_ONLY, TILE_RIGHT_IDX, TILE_EXPAND_RIGHT, TILE_NO_COUNTS, TILE_COUNTS_U8 = 1, 2, 4, 8, 16  # SQ_TILE_* flags
NULL_INDEX = 0xFFFFFFFF  # SQ_NULL_INDEX

u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)
vp = C.c_void_p

class SqTileOut(C.Structure):
    """struct sq_tile_out (include/sequila_cuda.h)"""
    _fields_ = [("n_pairs", C.c_uint64), ("n_rows", C.c_uint32), ("counts_width", C.c_uint32),
                ("left_idx", C.c_void_p), ("right_idx", C.c_void_p), ("counts", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/sequila_cuda.h declares
SIGNATURES = {
    "sq_abi_version": (C.c_int32, []),
    "sq_ctx_create": (C.c_int32, [C.c_int32, C.POINTER(vp)]),
    "sq_ctx_destroy": (None, [vp]),
    "sq_last_error": (C.c_char_p, [vp]),
    "sq_device_count": (C.c_int32, []),
    "sq_ctx_set_option": (C.c_int32, [vp, C.c_char_p, C.c_char_p]),
    "sq_ctx_get_option": (C.c_int32, [vp, C.c_char_p, C.c_char_p, C.c_size_t]),
    "sq_host_alloc": (C.c_int32, [vp, C.c_size_t, C.POINTER(vp)]),
    "sq_host_free": (None, [vp, vp]),
    "sq_index_build": (C.c_int32, [vp, vp, vp, vp, C.c_uint64, C.POINTER(vp)]),
    "sq_index_build_device": (C.c_int32, [vp, vp, vp, vp, C.c_uint64, vp, C.POINTER(vp)]),
    "sq_index_bytes": (C.c_uint64, [vp]),
    "sq_index_rows": (C.c_uint64, [vp]),
    "sq_index_keys": (C.c_uint64, [vp]),
    "sq_index_uses_packed": (C.c_int32, [vp]),
    "sq_index_uses_rank": (C.c_int32, [vp]),
    "sq_index_sort_key_bits": (C.c_int32, [vp]),
    "sq_index_uses_positions": (C.c_int32, [vp]),
    "sq_index_position_rows": (C.c_int32, [vp, vp]),
    "sq_index_position_rows_device": (vp, [vp]),
    "sq_index_build_ms": (C.c_float, [vp]),
    "sq_index_free": (None, [vp]),
    "sq_index_add_column": (C.c_int32, [vp, vp, C.c_uint32, C.POINTER(C.c_int32)]),
    "sq_index_add_column_device": (C.c_int32, [vp, vp, C.c_uint32, C.POINTER(C.c_int32)]),
    "sq_stream_create": (C.c_int32, [vp, C.POINTER(vp)]),
    "sq_stream_create_on": (C.c_int32, [vp, vp, C.POINTER(vp)]),
    "sq_stream_free": (None, [vp]),
    "sq_stream_last_error": (C.c_char_p, [vp]),
    "sq_stream_bytes": (C.c_uint64, [vp]),
    "sq_probe_count": (C.c_int32, [vp, vp, vp, vp, vp, C.c_uint32, u64p]),
    "sq_probe_emit_pairs": (C.c_int32, [vp, vp, vp, vp, C.c_uint64]),
    "sq_probe_join": (C.c_int32, [vp, vp, vp, vp, vp, C.c_uint32, vp, vp, vp, C.c_uint64, u64p]),
    "sq_probe_nearest": (C.c_int32, [vp, vp, vp, vp, vp, C.c_uint32, vp]),
    "sq_probe_nearest_device": (C.c_int32, [vp, vp, vp, vp, vp, C.c_uint32, vp]),
    "sq_probe_join_device": (C.c_int32, [vp, vp, vp, vp, vp, C.c_uint32, vp, vp, C.c_uint64, u64p]),
    "sq_probe_count_device": (C.c_int32, [vp, vp, vp, vp, vp, C.c_uint32, u64p]),
    "sq_probe_emit_pairs_device": (C.c_int32, [vp, vp, vp, C.c_uint64]),
    "sq_stream_counts_device": (vp, [vp]),
    "sq_gather_column": (C.c_int32, [vp, C.c_int32, C.c_int32, vp, C.c_uint32, vp, C.c_uint64]),
    "sq_gather_column_device": (C.c_int32, [vp, C.c_int32, C.c_int32, vp, C.c_uint32, vp, C.c_uint64]),
    "sq_index_pack_columns": (C.c_int32, [vp, i32p, C.c_int32, C.POINTER(C.c_int32)]),
    "sq_gather_pack_device": (C.c_int32, [vp, C.c_int32, C.POINTER(vp), C.c_int32, C.c_uint64]),
    "sq_gather_probe_columns_device": (C.c_int32, [vp, C.POINTER(vp), C.POINTER(vp), C.c_int32, C.c_uint64]),
    "sq_index_add_utf8_column": (C.c_int32, [vp, vp, vp, C.c_uint64, C.POINTER(C.c_int32)]),
    "sq_gather_utf8": (C.c_int32, [vp, C.c_int32, C.c_int32, vp, vp, C.c_uint64, vp, u64p]),
    "sq_gather_utf8_data": (C.c_int32, [vp, vp, C.c_uint64]),
    "sq_index_set_validity": (C.c_int32, [vp, C.c_int32, vp]),
    "sq_gather_validity": (C.c_int32, [vp, C.c_int32, C.c_int32, vp, vp, u64p]),
    "sq_stream_counts": (C.c_int32, [vp, vp]),
    "sq_stream_set_window": (C.c_int32, [vp, C.c_uint64, C.c_uint64]),
    "sq_fetch_pairs": (C.c_int32, [vp, vp, vp, C.c_uint64]),
    "sq_cast_i64_to_i32": (C.c_int32, [vp, vp, C.c_uint64, C.c_int64, vp]),
    "sq_rle_expand_variant": (C.c_int32, [C.c_int32, vp, C.c_uint32, vp, C.c_uint64]),
    "sq_pairs_digest_device": (C.c_int32, [vp, vp, vp, C.c_uint64, C.c_uint64, u64p]),
    "sq_stream_set_profiling": (C.c_int32, [vp, C.c_int32]),
    "sq_stream_phase_ms": (C.c_int32, [vp, f32p]),
    "sq_stream_launches": (C.c_uint64, [vp]),
    "sq_stream_submit": (C.c_int32, [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, u64p]),
    "sq_stream_set_key_dictionary": (C.c_int32, [vp, vp, C.c_uint32]),
    "sq_stream_submit_ids": (C.c_int32, [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, u64p]),
    "sq_stream_collect": (C.c_int32, [vp, C.c_uint64, C.POINTER(SqTileOut)]),
    "sq_stream_in_flight": (C.c_int32, [vp]),
    "sq_stream_pipeline_stats": (C.c_int32, [vp, C.POINTER(C.c_double)]),
}



class SqScanOptions(C.Structure):
    """struct sq_scan_options (include/sequila_scan.h)"""
    _fields_ = [("delimiter", C.c_uint8), ("has_header", C.c_uint8), ("comment", C.c_uint8), ("reserved", C.c_uint8),
                ("col_key", C.c_int32), ("col_start", C.c_int32), ("col_end", C.c_int32), ("reserved2", C.c_int32),
                ("start_minus", C.c_int64), ("end_minus", C.c_int64)]


# every symbol include/sequila_scan.h declares
SCAN_SIGNATURES = {
    "sq_scan_text": (C.c_int32, [vp, vp, C.c_uint64, C.POINTER(SqScanOptions), C.POINTER(vp)]),
    "sq_scan_text_device": (C.c_int32, [vp, vp, C.c_uint64, C.POINTER(SqScanOptions), C.POINTER(vp)]),
    "sq_scan_rows": (C.c_uint64, [vp]),
    "sq_scan_bytes": (C.c_uint64, [vp]),
    "sq_scan_key_hash_device": (vp, [vp]),
    "sq_scan_start_device": (vp, [vp]),
    "sq_scan_end_device": (vp, [vp]),
    "sq_scan_key_ids_device": (vp, [vp]),
    "sq_scan_dict_size": (C.c_uint32, [vp]),
    "sq_scan_dict_entry": (C.c_int32, [vp, C.c_uint32, C.POINTER(vp), u32p, u64p]),
    "sq_scan_fetch": (C.c_int32, [vp, vp, vp, vp, vp, vp]),
    "sq_scan_timing": (C.c_int32, [vp, f32p]),
    "sq_scan_free": (None, [vp]),
}


class SqDriveStats(C.Structure):
    """struct sq_drive_stats (include/sequila_driver.h)"""
    _fields_ = [("n_pairs", C.c_uint64), ("n_tiles", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("left_xor", C.c_uint64), ("regrown_tiles", C.c_uint64), ("seconds", C.c_double), ("h2d_ms", C.c_double),
                ("kernel_ms", C.c_double), ("d2h_ms", C.c_double)]


# every symbol include/sequila_driver.h declares
DRIVER_SIGNATURES = {
    "sq_driver_create": (C.c_int32, [vp, C.c_int32, C.POINTER(vp)]),
    "sq_driver_run": (C.c_int32, [vp, vp, vp, vp, vp, C.c_uint64, C.c_int32, C.c_uint32, C.c_int32, vp, vp,
                                  C.POINTER(SqDriveStats)]),
    "sq_driver_run_ids": (C.c_int32, [vp, vp, vp, C.c_uint32, vp, vp, vp, C.c_uint64, C.c_int32, C.c_uint32, C.c_int32, vp, vp,
                                      C.POINTER(SqDriveStats)]),
    "sq_driver_last_error": (C.c_char_p, [vp]),
    "sq_driver_free": (None, [vp]),
}


class SqExecConfig(C.Structure):
    """struct sq_exec_config (include/sequila_exec.h)"""
    _fields_ = [("device", C.c_int32), ("n_on", C.c_int32), ("on_left", i32p), ("on_right", i32p),
                ("left_start", C.c_int32), ("left_end", C.c_int32), ("right_start", C.c_int32),
                ("right_end", C.c_int32), ("left_end_minus_one", C.c_int32), ("right_end_minus_one", C.c_int32),
                ("n_projection", C.c_int32), ("projection", i32p), ("algorithm", C.c_int32),
                ("low_memory", C.c_int32), ("max_output_rows", C.c_int64)]


# every symbol include/sequila_exec.h declares (Arrow C Data Interface structs travel as addresses)
EXEC_SIGNATURES = {
    "sq_exec_create": (C.c_int32, [C.POINTER(SqExecConfig), vp, vp, C.POINTER(vp)]),
    "sq_exec_push_build": (C.c_int32, [vp, vp]),
    "sq_exec_finish_build": (C.c_int32, [vp]),
    "sq_exec_output_schema": (C.c_int32, [vp, vp]),
    "sq_exec_probe": (C.c_int32, [vp, C.c_int32, vp, vp]),
    "sq_exec_probe_push": (C.c_int32, [vp, C.c_int32, vp, i32p]),
    "sq_exec_probe_pop": (C.c_int32, [vp, C.c_int32, C.c_int32, vp, i32p]),
    "sq_exec_probe_begin": (C.c_int32, [vp, C.c_int32, vp]),
    "sq_exec_probe_next": (C.c_int32, [vp, C.c_int32, vp, i32p]),
    "sq_exec_metrics": (C.c_int32, [vp, u64p]),
    "sq_exec_last_error": (C.c_char_p, [vp]),
    "sq_exec_set_option": (C.c_int32, [vp, C.c_char_p, C.c_char_p]),
    "sq_exec_free": (None, [vp]),
}

_lib = None


class SequilaCudaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def lib():
    """Load the CUDA extension; raise loudly if it was not built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                " (make -C sequila_native_b200/csrc). The cuda interval join has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in list(SIGNATURES.items()) + list(EXEC_SIGNATURES.items()) + list(SCAN_SIGNATURES.items()) + list(DRIVER_SIGNATURES.items()):
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
