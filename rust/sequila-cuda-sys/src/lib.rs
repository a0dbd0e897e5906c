//! UNCOMPILED — see rust/README.md.
//! Raw bindings of `include/sequila_cuda.h`, `sequila_exec.h`, `sequila_driver.h` and `sequila_scan.h`
//! (ABI version 1).  One item per C declaration, same order as the headers; `tests/test_abi.py` keeps the
//! ctypes twin of this list (`sequila_native_b200/_native.py`) equal to the headers and to the exports of the
//! library, so a drift shows up there first.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_void};

#[repr(C)] pub struct sq_ctx { _p: [u8; 0] }
#[repr(C)] pub struct sq_index { _p: [u8; 0] }
#[repr(C)] pub struct sq_stream { _p: [u8; 0] }
#[repr(C)] pub struct sq_exec { _p: [u8; 0] }
#[repr(C)] pub struct sq_scan { _p: [u8; 0] }
#[repr(C)] pub struct sq_driver { _p: [u8; 0] }

pub const SQ_OK: i32 = 0;
pub const SQ_EINVAL: i32 = 1;
pub const SQ_ECUDA: i32 = 2;
pub const SQ_ENOMEM: i32 = 3;
pub const SQ_ESTATE: i32 = 4;
pub const SQ_ECAPACITY: i32 = 5;
pub const SQ_ECAST: i32 = 6;
pub const SQ_EPARSE: i32 = 7;
pub const SQ_EBUSY: i32 = 8;
pub const SQ_ABI_VERSION: i32 = 1;
pub const SQ_NULL_INDEX: u32 = 0xFFFF_FFFF;

pub const SQ_TILE_COUNT_ONLY: u32 = 1;
pub const SQ_TILE_RIGHT_IDX: u32 = 2;
pub const SQ_TILE_EXPAND_RIGHT: u32 = 4;
pub const SQ_TILE_NO_COUNTS: u32 = 8;
pub const SQ_TILE_COUNTS_U8: u32 = 16;

/// `struct sq_tile_out`: result of one collected tile; the three buffers are pinned host memory owned by the
/// caller after `sq_stream_collect` and go back to the pool with `sq_host_free`.
#[repr(C)]
pub struct sq_tile_out {
    pub n_pairs: u64,
    pub n_rows: u32,
    pub counts_width: u32, // bytes per element of `counts`: 4, or 1 with SQ_TILE_COUNTS_U8; 0: no counts
    pub left_idx: *mut u32,
    pub right_idx: *mut u32,
    pub counts: *mut u32,
}

#[repr(C)]
pub struct sq_drive_stats {
    pub n_pairs: u64, pub n_tiles: u64, pub h2d_bytes: u64, pub d2h_bytes: u64, pub left_xor: u64, pub regrown_tiles: u64,
    pub seconds: f64, pub h2d_ms: f64, pub kernel_ms: f64, pub d2h_ms: f64,
}
pub type sq_tile_consumer = Option<unsafe extern "C" fn(user: *mut c_void, partition: i32, first_row: u64, tile: *const sq_tile_out)>;

extern "C" {
    // ---- context -------------------------------------------------------------------------------------
    pub fn sq_abi_version() -> i32;
    pub fn sq_ctx_create(device: i32, out: *mut *mut sq_ctx) -> i32;
    pub fn sq_ctx_destroy(ctx: *mut sq_ctx);
    pub fn sq_last_error(ctx: *const sq_ctx) -> *const c_char;
    pub fn sq_device_count() -> i32;
    pub fn sq_ctx_set_option(ctx: *mut sq_ctx, key: *const c_char, value: *const c_char) -> i32;
    pub fn sq_ctx_get_option(ctx: *mut sq_ctx, key: *const c_char, value_out: *mut c_char, capacity: usize) -> i32;
    pub fn sq_host_alloc(ctx: *mut sq_ctx, bytes: usize, out: *mut *mut c_void) -> i32;
    pub fn sq_host_free(ctx: *mut sq_ctx, p: *mut c_void);

    // ---- build side (replaces interval_join.rs:662-683) ------------------------------------------------
    pub fn sq_index_build(ctx: *mut sq_ctx, key_hash: *const u64, start: *const i32, end: *const i32, n_rows: u64,
                          out: *mut *mut sq_index) -> i32;
    pub fn sq_index_build_device(ctx: *mut sq_ctx, d_key_hash: *const u64, d_start: *const i32, d_end: *const i32, n_rows: u64,
                                 cuda_stream: *mut c_void, out: *mut *mut sq_index) -> i32;
    pub fn sq_index_bytes(idx: *const sq_index) -> u64;
    pub fn sq_index_rows(idx: *const sq_index) -> u64;
    pub fn sq_index_keys(idx: *const sq_index) -> u64;
    pub fn sq_index_uses_packed(idx: *const sq_index) -> i32;
    pub fn sq_index_uses_rank(idx: *const sq_index) -> i32;
    pub fn sq_index_sort_key_bits(idx: *const sq_index) -> i32;
    pub fn sq_index_uses_positions(idx: *const sq_index) -> i32;
    pub fn sq_index_position_rows(idx: *const sq_index, rows_out: *mut u32) -> i32;
    pub fn sq_index_position_rows_device(idx: *const sq_index) -> *const u32;
    pub fn sq_index_build_ms(idx: *const sq_index) -> f32;
    pub fn sq_index_free(idx: *mut sq_index);
    pub fn sq_index_add_column(idx: *mut sq_index, values: *const c_void, width: u32, col_id_out: *mut i32) -> i32;
    pub fn sq_index_add_column_device(idx: *mut sq_index, d_values: *const c_void, width: u32, col_id_out: *mut i32) -> i32;

    // ---- probe side (replaces interval_join.rs:1582-1618) ------------------------------------------------
    pub fn sq_stream_create(ctx: *mut sq_ctx, out: *mut *mut sq_stream) -> i32;
    pub fn sq_stream_create_on(ctx: *mut sq_ctx, cuda_stream: *mut c_void, out: *mut *mut sq_stream) -> i32;
    pub fn sq_stream_free(s: *mut sq_stream);
    pub fn sq_stream_last_error(s: *const sq_stream) -> *const c_char;
    pub fn sq_stream_bytes(s: *const sq_stream) -> u64;
    pub fn sq_probe_count(s: *mut sq_stream, idx: *const sq_index, key_hash: *const u64, start: *const i32, end: *const i32,
                          n_rows: u32, n_pairs_out: *mut u64) -> i32;
    pub fn sq_probe_emit_pairs(s: *mut sq_stream, left_idx_out: *mut u32, right_idx_out: *mut u32, counts_out: *mut u32,
                               capacity: u64) -> i32;
    pub fn sq_probe_join(s: *mut sq_stream, idx: *const sq_index, key_hash: *const u64, start: *const i32, end: *const i32,
                         n_rows: u32, left_idx_out: *mut u32, right_idx_out: *mut u32, counts_out: *mut u32, capacity: u64,
                         n_pairs_out: *mut u64) -> i32;
    // asynchronous tile pipeline
    pub fn sq_stream_submit(s: *mut sq_stream, idx: *const sq_index, key_hash: *const u64, start: *const i32, end: *const i32,
                            n_rows: u32, flags: u32, ticket_out: *mut u64) -> i32;
    pub fn sq_stream_set_key_dictionary(s: *mut sq_stream, key_hashes: *const u64, n_entries: u32) -> i32;
    pub fn sq_stream_submit_ids(s: *mut sq_stream, idx: *const sq_index, key_id: *const u32, start: *const i32, end: *const i32,
                                n_rows: u32, flags: u32, ticket_out: *mut u64) -> i32;
    pub fn sq_stream_collect(s: *mut sq_stream, ticket: u64, out: *mut sq_tile_out) -> i32;
    pub fn sq_stream_in_flight(s: *const sq_stream) -> i32;
    pub fn sq_stream_pipeline_stats(s: *const sq_stream, out8: *mut f64) -> i32;
    // nearest
    pub fn sq_probe_nearest(s: *mut sq_stream, idx: *const sq_index, key_hash: *const u64, start: *const i32, end: *const i32,
                            n_rows: u32, left_idx_out: *mut u32) -> i32;
    pub fn sq_probe_nearest_device(s: *mut sq_stream, idx: *const sq_index, d_key_hash: *const u64, d_start: *const i32,
                                   d_end: *const i32, n_rows: u32, d_left_idx_out: *mut u32) -> i32;
    // device-pointer variants
    pub fn sq_probe_join_device(s: *mut sq_stream, idx: *const sq_index, d_key_hash: *const u64, d_start: *const i32,
                                d_end: *const i32, n_rows: u32, d_left_idx_out: *mut u32, d_right_idx_out: *mut u32,
                                capacity: u64, n_pairs_out: *mut u64) -> i32;
    pub fn sq_probe_count_device(s: *mut sq_stream, idx: *const sq_index, d_key_hash: *const u64, d_start: *const i32,
                                 d_end: *const i32, n_rows: u32, n_pairs_out: *mut u64) -> i32;
    pub fn sq_probe_emit_pairs_device(s: *mut sq_stream, d_left_idx_out: *mut u32, d_right_idx_out: *mut u32, capacity: u64) -> i32;
    pub fn sq_stream_counts_device(s: *const sq_stream) -> *const u32;

    // ---- materialise (interval_join.rs:1620-1632) ------------------------------------------------------------
    pub fn sq_gather_column(s: *mut sq_stream, side: i32, build_col_id: i32, probe_values: *const c_void, width: u32,
                            out: *mut c_void, capacity: u64) -> i32;
    pub fn sq_gather_column_device(s: *mut sq_stream, side: i32, build_col_id: i32, d_probe_values: *const c_void, width: u32,
                                   d_out: *mut c_void, capacity: u64) -> i32;
    pub fn sq_index_pack_columns(idx: *mut sq_index, col_ids: *const i32, n_cols: i32, pack_id_out: *mut i32) -> i32;
    pub fn sq_gather_pack_device(s: *mut sq_stream, pack_id: i32, d_outs: *const *mut c_void, n_outs: i32, capacity: u64) -> i32;
    pub fn sq_gather_probe_columns_device(s: *mut sq_stream, d_probe_values: *const *const c_void, d_outs: *const *mut c_void,
                                          n_cols: i32, capacity: u64) -> i32;
    pub fn sq_index_add_utf8_column(idx: *mut sq_index, offsets: *const i64, data: *const u8, data_bytes: u64,
                                    col_id_out: *mut i32) -> i32;
    pub fn sq_gather_utf8(s: *mut sq_stream, side: i32, build_col_id: i32, probe_offsets: *const i64, probe_data: *const u8,
                          probe_data_bytes: u64, out_offsets: *mut i32, total_bytes_out: *mut u64) -> i32;
    pub fn sq_gather_utf8_data(s: *mut sq_stream, out_data: *mut u8, capacity: u64) -> i32;
    pub fn sq_index_set_validity(idx: *mut sq_index, col_id: i32, bitmap: *const u8) -> i32;
    pub fn sq_gather_validity(s: *mut sq_stream, side: i32, build_col_id: i32, probe_bitmap: *const u8, out_bitmap: *mut u8,
                              null_count_out: *mut u64) -> i32;

    // ---- bounded output (interval_join.rs:1433-1530) ------------------------------------------------------------
    pub fn sq_stream_counts(s: *mut sq_stream, counts_out: *mut u32) -> i32;
    pub fn sq_stream_set_window(s: *mut sq_stream, pair_offset: u64, n_pairs: u64) -> i32;
    pub fn sq_fetch_pairs(s: *mut sq_stream, left_idx_out: *mut u32, right_idx_out: *mut u32, capacity: u64) -> i32;

    // ---- helpers -----------------------------------------------------------------------------------------------
    pub fn sq_cast_i64_to_i32(s: *mut sq_stream, values: *const i64, n: u64, minus: i64, out: *mut i32) -> i32;
    pub fn sq_pairs_digest_device(s: *mut sq_stream, d_left: *const u32, d_right: *const u32, n_pairs: u64, right_offset: u64,
                                  out3: *mut u64) -> i32;
    pub fn sq_rle_expand_variant(variant: i32, counts: *const u32, n_rows: u32, right_idx_out: *mut u32, n_pairs: u64) -> i32;
    pub fn sq_stream_set_profiling(s: *mut sq_stream, enabled: i32) -> i32;
    pub fn sq_stream_phase_ms(s: *mut sq_stream, out5: *mut f32) -> i32;
    pub fn sq_stream_launches(s: *const sq_stream) -> u64;

    // ---- include/sequila_driver.h ---------------------------------------------------------------------------------
    pub fn sq_driver_create(ctx: *mut sq_ctx, n_partitions: i32, out: *mut *mut sq_driver) -> i32;
    pub fn sq_driver_run(d: *mut sq_driver, idx: *const sq_index, key_hash: *const u64, start: *const i32, end: *const i32,
                         n_rows: u64, n_tiles: i32, flags: u32, checksum: i32, consume: sq_tile_consumer, user: *mut c_void,
                         stats_out: *mut sq_drive_stats) -> i32;
    pub fn sq_driver_run_ids(d: *mut sq_driver, idx: *const sq_index, dict_key_hashes: *const u64, dict_entries: u32,
                             key_id: *const u32, start: *const i32, end: *const i32, n_rows: u64, n_tiles: i32, flags: u32,
                             checksum: i32, consume: sq_tile_consumer, user: *mut c_void, stats_out: *mut sq_drive_stats) -> i32;
    pub fn sq_driver_last_error(d: *const sq_driver) -> *const c_char;
    pub fn sq_driver_free(d: *mut sq_driver);
}

/// `include/sequila_exec.h`: the exec node over the Arrow C Data Interface (arrow-rs: `arrow::ffi::{FFI_ArrowArray,
/// FFI_ArrowSchema}` are layout-compatible with the structs the header declares).
pub mod exec {
    use super::*;
    #[repr(C)]
    pub struct sq_exec_config {
        pub device: i32, pub n_on: i32, pub on_left: *const i32, pub on_right: *const i32,
        pub left_start: i32, pub left_end: i32, pub right_start: i32, pub right_end: i32,
        pub left_end_minus_one: i32, pub right_end_minus_one: i32, pub n_projection: i32, pub projection: *const i32,
        pub algorithm: i32, pub low_memory: i32, pub max_output_rows: i64,
    }
    pub const SQ_EXEC_OVERLAPS: i32 = 0;
    pub const SQ_EXEC_NEAREST: i32 = 1;
    extern "C" {
        pub fn sq_exec_create(cfg: *const sq_exec_config, left_schema: *const c_void, right_schema: *const c_void,
                              out: *mut *mut sq_exec) -> i32;
        pub fn sq_exec_push_build(e: *mut sq_exec, batch: *mut c_void) -> i32;
        pub fn sq_exec_finish_build(e: *mut sq_exec) -> i32;
        pub fn sq_exec_output_schema(e: *const sq_exec, out: *mut c_void) -> i32;
        pub fn sq_exec_probe(e: *mut sq_exec, partition: i32, batch: *const c_void, out: *mut c_void) -> i32;
        pub fn sq_exec_probe_push(e: *mut sq_exec, partition: i32, batch: *mut c_void, ready_out: *mut i32) -> i32;
        pub fn sq_exec_probe_pop(e: *mut sq_exec, partition: i32, flush: i32, out: *mut c_void, has_out: *mut i32) -> i32;
        pub fn sq_exec_probe_begin(e: *mut sq_exec, partition: i32, batch: *const c_void) -> i32;
        pub fn sq_exec_probe_next(e: *mut sq_exec, partition: i32, out: *mut c_void, has_more_out: *mut i32) -> i32;
        pub fn sq_exec_metrics(e: *const sq_exec, out16: *mut u64) -> i32;
        pub fn sq_exec_last_error(e: *const sq_exec) -> *const c_char;
        pub fn sq_exec_set_option(e: *mut sq_exec, key: *const c_char, value: *const c_char) -> i32;
        pub fn sq_exec_free(e: *mut sq_exec);
    }
}

/// `include/sequila_scan.h`: delimited text -> device columns.
pub mod scan {
    use super::*;
    #[repr(C)]
    pub struct sq_scan_options {
        pub delimiter: u8, pub has_header: u8, pub comment: u8, pub reserved: u8,
        pub col_key: i32, pub col_start: i32, pub col_end: i32, pub reserved2: i32, pub start_minus: i64, pub end_minus: i64,
    }
    extern "C" {
        pub fn sq_scan_text(s: *mut sq_stream, text: *const u8, n_bytes: u64, opt: *const sq_scan_options, out: *mut *mut sq_scan) -> i32;
        pub fn sq_scan_text_device(s: *mut sq_stream, d_text: *const u8, n_bytes: u64, opt: *const sq_scan_options,
                                   out: *mut *mut sq_scan) -> i32;
        pub fn sq_scan_rows(sc: *const sq_scan) -> u64;
        pub fn sq_scan_bytes(sc: *const sq_scan) -> u64;
        pub fn sq_scan_key_hash_device(sc: *const sq_scan) -> *const u64;
        pub fn sq_scan_start_device(sc: *const sq_scan) -> *const i32;
        pub fn sq_scan_end_device(sc: *const sq_scan) -> *const i32;
        pub fn sq_scan_key_ids_device(sc: *const sq_scan) -> *const u32;
        pub fn sq_scan_dict_size(sc: *const sq_scan) -> u32;
        pub fn sq_scan_dict_entry(sc: *const sq_scan, i: u32, bytes_out: *mut *const u8, len_out: *mut u32, hash_out: *mut u64) -> i32;
        pub fn sq_scan_fetch(s: *mut sq_stream, sc: *const sq_scan, key_hash_out: *mut u64, start_out: *mut i32, end_out: *mut i32,
                             key_ids_out: *mut u32) -> i32;
        pub fn sq_scan_timing(sc: *const sq_scan, out: *mut f32) -> i32;
        pub fn sq_scan_free(sc: *mut sq_scan);
    }
}
