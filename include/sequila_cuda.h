/* sequila_cuda.h — C ABI of the B200-native interval-overlap join (`alg=Cuda`).
 *
 * The reference (biodatageeks/sequila-native) has no FFI for this path: its seam is the
 * private Rust enum `IntervalJoinAlgorithm` with `new()` / `get()`
 * (sequila/sequila-core/src/physical_planner/joins/interval_join.rs, "IJ" below, IJ:767 and
 * IJ:957-960).  A per-row callback is the wrong granularity for a GPU, so this ABI sits one
 * level up: it replaces the tail of `collect_left_input` (IJ:662-683) and the body of
 * `process_probe_batch` (IJ:1582-1632) with batch-level calls.  Every entry point names the
 * reference lines whose work it takes over.  INTEGRATION.md shows the Rust `-sys` binding.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 on success or an SQ_E* code,
 *     never aborts; the message is read with sq_last_error / sq_stream_last_error;
 *   - "host" entry points borrow caller memory for the duration of the call only;
 *   - `_device` entry points take device pointers (used by the kernel-level benchmark so
 *     that timing excludes PCIe) and enqueue on the sq_stream's CUDA stream;
 *   - there is no CPU fallback: without a usable CUDA device sq_ctx_create fails.
 */
#ifndef SEQUILA_CUDA_H
#define SEQUILA_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQ_OK 0
#define SQ_EINVAL 1    /* bad argument                                  */
#define SQ_ECUDA 2     /* CUDA runtime/driver error (message has detail) */
#define SQ_ENOMEM 3    /* device or pinned-host allocation failed       */
#define SQ_ESTATE 4    /* call order violated (emit without count, ...) */
#define SQ_ECAPACITY 5 /* caller buffer smaller than the result         */
#define SQ_ECAST 6     /* value does not fit Int32 (IJ:1661-1672)       */
#define SQ_EPARSE 7    /* malformed delimited text (sequila_scan.h)     */
#define SQ_EBUSY 8     /* every tile slot of the stream is in flight    */

#define SQ_ABI_VERSION 1

typedef struct sq_ctx sq_ctx;       /* per process+device: error slot, device id            */
typedef struct sq_index sq_index;   /* immutable build-side index (+ resident build columns) */
typedef struct sq_stream sq_stream; /* per DataFusion partition: CUDA stream, staging, scratch */

/* ---- context ------------------------------------------------------------------------- */
int32_t sq_abi_version(void);
/* device = CUDA ordinal.  Fails (SQ_ECUDA) when no usable device exists: no CPU fallback. */
int32_t sq_ctx_create(int32_t device, sq_ctx** out);
void sq_ctx_destroy(sq_ctx* ctx);
const char* sq_last_error(const sq_ctx* ctx);
int32_t sq_device_count(void);

/* Tuning options = the `sequila.cuda_*` session keys, set the way the reference sets its own knobs
 * (`SET sequila.<key> TO <value>` -> ConfigExtension::set, SC:106-132); the "sequila." prefix is optional.
 * Nothing in the library reads the process environment.
 *   cuda_probe_layout         auto | packed | soa   which probe kernels serve an index that has both layouts
 *   cuda_staged_probe         auto | on | off       shared-memory (TMA) staged kernel for position-local probe tiles
 *   cuda_probe_block          64 | 128 | 256        probe rows per CTA of the packed-line kernels
 *   cuda_probe_tiles          1 | 2                 tiles per CTA of an emitting packed-line launch (default 1; 2 measured slower)
 *   cuda_lookback_backoff_ns  integer               sleep between polls of the chained scan's look-back
 *   cuda_rows_per_bin         0..1024               build: target rows per directory bin (0 = automatic: 1 up to 4M rows, else 8)
 *   cuda_right_idx_wire       rle | copy            host entry points: right_idx crosses PCIe as per-row counts (default)
 *                                                   or as itself
 *   cuda_l2_persist_mb        integer               L2 set-aside for the probe directory (device-wide limit; default 0)
 *   cuda_scan_dict_capacity   power of two          text scan: initial key-dictionary capacity
 *   cuda_exec_trace           0 | 1                 exec node: per-phase wall times on stderr
 *   cuda_pipeline_depth       2..8                  sq_stream_submit: tiles in flight per stream (default 3)
 *   cuda_coalesce_rows        1..2^27               exec node: probe rows that make one tile (default 524288)
 *   cuda_rank_count           on | off              build: rank structure over the ends (rank-difference count, one walk)
 *   cuda_build_sort           auto | wide           build: 32-bit sort keys when they fit (see sq_index_sort_key_bits) or always 64-bit
 *   cuda_build_ids            rows | positions      build: what left_idx means (see sq_index_uses_positions; default rows)
 * Unknown keys and invalid values return SQ_EINVAL with a message; values may be changed between calls. */
int32_t sq_ctx_set_option(sq_ctx* ctx, const char* key, const char* value);
int32_t sq_ctx_get_option(sq_ctx* ctx, const char* key, char* value_out, size_t capacity);

/* Pinned host memory for Arrow buffers the host wants copied without a staging hop. */
int32_t sq_host_alloc(sq_ctx* ctx, size_t bytes, void** out);
void sq_host_free(sq_ctx* ctx, void* p);

/* ---- build side ------------------------------------------------------------------------
 * Replaces IJ:662-683: `update_hashmap` (bucket rows by the u64 key hash only, IJ:1042-1048)
 * + `IntervalJoinAlgorithm::new` (one coitrees tree per key, IJ:769-793).
 *   key_hash[i] = create_hashes(on_left, RandomState::with_seeds(0,0,0,0))   IJ:136, IJ:1037
 *   start[i], end[i] = evaluate_as_i32(left_interval.start()/end())          IJ:1039-1040
 *   row i  <->  left index i of the concatenated build batch                 IJ:685, IJ:1043
 * n_rows must be < 2^32 - 1 (the reference truncates `pos as u32`, IJ:1590).
 * The index is immutable after return and may be probed from many threads/streams. */
int32_t sq_index_build(sq_ctx* ctx, const uint64_t* key_hash, const int32_t* start,
                       const int32_t* end, uint64_t n_rows, sq_index** out);
/* Same, inputs already in device memory; `cuda_stream` is a cudaStream_t (NULL = default). */
int32_t sq_index_build_device(sq_ctx* ctx, const uint64_t* d_key_hash, const int32_t* d_start,
                              const int32_t* d_end, uint64_t n_rows, void* cuda_stream,
                              sq_index** out);
uint64_t sq_index_bytes(const sq_index* idx); /* device bytes held (build_mem_used gauge, IJ:631) */
uint64_t sq_index_rows(const sq_index* idx);
uint64_t sq_index_keys(const sq_index* idx);  /* distinct key hashes = number of per-key trees    */
/* which probe kernels serve this index: 1 = the fused kernel over packed lines (every width < 65536, index
 * far larger than L2 and shallow), 0 = count / scan / write over the SoA arrays */
int32_t sq_index_uses_packed(const sq_index* idx);
/* 1 = the index carries the rank structure over the ends (built when the SoA kernels serve it and every build row has
 * start <= end): the probe is ONE kernel whose count is a rank difference and whose only candidate walk writes the
 * pairs; 0 = count / scan / write, two walks */
int32_t sq_index_uses_rank(const sq_index* idx);
/* Position ids (option cuda_build_ids = positions, read when the index is built).  The left_idx values every probe entry
 * point hands out are then POSITIONS in the index's (key, start) order instead of build rows, and every payload column
 * registered with sq_index_add_*column / sq_index_set_validity is stored in that order (permuted once, on the device):
 * the hits of a probe row are neighbouring positions, so the `take` of the build side (IJ:1620-1632) reads neighbouring
 * payload rows instead of one random row per pair (measured: 0.66 vs 1.85 ms for 80.5M pairs x three packed columns).
 * The joined rows are the same and in the same order; only the meaning of left_idx changes.
 * sq_index_position_rows copies the map position -> build row (n_rows values) for callers that need the rows.
 * The exec node (sequila_exec.h) builds its index this way: its build table is internal. */
int32_t sq_index_uses_positions(const sq_index* idx);
int32_t sq_index_position_rows(const sq_index* idx, uint32_t* rows_out);
const uint32_t* sq_index_position_rows_device(const sq_index* idx); /* device pointer, NULL unless position ids */
/* width of the build's sort keys: 32 when the keys' start ranges laid end to end fit 32 bits (at most 4096 keys; a
 * genome's contigs do) — four radix passes over 12-byte pairs — else 64 (key id : start); option cuda_build_sort = wide
 * forces 64.  The index is the same either way. */
int32_t sq_index_sort_key_bits(const sq_index* idx);
/* device time of the last build's kernels in ms (sort, scan, ...); 0 if unknown */
float sq_index_build_ms(const sq_index* idx);
void sq_index_free(sq_index* idx);

/* Build-side columns to materialise later (IJ:1623-1624 `take(build.batch.column(k), left)`).
 * Fixed-width values of `width` bytes (4, 8 or 16) per row, n_rows of them; copied to the device. */
int32_t sq_index_add_column(sq_index* idx, const void* values, uint32_t width, int32_t* col_id_out);
int32_t sq_index_add_column_device(sq_index* idx, const void* d_values, uint32_t width,
                                   int32_t* col_id_out); /* borrowed, must outlive the index */

/* ---- probe side ------------------------------------------------------------------------
 * One sq_stream per partition (IJ:528-556 creates one IntervalJoinStream per partition);
 * not thread-safe by itself, but different sq_streams may probe one sq_index concurrently. */
int32_t sq_stream_create(sq_ctx* ctx, sq_stream** out);
/* Run on a caller-provided cudaStream_t (e.g. the host framework's current stream). */
int32_t sq_stream_create_on(sq_ctx* ctx, void* cuda_stream, sq_stream** out);
void sq_stream_free(sq_stream* s);
const char* sq_stream_last_error(const sq_stream* s);
uint64_t sq_stream_bytes(const sq_stream* s); /* device scratch + pinned staging held */

/* Phase 1 of `process_probe_batch` full mode (IJ:1582-1609) for one probe tile (>= 1 probe
 * batch, order preserved):
 *   key_hash[i] = create_hashes(on_right)      IJ:1202-1211
 *   start/end   = evaluate_as_i32(right ...)   IJ:1430-1431
 * Computes, on the device, for every probe row the number of build rows with equal key hash
 * and  build.start <= probe.end && build.end >= probe.start  (IV:95-137; coitrees
 * nosimd.rs:647-649); a key hash absent from the build side yields zero (IJ:965).
 * *n_pairs_out = total number of output rows, so the caller can allocate. */
int32_t sq_probe_count(sq_stream* s, const sq_index* idx, const uint64_t* key_hash,
                       const int32_t* start, const int32_t* end, uint32_t n_rows,
                       uint64_t* n_pairs_out);
/* Phase 2 (IJ:1604-1618): writes the pairs of the tile just counted.
 *   left_idx_out[k]  = build row (== `pos as u32`, IJ:1590); NULL = keep the pairs on the device
 *                      only (capacity is then ignored past n_pairs; sq_gather_* calls follow)
 *   right_idx_out[k] = probe row within this tile, non-decreasing (IJ:1611-1618); may be NULL
 *   counts_out[i]    = hits of probe row i (`rle_right`, IJ:1604); may be NULL
 * The order of left hits inside one probe row is unspecified (as in the reference, where it
 * is coitrees' traversal order and every test sorts, IJ:1808).
 * capacity = number of elements left_idx_out/right_idx_out can hold (>= n_pairs).
 * right_idx crosses PCIe run-length encoded (the per-row counts, 4 B per probe row instead of 4 B per pair)
 * and is expanded into right_idx_out by the calling thread while left_idx is still arriving — the expansion
 * loop the reference runs at IJ:1611-1618.  Option cuda_right_idx_wire=copy copies right_idx itself. */
int32_t sq_probe_emit_pairs(sq_stream* s, uint32_t* left_idx_out, uint32_t* right_idx_out,
                            uint32_t* counts_out, uint64_t capacity);

/* Both phases in one call: count -> scan -> write run as ONE fused kernel pass when the pairs fit
 * `capacity`.  *n_pairs_out is always exact on return.  If n_pairs > capacity the call returns
 * SQ_ECAPACITY having written nothing usable; allocate n_pairs elements and call
 * sq_probe_emit_pairs (the tile stays counted). */
int32_t sq_probe_join(sq_stream* s, const sq_index* idx, const uint64_t* key_hash,
                      const int32_t* start, const int32_t* end, uint32_t n_rows,
                      uint32_t* left_idx_out, uint32_t* right_idx_out, uint32_t* counts_out,
                      uint64_t capacity, uint64_t* n_pairs_out);

/* ---- asynchronous tile pipeline ------------------------------------------------------------
 * The reference's stream state machine handles one probe batch at a time inside poll_next (IJ:1054-1167,
 * 1192-1233); behind PCIe the three legs of a tile (H2D of the probe columns, kernels, D2H of the result) have
 * to overlap ACROSS tiles.  sq_stream_submit enqueues all three legs of one tile on the stream's copy-in,
 * compute and copy-out CUDA streams and returns at once; up to `cuda_pipeline_depth` (default 3) tiles may be
 * in flight per sq_stream (SQ_EBUSY beyond that).  sq_stream_collect waits for the OLDEST ticket and hands over
 * its result in pinned host buffers taken from the library's pool, fresh for every tile (sized from the fan-out
 * of the previous tile; a tile that outgrows them is re-emitted, so the call never fails for lack of room):
 *   left_idx[n_pairs]   build rows (`pos as u32`, IJ:1590)
 *   counts[n_rows]      hits per probe row (`rle_right`, IJ:1604); right_idx is its run-length expansion
 *                       (IJ:1611-1618) and by default does not cross PCIe at all
 *   right_idx           NULL unless SQ_TILE_RIGHT_IDX (copied from the device) or SQ_TILE_EXPAND_RIGHT
 *                       (expanded from the counts by the collecting thread) was requested
 * The caller owns the buffers and returns each with sq_host_free (e.g. from an Arrow buffer's release
 * callback).  The probe columns passed to submit must stay valid until the ticket has been collected; they
 * should be pinned (sq_host_alloc) for the copy-in to be asynchronous.  Tiles of one stream are collected in
 * submission order (probe order is preserved, IJ:210-218). */
#define SQ_TILE_COUNT_ONLY 1u   /* count(*) of the join: no pairs are written or moved                         */
#define SQ_TILE_RIGHT_IDX 2u    /* also copy right_idx from the device (8 B per pair on the wire instead of 4)  */
#define SQ_TILE_EXPAND_RIGHT 4u /* expand right_idx from the counts on the collecting thread                    */
#define SQ_TILE_NO_COUNTS 8u    /* do not move the per-row counts (count-only callers that want the total only) */
#define SQ_TILE_COUNTS_U8 16u   /* counts travel as one byte per probe row when every count of the tile is < 256 (a tenth of
                                 * the cfg5 result's bytes); a tile with a bigger count hands over 4-byte counts as usual:
                                 * read sq_tile_out.counts_width */
typedef struct sq_tile_out {
  uint64_t n_pairs;
  uint32_t n_rows;
  uint32_t counts_width;  /* bytes per element of `counts`: 4, or 1 (SQ_TILE_COUNTS_U8; `counts` then points at bytes); 0: no counts */
  uint32_t* left_idx;
  uint32_t* right_idx;
  uint32_t* counts;
} sq_tile_out;
int32_t sq_stream_submit(sq_stream* s, const sq_index* idx, const uint64_t* key_hash, const int32_t* start,
                         const int32_t* end, uint32_t n_rows, uint32_t flags, uint64_t* ticket_out);
/* 12 bytes per probe row on the wire instead of 16: the key column travels as 4-byte ids into a dictionary of key hashes
 * (what a dictionary-encoded contig column already is, queries/q1-coitrees.sql:6-14; hash of dictionary value k =
 * key_hashes[k], the same hash the build side was given).  The dictionary is uploaded once per stream; an id outside it
 * — SQ_NULL_INDEX for a NULL key — matches nothing (create_hashes leaves NULL slots out, IJ:1211).  Everything else as
 * sq_stream_submit. */
int32_t sq_stream_set_key_dictionary(sq_stream* s, const uint64_t* key_hashes, uint32_t n_entries);
int32_t sq_stream_submit_ids(sq_stream* s, const sq_index* idx, const uint32_t* key_id, const int32_t* start,
                             const int32_t* end, uint32_t n_rows, uint32_t flags, uint64_t* ticket_out);
int32_t sq_stream_collect(sq_stream* s, uint64_t ticket, sq_tile_out* out);
int32_t sq_stream_in_flight(const sq_stream* s);
/* gauges of the pipeline since the stream was created (the reference's BuildProbeJoinMetrics, utils.rs:441-495,
 * has join_time only): [0..2] device ms summed over tiles of copy-in / kernels / copy-out, [3] bytes copied in,
 * [4] bytes copied out, [5] tiles collected, [6] tiles re-emitted into larger buffers, [7] pairs per probe row
 * of the last tile */
int32_t sq_stream_pipeline_stats(const sq_stream* s, double out8[8]);

/* Algorithm::CoitreesNearest on the same index (IJ:794-812 build, IJ:909-956 `nearest`, IJ:972-990
 * `get`, IJ:1593-1602 emit): ONE output row per probe row.
 *   left_idx_out[i] = a build row overlapping probe row i when one exists (the reference reports the
 *                     first one its tree traversal visits, "an arbitrary one", IJ:976; here: the
 *                     overlapping row that is last in (start) order), else the row `nearest()` picks
 *                     (two candidates around `end`, earlier one wins ties), else SQ_NULL_INDEX when the
 *                     key hash never occurred on the build side (NULL left side, IJ:1597-1598).
 *   right index of output row i is i.  May be NULL: the result then stays on the device for sq_gather_*
 *   (fixed-width gathers write zero, Utf8 gathers an empty string, sq_gather_validity a cleared bit for
 *   SQ_NULL_INDEX). */
#define SQ_NULL_INDEX 0xFFFFFFFFu
int32_t sq_probe_nearest(sq_stream* s, const sq_index* idx, const uint64_t* key_hash,
                         const int32_t* start, const int32_t* end, uint32_t n_rows,
                         uint32_t* left_idx_out);
int32_t sq_probe_nearest_device(sq_stream* s, const sq_index* idx, const uint64_t* d_key_hash,
                                const int32_t* d_start, const int32_t* d_end, uint32_t n_rows,
                                uint32_t* d_left_idx_out);

/* Device-pointer variants (kernel-level benchmark; outputs stay in HBM).  The count and join
 * variants synchronise the stream to return n_pairs; probe columns must stay valid until the
 * tile's emit. */
int32_t sq_probe_join_device(sq_stream* s, const sq_index* idx, const uint64_t* d_key_hash,
                             const int32_t* d_start, const int32_t* d_end, uint32_t n_rows,
                             uint32_t* d_left_idx_out, uint32_t* d_right_idx_out,
                             uint64_t capacity, uint64_t* n_pairs_out);
int32_t sq_probe_count_device(sq_stream* s, const sq_index* idx, const uint64_t* d_key_hash,
                              const int32_t* d_start, const int32_t* d_end, uint32_t n_rows,
                              uint64_t* n_pairs_out);
int32_t sq_probe_emit_pairs_device(sq_stream* s, uint32_t* d_left_idx_out,
                                   uint32_t* d_right_idx_out, uint64_t capacity);
/* device pointer to the per-probe-row counts of the last sq_probe_count* on this stream */
const uint32_t* sq_stream_counts_device(const sq_stream* s);

/* ---- materialise (IJ:1620-1632: one arrow::compute::take per output column) -------------
 * Gathers fixed-width values for the pairs of the last emit on this stream, on the device.
 * side 0 = build column registered with sq_index_add_column (indexed by left_idx),
 * side 1 = probe column of `width`-byte values for this tile (indexed by right_idx).
 * Host variant: `probe_values` (side 1 only) and `out` are host pointers. */
int32_t sq_gather_column(sq_stream* s, int32_t side, int32_t build_col_id,
                         const void* probe_values, uint32_t width, void* out, uint64_t capacity);
int32_t sq_gather_column_device(sq_stream* s, int32_t side, int32_t build_col_id,
                                const void* d_probe_values, uint32_t width, void* d_out,
                                uint64_t capacity);

/* Several columns per pass.  `take` column by column (what IJ:1620-1632 does on the CPU) costs the GPU one
 * random 32-byte sector per pair AND per build column, plus one more read of the index array; measured on
 * the cfg5 shard the six-column output took 5.6 ms that way against 1.0 ms for the join itself.
 * sq_index_pack_columns interleaves up to four 4-byte build columns row-wise (16 bytes per build row, built
 * once per query on the device) so that ONE random read per pair serves all of them;
 * sq_gather_pack_device writes the pack's columns for the pairs of the last emit into d_outs[0..n_cols).
 * sq_gather_probe_columns_device does the same for up to four 4-byte probe columns of this tile (right_idx
 * is non-decreasing: sequential reads, one pass instead of one per column). */
int32_t sq_index_pack_columns(sq_index* idx, const int32_t* col_ids, int32_t n_cols, int32_t* pack_id_out);
int32_t sq_gather_pack_device(sq_stream* s, int32_t pack_id, void* const* d_outs, int32_t n_outs,
                              uint64_t capacity);
int32_t sq_gather_probe_columns_device(sq_stream* s, const void* const* d_probe_values, void* const* d_outs,
                                       int32_t n_cols, uint64_t capacity);

/* Utf8 columns (Arrow `take` of a string column, IJ:1624-1627).  Build side: offsets are n_rows+1
 * int64 (the host concatenates the build batches, IJ:685).  Gathering is two-phase so the caller
 * can allocate: sq_gather_utf8 computes the n_pairs+1 output offsets (32-bit, Arrow Utf8) and the
 * byte total; sq_gather_utf8_data then copies the bytes.  Probe side (side 1): this tile's column
 * is passed with int64 offsets. */
int32_t sq_index_add_utf8_column(sq_index* idx, const int64_t* offsets, const uint8_t* data,
                                 uint64_t data_bytes, int32_t* col_id_out);
int32_t sq_gather_utf8(sq_stream* s, int32_t side, int32_t build_col_id, const int64_t* probe_offsets,
                       const uint8_t* probe_data, uint64_t probe_data_bytes, int32_t* out_offsets,
                       uint64_t* total_bytes_out);
int32_t sq_gather_utf8_data(sq_stream* s, uint8_t* out_data, uint64_t capacity);
/* Arrow validity bitmaps of payload columns: out bit k = in bit idx[k] (a build column without a bitmap
 * counts as all-valid), 0 where idx[k] == SQ_NULL_INDEX; *null_count_out = zeros. */
int32_t sq_index_set_validity(sq_index* idx, int32_t col_id, const uint8_t* bitmap);
int32_t sq_gather_validity(sq_stream* s, int32_t side, int32_t build_col_id, const uint8_t* probe_bitmap,
                           uint8_t* out_bitmap, uint64_t* null_count_out);

/* ---- bounded output --------------------------------------------------------------------------
 * The GPU analogue of `sequila.interval_join_low_memory` (IJ:1433-1530: the reference caps an output batch
 * at 1M rows and continues with the remaining probe rows): the whole tile is emitted ONCE into device
 * memory (sq_probe_emit_pairs with NULL outputs), the host then walks it in windows of the pair sequence —
 * cut at probe-row boundaries with the counts of sq_stream_counts — and only one window at a time is
 * gathered / copied to the host.  right_idx stays relative to the tile. */
int32_t sq_stream_counts(sq_stream* s, uint32_t* counts_out /* n_rows of the counted tile: rle_right */);
int32_t sq_stream_set_window(sq_stream* s, uint64_t pair_offset, uint64_t n_pairs);
/* D2H copy of the current window (the whole tile if none was set); either output may be NULL */
int32_t sq_fetch_pairs(sq_stream* s, uint32_t* left_idx_out, uint32_t* right_idx_out, uint64_t capacity);

/* ---- helpers on the boundary -------------------------------------------------------------
 * `evaluate_as_i32` for BIGINT columns (IJ:1661-1672): checked cast Int64 -> Int32 on the
 * device, optionally subtracting `minus` first (the `end - 1` of strict comparisons,
 * IV:67-69).  On overflow returns SQ_ECAST and the message is exactly the reference's
 * "Arrow error: Cast error: Can't cast value {v} to type Int32" (IJ:1959-1965), v being the
 * first offending value in row order. */
int32_t sq_cast_i64_to_i32(sq_stream* s, const int64_t* values, uint64_t n, int64_t minus,
                           int32_t* out);

/* Order-independent multiset digest of (left,right+right_offset) pairs held in device memory:
 * out3 = {count, sum of mix64(left<<32|right) mod 2^64, xor of the same}.  Parity at scale. */
int32_t sq_pairs_digest_device(sq_stream* s, const uint32_t* d_left, const uint32_t* d_right,
                               uint64_t n_pairs, uint64_t right_offset, uint64_t out3[3]);

/* Test / tuning hook of the host-side expansion of right_idx from the per-row counts (the wire format of the
 * host entry points, IJ:1611-1618): runs ONE named variant (-1 = the widest the CPU supports = what the library
 * uses, 0 scalar, 1 SSE2, 2 AVX2, 3 AVX-512); returns 0 when the CPU lacks it, else 1.  No device involved. */
int32_t sq_rle_expand_variant(int32_t variant, const uint32_t* counts, uint32_t n_rows, uint32_t* right_idx_out,
                              uint64_t n_pairs);

/* Per-phase device timings (ms), averaged over the calls since sq_stream_set_profiling(s, 1),
 * measured with CUDA events recorded on the stream around each phase:
 * [0]=h2d [1]=fused probe kernel of count/join calls [2]=probe kernel re-run by emit calls
 * [3]=d2h [4]=gather.  Enabling adds only event
 * records to the stream; sq_stream_phase_ms synchronises the stream to read them. */
int32_t sq_stream_set_profiling(sq_stream* s, int32_t enabled);
int32_t sq_stream_phase_ms(sq_stream* s, float out5[5]);
/* number of kernels this library launched on the stream since creation (bench: gpu_launches) */
uint64_t sq_stream_launches(const sq_stream* s);

#ifdef __cplusplus
}
#endif
#endif /* SEQUILA_CUDA_H */
