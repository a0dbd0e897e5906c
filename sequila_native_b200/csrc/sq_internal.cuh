// Internal declarations shared by the translation units of libsequila_cuda.so.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "sequila_cuda.h"

namespace sq {

constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;  // hash-table sentinel (handled out of band)
constexpr uint32_t kNoKey = 0xFFFFFFFFu;               // probe key hash absent from the build side
constexpr int kProbeBlock = 256;                       // threads per probe CTA = probe rows per CTA
constexpr int kWarpsPerBlock = kProbeBlock / 32;

// 64-bit finalizer used for the key hash table and the pair digest.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31; return x;
}

// Per key segment: where it lives in the sorted arrays and its direct-address bin directory.
// bin(x) = (x - min_start) >> shift ; dir[dir_base + b] = first row of the segment whose bin >= b,
// dir[dir_base + nbins] = se.  Bins hold 8-16 rows for uniform data, so locating the upper
// bound of a probe costs one directory load plus one round of loads inside one cache line.
struct SegMeta {
  uint32_t sb, se;     // [sb, se) rows of this key in the sorted arrays
  int32_t min_start;   // start[sb]
  uint32_t shift;
  uint32_t dir_base;
  uint32_t nbins;
  uint32_t line_base;  // first packed line of this key (packed layout, see PackedLine)
  uint32_t pad1;
};

// Packed layout of the same sorted rows ("narrow" indexes: 0 <= end - start < 65536 for every row):
// 128-byte lines of up to 15 rows; a line never crosses a key segment or a 65536-wide window of starts,
// so in-line start offsets fit 16 bits whatever the gaps in the data.  One cooperative 8-lane load (8 x 16 bytes = one L2 request) brings
//   lane 0      : { base_start, exmax, row0.lo, row0.id }
//   lanes 1..7  : { row(2k-1).lo, row(2k-1).id, row(2k).lo, row(2k).id }
// with row.lo = (start - base_start) | (end - start) << 16, row.id = build row (0xFFFFFFFF = empty
// slot) and exmax = max end over all earlier rows of the key segment (INT32_MIN on its first line):
// a probe (qs, qe) walks lines backwards from the one holding the last start <= qe while
// exmax >= qs, testing the exact predicate per row.
constexpr uint32_t kLineRows = 15;
constexpr uint32_t kEmptyRow = 0xFFFFFFFFu;

// Flat, tree-free build index in HBM (all arrays length n_rows, sorted by (key id, start)).
struct IndexView {
  const int32_t* __restrict__ start;    // sorted starts (searched)
  const int32_t* __restrict__ runmax;   // running max of end inside the key segment (non-decreasing)
  const int32_t* __restrict__ end;      // end of the same row
  const uint32_t* __restrict__ row;     // original build row (left index)
  const SegMeta* __restrict__ meta;     // [n_keys]
  const uint32_t* __restrict__ dir;     // bin directory, all segments back to back
  const uint64_t* __restrict__ ht_keys; // open-addressing table of key hashes
  const uint32_t* __restrict__ ht_ids;  // slot -> dense key id
  uint32_t ht_mask;                     // capacity - 1
  uint32_t sentinel_id;                 // id of the key hash equal to kEmptyKey, or kNoKey
  uint32_t n_keys;
  uint32_t n_rows;
  const uint4* __restrict__ lines;      // packed lines (nullptr when the index is not "narrow")
  const uint2* __restrict__ dir_line;   // directory entry -> {line holding the last row below it, that line's first start}
  // rank structure over the ENDS (nullptr when not built): the same rows' ends sorted inside each key segment, with
  // their own bin directory.  For rows with start <= end and a probe with qs <= qe + 1 the hit count is a rank
  // difference, |{start <= qe}| - |{end < qs}|, no candidate is touched to count
  // max of end over aligned blocks of 32 / 1024 / 32768 sorted rows (blocks ignore key segments: a superset test).  One
  // very long build interval keeps runmax >= qs for every later row of its key, so candidate ranges grow to O(segment);
  // walks over long ranges descend this pyramid and touch only the blocks that hold a hit (coitrees prunes the same
  // way through subtree_last, CT/nosimd.rs:343-384)
  const int32_t* __restrict__ bmax1;
  const int32_t* __restrict__ bmax2;
  const int32_t* __restrict__ bmax3;
  const int32_t* __restrict__ send;
  const SegMeta* __restrict__ emeta;    // min_start = smallest end of the segment; sb / se as in meta
  const uint32_t* __restrict__ edir;
};

#ifdef __CUDACC__
// key hash -> dense key id; kNoKey when the hash never occurred on the build side (IJ:965)
__device__ __forceinline__ uint32_t ht_lookup(const uint64_t* __restrict__ ht_keys,
                                              const uint32_t* __restrict__ ht_ids, uint32_t mask,
                                              uint32_t sentinel_id, uint64_t key) {
  if (key == kEmptyKey) return sentinel_id;
  uint32_t slot = uint32_t(mix64(key)) & mask;
  for (uint32_t step = 0; step <= mask; ++step) {
    const uint64_t cur = __ldg(ht_keys + slot);
    if (cur == key) return __ldg(ht_ids + slot);
    if (cur == kEmptyKey) return kNoKey;
    slot = (slot + 1) & mask;
  }
  return kNoKey;
}
#endif

struct ErrorSlot {
  mutable std::mutex mu;
  std::string msg;
  void set(const std::string& m) { std::lock_guard<std::mutex> g(mu); msg = m; }
};

}  // namespace sq

// Pool behind sq_host_alloc / sq_host_free: cudaHostAlloc costs milliseconds per call, and the exec node
// hands one pinned buffer per output column per batch to Arrow, so released buffers are kept (size classes
// of powers of two) and reused.
struct HostPool {
  std::mutex mu;
  std::multimap<size_t, void*> idle;          // capacity -> buffer
  std::unordered_map<void*, size_t> capacity; // every buffer this pool handed out
  size_t idle_bytes = 0;
};

// Tuning options of a context = the `sequila.cuda_*` session keys (sq_ctx_set_option; the reference keeps its
// knobs in SequilaConfig, session_context.rs:50-60).  Plain relaxed atomics: a host may SET from one thread
// while partitions probe on others; every reader takes one consistent value per call.
struct sq_options {
  std::atomic<int> probe_layout{0};         // 0 auto, 1 packed lines (when the index has them), 2 SoA arrays
  std::atomic<int> build_sort{0};           // sort keys of the build: 0 auto (32-bit when the key ranges fit), 1 always 64-bit
  std::atomic<int> build_ids{0};            // ids an index hands out: 0 build rows, 1 sorted positions (payload kept in that order)
  std::atomic<int> probe_tiles{1};          // tiles per CTA of an emitting packed-line launch: 1 / 2
  std::atomic<int> probe_block{128};        // rows per CTA of the packed-line kernels: 64 / 128 / 256
  std::atomic<int> lookback_backoff_ns{64}; // sleep between polls of a predecessor's chained-scan word
  std::atomic<int> rows_per_bin{0};         // build: target rows per directory bin; 0 = 8, or 1 for indexes of up to 4M rows (the
                                            // directory then costs 4 B per row of a cache-resident index and every in-bin search
                                            // is shorter: cfg3 1.32 -> 1.27 ms, profiles/r02_subcfg_rows_per_bin_sweep.txt)
  std::atomic<int> right_idx_wire{0};       // host entry points: 0 = per-row counts cross PCIe and right_idx is expanded on
                                            // the host (IJ:1611-1618), 1 = right_idx itself is copied
  std::atomic<int> staged_probe{0};         // 0 auto (adaptive per stream), 1 always try the staged kernel, 2 never
  std::atomic<int> scan_dict_capacity{1 << 16};  // text scan: initial capacity of the per-call key dictionary
  std::atomic<int> exec_trace{0};           // exec node: per-phase wall-clock trace on stderr
  std::atomic<int> pipeline_depth{3};       // sq_stream_submit: tiles in flight per stream (2..8)
  std::atomic<int> rank_count{1};           // rank-difference kernel for the indexes the SoA kernels serve: 0 off, 1 when the
                                            // index is deep enough to pay for it (measured crossover), 2 whenever possible
  std::atomic<int> coalesce_rows{1 << 19};  // exec node: probe rows that make one tile (sq_exec_probe_push / _pop)
};

struct sq_ctx {
  int device = 0;
  int sm_count = 148;
  size_t l2_persist_bytes = 0;  // L2 set-aside granted for persisting accesses
  size_t l2_window_max = 0;
  size_t l2_persist_max = 0;
  sq_options opt;
  sq::ErrorSlot err;
  std::mutex pool_mu;
  void* pool = nullptr;  // cudaMemPool_t of this context (sq_build.cu: ctx_pool); destroyed with the context
};

struct sq_column {
  void* d_values = nullptr;          // fixed-width values, or the bytes of a Utf8 column
  uint32_t width = 0;                // 4/8/16, or 0 for Utf8
  bool owned = false;
  int64_t* d_offsets = nullptr;      // Utf8: n_rows + 1 offsets into d_values
  uint64_t data_bytes = 0;
  uint8_t* d_validity = nullptr;     // optional Arrow validity bitmap (bit i = row i is valid)
};

struct sq_pack {  // up to four 4-byte build columns interleaved row-wise (sq_index_pack_columns)
  uint4* d_rows = nullptr;
  int n_cols = 0;
};

struct sq_index {
  sq_ctx* ctx = nullptr;
  uint64_t n_rows = 0;
  uint32_t n_keys = 0;
  // device arrays
  int32_t* d_start = nullptr;
  int32_t* d_runmax = nullptr;
  int32_t* d_end = nullptr;
  uint32_t* d_row = nullptr;
  bool narrow_sort = false;    // the build sorted 32-bit keys (key ranges laid end to end)
  bool pos_ids = false;        // option cuda_build_ids positions: d_row[j] = j, payload columns are stored in sorted order
  uint32_t* d_perm = nullptr;  // pos_ids: sorted position -> build row
  std::mutex perm_mu;          // pos_ids: payload columns are permuted one at a time on perm_stream
  struct sq_stream* perm_stream = nullptr;
  sq::SegMeta* d_meta = nullptr;
  uint32_t* d_dir = nullptr;
  uint64_t dir_bytes = 0;
  uint4* d_lines = nullptr;    // packed lines, or nullptr (wide / inverted intervals: SoA path only)
  uint2* d_dir_line = nullptr;
  int32_t* d_bmax = nullptr;   // block maxima of end: level 1, then level 2, then level 3 (one allocation)
  uint64_t bmax_n1 = 0, bmax_n2 = 0;
  int32_t* d_send = nullptr;   // ends sorted inside each key segment (rank-difference count), or nullptr
  sq::SegMeta* d_emeta = nullptr;
  uint32_t* d_edir = nullptr;
  uint64_t n_lines = 0;
  float mean_back_lines = 0.f; // mean number of extra lines a probe landing on a line's last row walks back
  float mean_depth = 0.f;      // sampled: rows before a row (same key) whose running max end reaches its start = overlap depth
  uint64_t* d_ht_keys = nullptr;
  uint32_t* d_ht_ids = nullptr;
  uint32_t ht_cap = 0;
  uint32_t sentinel_id = sq::kNoKey;
  uint64_t bytes = 0;
  float build_ms = 0.f;
  std::mutex col_mu;
  std::vector<sq_column> columns;
  std::vector<sq_pack> packs;

  sq::IndexView view() const {
    sq::IndexView v;
    v.start = d_start; v.runmax = d_runmax; v.end = d_end; v.row = d_row; v.meta = d_meta; v.dir = d_dir;
    v.ht_keys = d_ht_keys; v.ht_ids = d_ht_ids; v.ht_mask = ht_cap - 1; v.sentinel_id = sentinel_id;
    v.n_keys = n_keys; v.n_rows = uint32_t(n_rows);
    v.lines = d_lines;
    v.dir_line = d_dir_line;
    v.bmax1 = d_bmax;
    v.bmax2 = d_bmax ? d_bmax + bmax_n1 : nullptr;
    v.bmax3 = d_bmax ? d_bmax + bmax_n1 + bmax_n2 : nullptr;
    v.send = d_send;
    v.emeta = d_emeta;
    v.edir = d_edir;
    return v;
  }
};

// Grow-only buffer (device or pinned host).
struct sq_buf {
  void* p = nullptr;
  size_t cap = 0;
  bool pinned = false;
};

struct sq_tile_slot;  // sq_pipeline.cu

struct sq_stream {
  sq_ctx* ctx = nullptr;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  sq::ErrorSlot err;

  // asynchronous tile pipeline (sq_stream_submit / sq_stream_collect, sq_pipeline.cu)
  std::vector<sq_tile_slot*> slots;
  cudaStream_t stream_in = nullptr, stream_out = nullptr;  // H2D and D2H run beside the kernels of other tiles
  uint64_t next_ticket = 1, oldest_ticket = 1;
  double pairs_per_row = -1.0;         // of the last collected tile: sizes the next tiles' buffers and speculative copies
  double pipe_ms[3] = {0, 0, 0};       // summed device time of h2d / kernels / d2h over collected tiles
  uint64_t pipe_bytes[2] = {0, 0};     // bytes moved h2d / d2h
  uint64_t pipe_tiles = 0, pipe_regrow = 0;
  sq_buf d_dict;                       // key dictionary of sq_stream_submit_ids: u64 hash per dictionary entry
  uint32_t dict_n = 0;

  // which packed-line kernel serves the next tile (sq_api.cu: pick_staged / staged_feedback)
  uint32_t staged_skip = 0;      // tiles to send to k_probe_packed before the staged kernel is tried again
  uint32_t staged_backoff = 0;   // grows while tiles keep failing to stage
  uint64_t staged_tiles = 0, staged_global_ctas = 0;
  bool order_local = false;      // verdict of the last look at a device tile's row order
  uint32_t order_age = 0;        // device tiles until the next look

  // state of the tile currently between count and emit
  const sq_index* idx = nullptr;
  const sq_index* l2_window_idx = nullptr;  // index whose directory this stream's L2 window covers
  uint32_t n_rows = 0;
  uint64_t n_pairs = 0;
  bool counted = false;
  bool emitted = false;
  const uint64_t* d_q_key = nullptr;   // probe columns of the tile (device), kept for the emit pass
  const int32_t* d_q_start = nullptr;
  const int32_t* d_q_end = nullptr;
  bool spec_valid = false;             // the count pass already wrote the pairs (speculative emit)
  uint32_t* d_spec_left = nullptr;
  uint32_t* d_spec_right = nullptr;
  const uint32_t* d_last_left = nullptr;
  const uint32_t* d_last_right = nullptr;
  bool win_set = false;                // sq_stream_set_window: gathers / fetches see pairs [win_off, win_off + win_n)
  uint64_t win_off = 0, win_n = 0;

  // device scratch
  sq_buf d_in;       // staged probe key/start/end (host entry points)
  sq_buf d_cnt;      // hit count per probe row (rle_right)
  sq_buf d_cnt8;     // the same as bytes (SQ_TILE_COUNTS_U8)
  sq_buf d_state;    // per probe row: lo, nc, hit mask (3 x u32 arrays)
  sq_buf d_tile;     // per-CTA pair totals -> offsets, + chained-scan words
  sq_buf d_scalar;   // n_pairs, ticket counter, cast-error slot, digest
  sq_buf d_left, d_right;  // emitted pairs (host entry points)
  sq_buf d_gather;   // gather staging
  sq_buf d_gather2;  // Utf8 gather: staged probe column / output offsets
  sq_buf d_chain;    // chained-scan words of launch_scan_u64
  sq_buf d_strblk;   // Utf8 gather: per-block length sums
  sq_buf d_strdata;  // Utf8 gather: output bytes
  // Utf8 gather in flight between sq_gather_utf8 (offsets) and sq_gather_utf8_data (bytes)
  const int64_t* str_src_off = nullptr;
  const uint8_t* str_src_data = nullptr;
  const uint32_t* str_idx = nullptr;
  int64_t* str_out_off = nullptr;
  uint64_t str_total = 0;
  bool str_pending = false;
  // pinned staging
  sq_buf h_in, h_out, h_scalar;
  sq_buf h_scan;     // ring of pinned chunks the text scanner copies pageable host text through

  // profiling
  bool profiling = false;
  cudaEvent_t ev[8] = {};
  bool ev_ready = false;
  cudaEvent_t ev_rle = nullptr;          // "the counts have arrived" (right_idx is decoded from them on the host)
  float phase_ms[5] = {0, 0, 0, 0, 0};   // last value per phase
  double phase_sum[5] = {0, 0, 0, 0, 0}; // running sums since profiling was enabled
  uint64_t phase_n[5] = {0, 0, 0, 0, 0};
  uint32_t pending = 0;                  // phases with a recorded, unread event pair
  uint64_t launches = 0;
};

namespace sq {

// error helpers ---------------------------------------------------------------------------
int fail(ErrorSlot& e, int code, const char* fmt, ...);
#define SQ_CUDA(slot, expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return sq::fail((slot), _e == cudaErrorMemoryAllocation ? SQ_ENOMEM : SQ_ECUDA,        \
                      "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

int ensure(ErrorSlot& e, sq_buf& b, size_t bytes, bool pinned);
void release(sq_buf& b);

// build.cu
int build_index_device(sq_ctx* ctx, const uint64_t* d_key, const int32_t* d_start, const int32_t* d_end,
                       uint64_t n, cudaStream_t st, sq_index** out);
void free_index(sq_index* idx);

// probe.cu
// K1 + K2: per-row state (lo, nc, hit mask, count), tile offsets, result[0] = n_pairs
int launch_count(sq_stream* s, const sq_index* idx, const uint64_t* d_key, const int32_t* d_start,
                 const int32_t* d_end, uint32_t n);
// K3: pairs from the saved state; sets result[1] instead of writing when n_pairs > capacity
int launch_write(sq_stream* s, const sq_index* idx, const int32_t* d_start, uint32_t n, uint32_t* d_left,
                 uint32_t* d_right, uint64_t capacity);

// probe_rank.cu: indexes with the rank structure over the ends (idx->d_send): ONE kernel — search, rank-difference count,
// chained scan, candidate walk + write (count only when d_left == nullptr: no candidate is touched at all).
// Same outputs as launch_count + launch_write: cnt per row in s->d_cnt, result[0] = n_pairs, result[1] = overflow.
int launch_rank_join(sq_stream* s, const sq_index* idx, const uint64_t* d_key, const int32_t* d_start,
                     const int32_t* d_end, uint32_t n, uint32_t* d_left, uint32_t* d_right, uint64_t capacity);

// one output row per probe row: an overlapping build row, else the nearest one, else kEmptyRow (NULL)
int launch_nearest(sq_stream* s, const sq_index* idx, const uint64_t* d_key, const int32_t* d_start,
                   const int32_t* d_end, uint32_t n, uint32_t* d_left);
// right_idx of a one-row-per-probe-row result: 0, 1, ..., n-1
int launch_iota(sq_stream* s, uint32_t* d_out, uint64_t n);
int launch_narrow_counts(sq_stream* s, cudaStream_t st, const uint32_t* d_cnt, uint32_t n, uint8_t* d_out, unsigned long long* d_too_big);
int launch_narrow_offsets(sq_stream* s, const int64_t* d_in, uint64_t n, int32_t* d_out);

// probe_packed.cu: one fused pass over the packed lines (count + chained scan + write when
// d_left != nullptr; count only otherwise).  Leaves cnt per row in s->d_cnt, result[0] = n_pairs and
// result[1] = 1 when the pairs did not fit `capacity` (nothing useful was written then).
int launch_packed(sq_stream* s, const sq_index* idx, const uint64_t* d_key, const int32_t* d_start,
                  const int32_t* d_end, uint32_t n, uint32_t* d_left, uint32_t* d_right, uint64_t capacity);
bool use_packed(const sq_index* idx);
bool use_rank(const sq_index* idx);  // probe_rank.cu: the rank-difference kernel serves this (SoA-served) index
// probe_staged.cu: the same contract as launch_packed for position-local tiles (build tile staged in shared memory by
// TMA, one thread per probe row); result[2] = CTAs that could not stage and walked global memory instead
int launch_staged(sq_stream* s, const sq_index* idx, const uint64_t* d_key, const int32_t* d_start,
                  const int32_t* d_end, uint32_t n, uint32_t* d_left, uint32_t* d_right, uint64_t capacity);
uint32_t staged_tile_rows();
// api.cu: adaptive choice between the two (option cuda_staged_probe; `policy` = the stream that keeps the statistics,
// `host_key` / `host_start` = the tile's host columns when the caller has them: a cheap look at their order)
bool pick_staged(sq_stream* policy, const sq_index* idx, const uint64_t* host_key, const int32_t* host_start,
                 const uint64_t* d_key, const int32_t* d_start, uint32_t n, bool emits, const uint32_t* host_ids = nullptr);
// probe_staged.cu: samples up to 1024 adjacent row pairs of a device tile: {pairs with equal key and non-decreasing
// start, pairs looked at} (synchronises the stream: device entry points only)
int sample_order_device(sq_stream* s, const uint64_t* d_key, const int32_t* d_start, uint32_t n, uint32_t out2[2]);
void staged_feedback(sq_stream* policy, uint32_t n_rows, uint64_t global_ctas, uint64_t staged_lines);
int launch_packed_any(sq_stream* s, sq_stream* policy, bool staged, const sq_index* idx, const uint64_t* d_key,
                      const int32_t* d_start, const int32_t* d_end, uint32_t n, uint32_t* d_left, uint32_t* d_right,
                      uint64_t capacity);


// api.cu: per-tile state of a stream (count -> emit protocol) and the write pass of a counted tile
void tile_begin(sq_stream* s, const sq_index* idx, const uint64_t* dk, const int32_t* ds, const int32_t* de, uint32_t n);
int tile_emit(sq_stream* s, uint32_t* d_left, uint32_t* d_right, uint64_t capacity);
// gather.cu: key ids -> key hashes through a stream's dictionary; rows with an id outside it get an interval nothing overlaps
int launch_expand_ids(sq_stream* s, cudaStream_t st, const uint32_t* d_ids, const uint64_t* d_dict, uint32_t dict_n, uint32_t n,
                      uint64_t* d_key_out, int32_t* d_start, int32_t* d_end);
// pipeline.cu
void pipeline_destroy(sq_stream* s);
uint64_t pipeline_bytes(const sq_stream* s);

// rle.cpp (plain C++): right_idx from the per-row counts, on the host (interval_join.rs:1611-1618)
void expand_counts(const uint32_t* counts, uint32_t n_rows, uint32_t* right, uint64_t n_pairs);

// gather.cu
int launch_gather(sq_stream* s, const void* d_values, const uint32_t* d_idx, uint64_t n, uint32_t width,
                  void* d_out);
// row-wise pack of up to four 4-byte columns, and the multi-column gathers over it / over probe columns
int launch_pack_columns(sq_ctx* ctx, const uint32_t* const* d_cols, int n_cols, uint64_t n_rows, uint4* d_rows);
int launch_gather_pack(sq_stream* s, const uint4* d_rows, const uint32_t* d_idx, uint64_t n, uint32_t* const* d_outs, int n_outs);
int launch_gather_probe_columns(sq_stream* s, const uint32_t* const* d_cols, const uint32_t* d_idx, uint64_t n,
                                uint32_t* const* d_outs, int n_cols);
int launch_cast_i64(sq_stream* s, const int64_t* d_in, uint64_t n, int64_t minus, int32_t* d_out,
                    int64_t* bad_value, bool* bad);
int launch_digest(sq_stream* s, const uint32_t* d_left, const uint32_t* d_right, uint64_t n,
                  uint64_t right_offset, uint64_t out3[3]);
// Utf8 take: d_out_off[n+1] = exclusive scan of the gathered string lengths; *total = bytes (stream synced)
int launch_str_offsets(sq_stream* s, const int64_t* d_src_off, const uint32_t* d_idx, uint64_t n,
                       int64_t* d_out_off, uint64_t* total);
int launch_str_copy(sq_stream* s, const int64_t* d_src_off, const uint8_t* d_src_data, const uint32_t* d_idx,
                    uint64_t n, const int64_t* d_out_off, uint8_t* d_out_data);
int launch_gather_bits(sq_stream* s, const uint8_t* d_bitmap, const uint32_t* d_idx, uint64_t n, uint8_t* d_out,
                       uint64_t* null_count);
// exclusive scan, in place, of n u64 values (chained scan with decoupled look-back); total -> d_total[0]
int launch_scan_u64(sq_stream* s, unsigned long long* d_vals, uint32_t n, unsigned long long* d_total);

}  // namespace sq
