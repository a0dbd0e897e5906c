/* sequila_exec.h — host-side exec node of the `Cuda` interval join over the Arrow C Data Interface.
 *
 * C++ mirror (the Rust toolchain is absent from this image) of what the reference's
 * `IntervalJoinExec` / `IntervalJoinStream` do around the index
 * (sequila/sequila-core/src/physical_planner/joins/interval_join.rs, "IJ"):
 *
 *   sq_exec_create        IntervalJoinExec::try_new  IJ:112-172  (on, intervals, projection, schema)
 *   sq_exec_push_build    collect_left_input fold    IJ:624-640  (batches + metrics)
 *   sq_exec_finish_build  update_hashmap + ::new + concat_batches  IJ:662-685  -> sq_index_build
 *   sq_exec_probe         fetch_probe_batch + process_probe_batch (full mode)  IJ:1192-1233, 1580-1640:
 *                         key hashes, evaluate_as_i32, probe on the GPU, `take` of every projected
 *                         column on the GPU, one output RecordBatch per probe batch
 *   sq_exec_probe_begin / _next   the same in low-memory mode (IJ:1433-1530): bounded output batches
 *   sq_exec_metrics       BuildProbeJoinMetrics      utils.rs:441-495
 *
 * Batches cross the boundary as Arrow C Data Interface struct arrays (a RecordBatch).  Key hashing
 * is this layer's own 64-bit hash: the join only needs it to be injective on the keys present
 * (IJ:1042-1048 groups by the hash alone).  Everything compute-heavy goes through
 * include/sequila_cuda.h; there is no CPU fallback.
 */
#ifndef SEQUILA_EXEC_H
#define SEQUILA_EXEC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
struct ArrowSchema {
  const char* format;
  const char* name;
  const char* metadata;
  int64_t flags;
  int64_t n_children;
  struct ArrowSchema** children;
  struct ArrowSchema* dictionary;
  void (*release)(struct ArrowSchema*);
  void* private_data;
};
struct ArrowArray {
  int64_t length;
  int64_t null_count;
  int64_t offset;
  int64_t n_buffers;
  int64_t n_children;
  const void** buffers;
  struct ArrowArray** children;
  struct ArrowArray* dictionary;
  void (*release)(struct ArrowArray*);
  void* private_data;
};
#endif

typedef struct sq_exec sq_exec;

typedef struct sq_exec_config {
  int32_t device;             /* CUDA ordinal */
  int32_t n_on;               /* equi-key column pairs; 0 = range-only join, on=[(lit(1), lit(1))] (PP:127-148) */
  const int32_t* on_left;     /* build-side column indices */
  const int32_t* on_right;    /* probe-side column indices */
  int32_t left_start, left_end;   /* build interval columns (ColIntervals.left_interval, IV:24-28) */
  int32_t right_start, right_end; /* probe interval columns */
  int32_t left_end_minus_one;     /* 1 when the parser wrapped the end in `- 1` (strict operator, IV:67-69) */
  int32_t right_end_minus_one;
  int32_t n_projection;       /* -1 = all columns: left then right (build_join_schema, IJ:133) */
  const int32_t* projection;  /* indices into [left columns..., right columns...] (IJ:520-526) */
  int32_t algorithm;          /* SQ_EXEC_OVERLAPS: every overlapping pair (alg=Cuda, the Coitrees semantics);
                                 SQ_EXEC_NEAREST: one row per probe row, left side = an overlapping build row,
                                 else the nearest one, else NULL (alg=CudaNearest = CoitreesNearest, IJ:972-990) */
  int32_t low_memory;         /* sequila.interval_join_low_memory (SC:54): informational here, the caller picks
                                 sq_exec_probe (one output batch per probe batch) or the begin/next protocol */
  int64_t max_output_rows;    /* cap of one output batch in the begin/next protocol; <= 0: 1,000,000 (IJ:1439) */
} sq_exec_config;
#define SQ_EXEC_OVERLAPS 0
#define SQ_EXEC_NEAREST 1

/* schemas are struct ("+s") schemas of the two inputs; they are only read during the call */
int32_t sq_exec_create(const sq_exec_config* cfg, const struct ArrowSchema* left_schema,
                       const struct ArrowSchema* right_schema, sq_exec** out);
/* takes ownership of *batch (moves it; batch->release is set to NULL) */
int32_t sq_exec_push_build(sq_exec* e, struct ArrowArray* batch);
int32_t sq_exec_finish_build(sq_exec* e);
/* fills *out with the projected join schema (caller releases it) */
int32_t sq_exec_output_schema(const sq_exec* e, struct ArrowSchema* out);
/* probes one batch (borrowed for the duration of the call) for `partition`; fills *out with one
 * struct array = the output RecordBatch of this probe batch (caller releases it) */
int32_t sq_exec_probe(sq_exec* e, int32_t partition, const struct ArrowArray* batch, struct ArrowArray* out);
/* Low-memory mode (IJ:1433-1530): the probe batch is joined once on the GPU, then handed back in output
 * batches of at most max_output_rows rows cut at probe-row boundaries (a single probe row with more matches
 * forms a batch of its own).  `batch` must stay valid until sq_exec_probe_next reported *has_more_out = 0. */
int32_t sq_exec_probe_begin(sq_exec* e, int32_t partition, const struct ArrowArray* batch);
int32_t sq_exec_probe_next(sq_exec* e, int32_t partition, struct ArrowArray* out, int32_t* has_more_out);
/* Probe-batch coalescing.  The reference joins every probe batch on its own (<= 8192 rows by default, IJ:1192-1233); a GPU
 * launch chain and a PCIe round trip per 8192 rows are bound by their latencies, so batches are pushed (ownership moves,
 * like sq_exec_push_build) and leave as ONE tile once `cuda_coalesce_rows` rows (sq_exec_set_option, default 524288) are
 * waiting: *ready_out = 1 then.  sq_exec_probe_pop takes the waiting batches as a tile if they reached the target (or if
 * `flush` != 0: end of the probe stream).  Tiles are pipelined two deep per partition: the call concatenates, hashes and
 * casts the new tile on the calling thread while a worker thread still runs the previous tile on the GPU, then fills *out
 * with the PREVIOUS tile's output batch (ONE batch per tile, rows in probe order, *has_out = 1) and starts the worker of the
 * new one; *has_out = 0 when no finished tile was due.  At the end of the probe stream call it with flush != 0 until
 * *has_out = 0. */
int32_t sq_exec_probe_push(sq_exec* e, int32_t partition, struct ArrowArray* batch, int32_t* ready_out);
int32_t sq_exec_probe_pop(sq_exec* e, int32_t partition, int32_t flush, struct ArrowArray* out, int32_t* has_out);
/* [0] build_input_batches [1] build_input_rows [2] build_mem_used [3] input_batches [4] input_rows
 * [5] output_batches [6] output_rows [7] build_time_ns [8] join_time_ns [9] index_bytes [10] keys
 * [11] coalesced probe tiles joined */
int32_t sq_exec_metrics(const sq_exec* e, uint64_t out[16]);
const char* sq_exec_last_error(const sq_exec* e);
/* `SET sequila.cuda_<name> TO <value>` for this node's context (keys of sq_ctx_set_option, sequila_cuda.h);
 * the reference carries its knobs the same way, as SequilaConfig fields set through SET (SC:50-60, 106-132) */
int32_t sq_exec_set_option(sq_exec* e, const char* key, const char* value);
void sq_exec_free(sq_exec* e);

#ifdef __cplusplus
}
#endif
#endif /* SEQUILA_EXEC_H */
