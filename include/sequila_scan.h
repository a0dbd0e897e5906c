/* sequila_scan.h — scan side of the interval join: delimited text (BED / CSV) -> device columns.
 *
 * SURVEY.md §8(f) rank 4 (the step in front of the hot path).  In the reference the two join inputs
 * are `CREATE EXTERNAL TABLE .. (contig VARCHAR NOT NULL, start BIGINT NOT NULL, end BIGINT NOT NULL)
 * STORED AS CSV LOCATION '...bed' OPTIONS ('delimiter' '\t', 'has_header' 'false')`
 * (queries/q1-coitrees.sql:6-14; testing/data/interval/reads.csv, targets.csv with a header and ','):
 * DataFusion's CSV reader produces Arrow batches on the host, `create_hashes` hashes the contig
 * strings row by row (interval_join.rs "IJ":1037, IJ:1211) and `evaluate_as_i32` casts the BIGINT
 * columns (IJ:1661-1672).  Here the file bytes are parsed ON THE DEVICE straight into the three
 * columns the join consumes — key hash (u64), start, end (i32, after the checked cast) — plus a
 * dictionary encoding of the key column (u32 id per row, ids in order of first occurrence, the
 * distinct strings kept on the host), so no host-side Arrow round trip and no host string hashing
 * is left: sq_index_build_device / sq_probe_*_device take the columns as they are, and
 * sq_index_add_column_device takes the ids as a 4-byte payload column.
 *
 * Text format accepted (what the reference's tables declare, nothing more):
 *   - rows end with '\n' or "\r\n"; the last row may lack the terminator; empty rows are skipped;
 *     rows starting with `comment` (when non-zero) are skipped; with `has_header` the first
 *     remaining row is skipped;
 *   - fields are separated by `delimiter`; no quoting (a '"' opening a used field is an error),
 *     no trimming; fields beyond the ones named are ignored (BED4+);
 *   - integer fields: [+-]?[0-9]+ within Int64; `*_minus` is subtracted (the `end - 1` of strict
 *     comparisons, intervals.rs:67-69, or BED's half-open end), then the value must fit Int32:
 *     otherwise SQ_ECAST with the reference's text "Arrow error: Cast error: Can't cast value {v}
 *     to type Int32" (IJ:1959-1965) for the first offending row (start before end).
 *   - anything else: SQ_EPARSE naming the byte offset and the row.
 * key hash = sq_keyhash.h over the key field's bytes == what the exec node computes for a Utf8 `on`
 * column, so a scanned side and an Arrow side of one join agree.
 */
#ifndef SEQUILA_SCAN_H
#define SEQUILA_SCAN_H

#include "sequila_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sq_scan sq_scan; /* device-resident columns of one scanned table */

typedef struct sq_scan_options {
  uint8_t delimiter;  /* OPTIONS ('delimiter' ..): ',' or '\t'                                   */
  uint8_t has_header; /* OPTIONS ('has_header' ..)                                               */
  uint8_t comment;    /* 0 = none                                                                */
  uint8_t reserved;
  int32_t col_key;    /* 0-based field of the `on` column (contig); -1 = range-only join: every
                         row carries the key of on=[(1,1)] (sequila_physical_planner.rs:136)      */
  int32_t col_start;  /* 0-based field of the interval start                                     */
  int32_t col_end;    /* 0-based field of the interval end                                       */
  int32_t reserved2;
  int64_t start_minus; /* subtracted before the Int32 cast                                       */
  int64_t end_minus;
} sq_scan_options;

/* Host text (borrowed for the call; copied to the device inside).  The scan runs on the stream's
 * CUDA stream and is complete when the call returns. */
int32_t sq_scan_text(sq_stream* s, const uint8_t* text, uint64_t n_bytes, const sq_scan_options* opt,
                     sq_scan** out);
/* Text already in device memory (16-byte aligned; borrowed for the call). */
int32_t sq_scan_text_device(sq_stream* s, const uint8_t* d_text, uint64_t n_bytes,
                            const sq_scan_options* opt, sq_scan** out);

uint64_t sq_scan_rows(const sq_scan* sc);
uint64_t sq_scan_bytes(const sq_scan* sc); /* device bytes held */
/* device columns, n_rows long, valid until sq_scan_free */
const uint64_t* sq_scan_key_hash_device(const sq_scan* sc);
const int32_t* sq_scan_start_device(const sq_scan* sc);
const int32_t* sq_scan_end_device(const sq_scan* sc);
const uint32_t* sq_scan_key_ids_device(const sq_scan* sc); /* NULL when col_key < 0 */
/* dictionary of the key column: id -> bytes (not NUL-terminated) and the id's key hash */
uint32_t sq_scan_dict_size(const sq_scan* sc);
int32_t sq_scan_dict_entry(const sq_scan* sc, uint32_t id, const uint8_t** bytes, uint32_t* len,
                           uint64_t* key_hash);
/* copy columns to host buffers (any pointer may be NULL) */
int32_t sq_scan_fetch(sq_stream* s, const sq_scan* sc, uint64_t* key_hash, int32_t* start, int32_t* end,
                      uint32_t* key_ids);
/* device time of the last scan in ms, CUDA events on the stream:
 * [0] = host-to-device copy of the text, [1] = row location + parse kernels, [2] = id assignment */
int32_t sq_scan_timing(const sq_scan* sc, float out3[3]);
void sq_scan_free(sq_scan* sc);

#ifdef __cplusplus
}
#endif
#endif /* SEQUILA_SCAN_H */
