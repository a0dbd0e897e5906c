/* sequila_driver.h — the partition loop of a host, natively.
 *
 * In the reference DataFusion spawns one task per output partition; each polls its IntervalJoinStream, which
 * fetches a probe batch, joins it and yields one output batch (interval_join.rs:449-557, 1054-1167, 1192-1233).
 * sq_driver_run is that loop over the C ABI of sequila_cuda.h for a probe side that already sits in (pinned)
 * host memory: `n_partitions` OS threads, one sq_stream each, the probe rows cut into `n_tiles` tiles dealt
 * round-robin (tile t -> partition t % n_partitions, as DataFusion's round-robin repartitioning would), every
 * partition keeping `cuda_pipeline_depth` tiles in flight through sq_stream_submit / sq_stream_collect and handing
 * each collected tile to `consume` (NULL = drop it) before returning the tile's pinned buffers to the pool.
 * It is the harness bench.py times end to end and the shape a Rust exec node's `execute()` would take; it adds
 * nothing to the data path.
 */
#ifndef SEQUILA_DRIVER_H
#define SEQUILA_DRIVER_H

#include "sequila_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* called on the partition's thread with the tile's first probe row and its result; the buffers belong to the
 * library again when the callback returns */
typedef void (*sq_tile_consumer)(void* user, int32_t partition, uint64_t first_row, const sq_tile_out* tile);

typedef struct sq_drive_stats {
  uint64_t n_pairs;        /* sum over tiles */
  uint64_t n_tiles;
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t left_xor;       /* xor over all left_idx values: a cheap witness that every byte arrived */
  uint64_t regrown_tiles;  /* tiles re-emitted into larger buffers */
  double seconds;          /* wall time of the whole pass, all partitions (steady clock) */
  double h2d_ms, kernel_ms, d2h_ms; /* device time summed over tiles */
} sq_drive_stats;

/* A driver owns `n_partitions` sq_streams (their device scratch, pinned staging and CUDA streams live as long as it
 * does, like the IntervalJoinStreams of a running query).  sq_driver_run makes one pass over a probe side:
 * flags = SQ_TILE_* of sq_stream_submit; `checksum` != 0 makes every partition xor the left_idx of its tiles (reads all
 * output bytes once on the host: a stand-in for a consumer that touches the result).  One run at a time per driver. */
typedef struct sq_driver sq_driver;
int32_t sq_driver_create(sq_ctx* ctx, int32_t n_partitions, sq_driver** out);
int32_t sq_driver_run(sq_driver* d, const sq_index* idx, const uint64_t* key_hash, const int32_t* start,
                      const int32_t* end, uint64_t n_rows, int32_t n_tiles, uint32_t flags, int32_t checksum,
                      sq_tile_consumer consume, void* user, sq_drive_stats* stats_out);
/* the same pass with the key column as 4-byte dictionary ids (sq_stream_submit_ids; 12 bytes per probe row on the wire);
 * the dictionary is uploaded to every partition's stream first */
int32_t sq_driver_run_ids(sq_driver* d, const sq_index* idx, const uint64_t* dict_key_hashes, uint32_t dict_entries,
                          const uint32_t* key_id, const int32_t* start, const int32_t* end, uint64_t n_rows, int32_t n_tiles,
                          uint32_t flags, int32_t checksum, sq_tile_consumer consume, void* user, sq_drive_stats* stats_out);
const char* sq_driver_last_error(const sq_driver* d);
void sq_driver_free(sq_driver* d);

#ifdef __cplusplus
}
#endif
#endif /* SEQUILA_DRIVER_H */
