/* sequila_driver.h — the partition loop of a host, natively.
 *
 * In the reference DataFusion spawns one task per output partition; each polls its IntervalJoinStream, which
 * fetches a probe batch, joins it and yields one output batch (interval_join.rs:449-557, 1054-1167, 1192-1233).
 * sq_drive_partitions is that loop over the C ABI of sequila_cuda.h for a probe side that already sits in (pinned)
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

/* flags = SQ_TILE_* of sq_stream_submit.  `checksum` != 0 makes every partition xor the left_idx of its tiles
 * (reads all output bytes once on the host: a stand-in for a consumer that touches the result). */
int32_t sq_drive_partitions(sq_ctx* ctx, const sq_index* idx, const uint64_t* key_hash, const int32_t* start,
                            const int32_t* end, uint64_t n_rows, int32_t n_partitions, int32_t n_tiles,
                            uint32_t flags, int32_t checksum, sq_tile_consumer consume, void* user,
                            sq_drive_stats* stats_out);

#ifdef __cplusplus
}
#endif
#endif /* SEQUILA_DRIVER_H */
