// sq_driver_* (include/sequila_driver.h): the partition loop of a host over the public C ABI — what
// IntervalJoinExec::execute + DataFusion's per-partition tasks do around the stream state machine
// (interval_join.rs:449-557, 1054-1167).  Uses nothing but sequila_cuda.h.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <thread>
#include <vector>

#include "sequila_driver.h"

#define SQ_API extern "C" __attribute__((visibility("default")))

struct sq_driver {
  sq_ctx* ctx = nullptr;
  std::vector<sq_stream*> streams;
  std::string err;
};

SQ_API int32_t sq_driver_create(sq_ctx* ctx, int32_t n_partitions, sq_driver** out) {
  if (!ctx || !out || n_partitions < 1) return SQ_EINVAL;
  *out = nullptr;
  auto* d = new sq_driver();
  d->ctx = ctx;
  for (int p = 0; p < n_partitions; ++p) {
    sq_stream* s = nullptr;
    const int rc = sq_stream_create(ctx, &s);
    if (rc != SQ_OK) {
      for (sq_stream* q : d->streams) sq_stream_free(q);
      delete d;
      return rc;
    }
    d->streams.push_back(s);
  }
  *out = d;
  return SQ_OK;
}

SQ_API void sq_driver_free(sq_driver* d) {
  if (!d) return;
  for (sq_stream* s : d->streams) sq_stream_free(s);
  delete d;
}

SQ_API const char* sq_driver_last_error(const sq_driver* d) { return d ? d->err.c_str() : ""; }

namespace {
struct Part {
  uint64_t pairs = 0, tiles = 0, x = 0;
  int rc = SQ_OK;
  std::string err;
};
}  // namespace

static int32_t run_pass(sq_driver* d, const sq_index* idx, const uint64_t* key_hash, const uint32_t* key_id, const int32_t* start,
                        const int32_t* end, uint64_t n_rows, int32_t n_tiles, uint32_t flags, int32_t checksum,
                        sq_tile_consumer consume, void* user, sq_drive_stats* out) {
  if (!d || !idx || !out || n_tiles < 1) return SQ_EINVAL;
  memset(out, 0, sizeof *out);
  sq_ctx* ctx = d->ctx;
  const int T = int(d->streams.size());
  std::vector<Part> parts(static_cast<size_t>(T));
  std::vector<double> before(size_t(T) * 8, 0.0);
  for (int w = 0; w < T; ++w) sq_stream_pipeline_stats(d->streams[size_t(w)], &before[size_t(w) * 8]);
  char dv[16] = "3";
  sq_ctx_get_option(ctx, "cuda_pipeline_depth", dv, sizeof dv);
  const size_t depth = size_t(atoi(dv) > 0 ? atoi(dv) : 3);
  auto bound = [&](int64_t t) { return uint64_t((unsigned __int128)n_rows * uint64_t(t) / uint64_t(n_tiles)); };

  auto work = [&](int w) {
    Part& p = parts[size_t(w)];
    sq_stream* st = d->streams[size_t(w)];
    std::deque<std::pair<uint64_t, int>> pend;  // (ticket, tile)
    auto collect_one = [&]() -> bool {
      const uint64_t ticket = pend.front().first;
      const int t = pend.front().second;
      pend.pop_front();
      sq_tile_out res;
      const int rc = sq_stream_collect(st, ticket, &res);
      if (rc != SQ_OK) { p.rc = rc; p.err = sq_stream_last_error(st); return false; }
      p.pairs += res.n_pairs;
      p.tiles += 1;
      if (checksum && res.left_idx) {
        uint64_t x = 0;
        const uint32_t* l = res.left_idx;
        for (uint64_t k = 0; k < res.n_pairs; ++k) x ^= l[k];
        p.x ^= x;
      }
      if (consume) consume(user, w, bound(t), &res);
      sq_host_free(ctx, res.left_idx);
      sq_host_free(ctx, res.right_idx);
      sq_host_free(ctx, res.counts);
      return true;
    };
    auto drain = [&]() {  // after a failure: the stream's slots must be empty again for the next run
      while (!pend.empty()) {
        sq_tile_out res;
        if (sq_stream_collect(st, pend.front().first, &res) == SQ_OK) {
          sq_host_free(ctx, res.left_idx);
          sq_host_free(ctx, res.right_idx);
          sq_host_free(ctx, res.counts);
        }
        pend.pop_front();
      }
    };
    for (int t = w; t < n_tiles; t += T) {
      if (pend.size() == depth && !collect_one()) { drain(); return; }
      const uint64_t lo = bound(t), hi = bound(t + 1);
      uint64_t ticket = 0;
      const int rc = key_id ? sq_stream_submit_ids(st, idx, key_id + lo, start + lo, end + lo, uint32_t(hi - lo), flags, &ticket)
                            : sq_stream_submit(st, idx, key_hash + lo, start + lo, end + lo, uint32_t(hi - lo), flags, &ticket);
      if (rc != SQ_OK) { p.rc = rc; p.err = sq_stream_last_error(st); drain(); return; }
      pend.emplace_back(ticket, t);
    }
    while (!pend.empty())
      if (!collect_one()) { drain(); return; }
  };

  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int w = 1; w < T; ++w) pool.emplace_back(work, w);
  work(0);
  for (auto& th : pool) th.join();
  out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  int rc = SQ_OK;
  for (int w = 0; w < T; ++w) {
    const Part& p = parts[size_t(w)];
    out->n_pairs += p.pairs;
    out->n_tiles += p.tiles;
    out->left_xor ^= p.x;
    double s8[8] = {0};
    sq_stream_pipeline_stats(d->streams[size_t(w)], s8);
    const double* b8 = &before[size_t(w) * 8];
    out->h2d_ms += s8[0] - b8[0];
    out->kernel_ms += s8[1] - b8[1];
    out->d2h_ms += s8[2] - b8[2];
    out->h2d_bytes += uint64_t(s8[3] - b8[3]);
    out->d2h_bytes += uint64_t(s8[4] - b8[4]);
    out->regrown_tiles += uint64_t(s8[6] - b8[6]);
    if (p.rc != SQ_OK && rc == SQ_OK) { rc = p.rc; d->err = p.err; }
  }
  return rc;
}

SQ_API int32_t sq_driver_run(sq_driver* d, const sq_index* idx, const uint64_t* key_hash, const int32_t* start,
                             const int32_t* end, uint64_t n_rows, int32_t n_tiles, uint32_t flags, int32_t checksum,
                             sq_tile_consumer consume, void* user, sq_drive_stats* out) {
  return run_pass(d, idx, key_hash, nullptr, start, end, n_rows, n_tiles, flags, checksum, consume, user, out);
}

SQ_API int32_t sq_driver_run_ids(sq_driver* d, const sq_index* idx, const uint64_t* dict_key_hashes, uint32_t dict_entries,
                                 const uint32_t* key_id, const int32_t* start, const int32_t* end, uint64_t n_rows,
                                 int32_t n_tiles, uint32_t flags, int32_t checksum, sq_tile_consumer consume, void* user,
                                 sq_drive_stats* out) {
  if (!d || !key_id) return SQ_EINVAL;
  for (sq_stream* st : d->streams) {
    const int rc = sq_stream_set_key_dictionary(st, dict_key_hashes, dict_entries);
    if (rc != SQ_OK) { d->err = sq_stream_last_error(st); return rc; }
  }
  return run_pass(d, idx, nullptr, key_id, start, end, n_rows, n_tiles, flags, checksum, consume, user, out);
}
