// Asynchronous tile pipeline of a probe stream: sq_stream_submit / sq_stream_collect (include/sequila_cuda.h).
//
// The reference's IntervalJoinStream handles one probe batch at a time, synchronously, inside poll_next
// (interval_join.rs:1146-1151, 1192-1233, 1580-1640).  A GPU behind PCIe needs the three legs of a tile — H2D of the
// probe columns, the probe kernels, D2H of the pairs — to overlap ACROSS tiles, and the host must never wait in the
// middle of a tile.  So a stream keeps `cuda_pipeline_depth` tile slots, each with its own device scratch, and three
// CUDA streams (copy-in, kernels, copy-out) chained by events:
//
//   submit(t)   H2D(t) on stream_in -> kernels(t) on the compute stream -> on stream_out: {n_pairs, overflow} and the
//               per-row counts (= rle_right, IJ:1604), then a SPECULATIVE copy of the first est(t) pairs, est(t) from
//               the pairs-per-row of the last collected tile.  Nothing here waits for the device.
//   collect(t)  waits for the scalars only; copies the remainder [est, n_pairs) if the estimate was short (the first
//               tile of a stream, or a change of fan-out), waits for the copy-out, hands the pinned buffers over.
//
// Output buffers are pinned host memory from the library's pool (sq_host_alloc), fresh for every tile, sized from the
// estimate; the caller owns them after collect and returns them with sq_host_free (an Arrow buffer's release callback).
// A tile whose pairs outgrow the device buffer is re-emitted into a larger one at collect (the count is exact by then),
// the GPU analogue of retrying on SQ_ECAPACITY.  right_idx is not moved at all by default: it is the run-length
// expansion of the counts (IJ:1611-1618), which the consumer can fuse into its own pass.
#include "sq_internal.cuh"

#define SQ_API extern "C" __attribute__((visibility("default")))

using namespace sq;

struct sq_tile_slot {
  sq_stream* sub = nullptr;  // private device scratch of this slot; runs on the parent's compute stream
  bool busy = false;
  uint64_t ticket = 0;
  uint32_t n_rows = 0, flags = 0;
  const sq_index* idx = nullptr;
  sq_buf h_scalar;
  uint32_t* h_left = nullptr;
  uint32_t* h_right = nullptr;
  uint32_t* h_counts = nullptr;
  uint64_t h_cap = 0;    // pairs the pinned pair buffers hold
  uint64_t dev_cap = 0;  // pairs the device pair buffers were offered
  uint64_t spec = 0;     // pairs whose copy-out is already enqueued
  bool staged = false;   // served by the staged kernel: its report feeds the stream's kernel choice
  uint32_t key_bytes = 8; // bytes per key on the wire (8: hash, 4: dictionary id)
  cudaEvent_t ev[6] = {};  // 0 h2d begin, 1 h2d end, 2 kernels end, 3 scalars + counts arrived, 4 d2h begin, 5 d2h end
  bool ev_ready = false;
};

namespace sq {

void pipeline_destroy(sq_stream* s) {
  if (s->stream_in) cudaStreamSynchronize(s->stream_in);
  if (s->stream_out) cudaStreamSynchronize(s->stream_out);
  for (sq_tile_slot* sl : s->slots) {
    if (sl->h_left) sq_host_free(s->ctx, sl->h_left);
    if (sl->h_right) sq_host_free(s->ctx, sl->h_right);
    if (sl->h_counts) sq_host_free(s->ctx, sl->h_counts);
    release(sl->h_scalar);
    if (sl->ev_ready) for (auto& e : sl->ev) cudaEventDestroy(e);
    if (sl->sub) sq_stream_free(sl->sub);
    delete sl;
  }
  s->slots.clear();
  release(s->d_dict);
  if (s->stream_in) cudaStreamDestroy(s->stream_in);
  if (s->stream_out) cudaStreamDestroy(s->stream_out);
  s->stream_in = s->stream_out = nullptr;
}

uint64_t pipeline_bytes(const sq_stream* s) {
  uint64_t t = 0;
  for (const sq_tile_slot* sl : s->slots)
    if (sl->sub) t += sq_stream_bytes(sl->sub);
  return t;
}

}  // namespace sq

static int32_t pipeline_init_slots(sq_stream* s);

static int32_t pipeline_init(sq_stream* s) {
  if (!s->slots.empty()) return SQ_OK;
  const int32_t rc = pipeline_init_slots(s);
  if (rc != SQ_OK) pipeline_destroy(s);  // never leave half-built slots behind: the next submit starts over
  return rc;
}

static int32_t pipeline_init_slots(sq_stream* s) {
  ErrorSlot& E = s->err;
  SQ_CUDA(E, cudaStreamCreateWithFlags(&s->stream_in, cudaStreamNonBlocking));
  SQ_CUDA(E, cudaStreamCreateWithFlags(&s->stream_out, cudaStreamNonBlocking));
  const int depth = s->ctx->opt.pipeline_depth.load(std::memory_order_relaxed);
  for (int i = 0; i < depth; ++i) {
    auto* sl = new sq_tile_slot();
    s->slots.push_back(sl);
    int rc = sq_stream_create_on(s->ctx, s->stream, &sl->sub);
    if (rc != SQ_OK) return fail(E, rc, "%s", sq_last_error(s->ctx));
    for (auto& e : sl->ev) SQ_CUDA(E, cudaEventCreate(&e));
    sl->ev_ready = true;
    if ((rc = ensure(E, sl->h_scalar, 256, true))) return rc;
  }
  return SQ_OK;
}

static void drop_outputs(sq_stream* s, sq_tile_slot* sl) {
  if (sl->h_left) sq_host_free(s->ctx, sl->h_left);
  if (sl->h_right) sq_host_free(s->ctx, sl->h_right);
  if (sl->h_counts) sq_host_free(s->ctx, sl->h_counts);
  sl->h_left = sl->h_right = sl->h_counts = nullptr;
  sl->h_cap = 0;
}

static int32_t pinned(sq_stream* s, size_t bytes, uint32_t** out) {
  void* p = nullptr;
  const int rc = sq_host_alloc(s->ctx, bytes ? bytes : 4, &p);
  if (rc != SQ_OK) return fail(s->err, rc, "%s", sq_last_error(s->ctx));
  *out = static_cast<uint32_t*>(p);
  return SQ_OK;
}

// key_hash: u64 hash per row, or (key_id != nullptr) 4-byte ids into the stream's key dictionary
static int32_t submit_tile(sq_stream* s, const sq_index* idx, const uint64_t* key_hash, const uint32_t* key_id, const int32_t* start,
                           const int32_t* end, uint32_t n_rows, uint32_t flags, uint64_t* ticket_out) {
  if (!s) return SQ_EINVAL;
  ErrorSlot& E = s->err;
  if (!idx || !ticket_out) return fail(E, SQ_EINVAL, "null index or ticket pointer");
  if (n_rows && ((!key_hash && !key_id) || !start || !end)) return fail(E, SQ_EINVAL, "null probe column");
  if (key_id && !s->d_dict.p) return fail(E, SQ_ESTATE, "sq_stream_submit_ids without sq_stream_set_key_dictionary");
  if (flags & ~(SQ_TILE_COUNT_ONLY | SQ_TILE_RIGHT_IDX | SQ_TILE_EXPAND_RIGHT | SQ_TILE_NO_COUNTS | SQ_TILE_COUNTS_U8))
    return fail(E, SQ_EINVAL, "unknown tile flag bits 0x%x", flags);
  if ((flags & SQ_TILE_COUNTS_U8) && (flags & (SQ_TILE_EXPAND_RIGHT | SQ_TILE_NO_COUNTS)))
    return fail(E, SQ_EINVAL, "SQ_TILE_COUNTS_U8 goes with neither SQ_TILE_EXPAND_RIGHT (it reads 4-byte counts) nor SQ_TILE_NO_COUNTS");
  if (idx->ctx->device != s->ctx->device)
    return fail(E, SQ_EINVAL, "index lives on device %d, stream on %d", idx->ctx->device, s->ctx->device);
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  int rc;
  if ((rc = pipeline_init(s))) return rc;
  const size_t depth = s->slots.size();
  if (s->next_ticket - s->oldest_ticket >= depth)
    return fail(E, SQ_EBUSY, "%zu tiles in flight: collect ticket %llu first", depth, (unsigned long long)s->oldest_ticket);
  sq_tile_slot* sl = s->slots[s->next_ticket % depth];
  sq_stream* sub = sl->sub;
  // a call that fails half-way may have enqueued work on this slot's buffers: drain the streams before the slot can be
  // handed out again (the slot is only marked busy on success)
  struct Drain {
    sq_stream* s;
    bool armed;
    ~Drain() {
      if (!armed) return;
      cudaStreamSynchronize(s->stream_in);
      cudaStreamSynchronize(s->stream);
      cudaStreamSynchronize(s->stream_out);
      cudaGetLastError();
    }
  } drain{s, true};
  const bool count_only = (flags & SQ_TILE_COUNT_ONLY) != 0;
  const bool want_right = !count_only && (flags & SQ_TILE_RIGHT_IDX) != 0;
  const bool counts_u8 = (flags & SQ_TILE_COUNTS_U8) != 0;
  const size_t n = n_rows;

  // buffers of this tile, sized from the fan-out of the last collected tile (first tile: a guess, nothing speculative)
  uint64_t est = 0, dev_cap = 0;
  if (!count_only && n) {
    if (s->pairs_per_row >= 0.0) est = uint64_t(s->pairs_per_row * 1.02 * double(n)) + 256;
    dev_cap = est ? est + est / 4 + 1024 : n * 8 + 1024;
  }
  if (n) {
    if ((rc = ensure(E, sub->d_in, n * (key_id ? 20 : 16), false))) return rc;
    if (dev_cap && (rc = ensure(E, sub->d_left, dev_cap * 4, false))) return rc;
    if (want_right && (rc = ensure(E, sub->d_right, dev_cap * 4, false))) return rc;
    if (counts_u8 && (rc = ensure(E, sub->d_cnt8, n + 16, false))) return rc;
  }
  drop_outputs(s, sl);
  if ((flags & SQ_TILE_NO_COUNTS) == 0 && n && (rc = pinned(s, n * 4, &sl->h_counts))) return rc;
  if (dev_cap) {
    if ((rc = pinned(s, dev_cap * 4, &sl->h_left))) { drop_outputs(s, sl); return rc; }
    if (want_right && (rc = pinned(s, dev_cap * 4, &sl->h_right))) { drop_outputs(s, sl); return rc; }
    sl->h_cap = dev_cap;
  }
  if (est > dev_cap) est = dev_cap;
  sl->dev_cap = dev_cap;
  sl->spec = 0;
  sl->n_rows = n_rows;
  sl->flags = flags;
  sl->idx = idx;
  auto* hs = static_cast<unsigned long long*>(sl->h_scalar.p);
  hs[0] = hs[1] = hs[2] = hs[3] = hs[4] = 0;

  auto* dk = static_cast<uint64_t*>(sub->d_in.p);
  auto* ds = reinterpret_cast<int32_t*>(dk + n);
  auto* de = ds + n;
  tile_begin(sub, idx, dk, ds, de, n_rows);
  if (n) {
    // ---- copy-in
    SQ_CUDA(E, cudaEventRecord(sl->ev[0], s->stream_in));
    auto* d_ids = reinterpret_cast<uint32_t*>(de + n);
    if (key_id) SQ_CUDA(E, cudaMemcpyAsync(d_ids, key_id, n * 4, cudaMemcpyHostToDevice, s->stream_in));
    else SQ_CUDA(E, cudaMemcpyAsync(dk, key_hash, n * 8, cudaMemcpyHostToDevice, s->stream_in));
    SQ_CUDA(E, cudaMemcpyAsync(ds, start, n * 4, cudaMemcpyHostToDevice, s->stream_in));
    SQ_CUDA(E, cudaMemcpyAsync(de, end, n * 4, cudaMemcpyHostToDevice, s->stream_in));
    SQ_CUDA(E, cudaEventRecord(sl->ev[1], s->stream_in));
    // ---- kernels: the whole count -> scan -> write chain is enqueued without a host round trip; a tile that does
    // not fit dev_cap reports it in result[1] and is re-emitted at collect
    SQ_CUDA(E, cudaStreamWaitEvent(s->stream, sl->ev[1], 0));
    if (key_id && (rc = launch_expand_ids(s, s->stream, d_ids, static_cast<const uint64_t*>(s->d_dict.p), s->dict_n, n_rows, dk, ds, de)))
      return rc;
    const uint64_t* host_key = key_id ? nullptr : key_hash;  // the staged kernel's look at the row order needs the hashes
    uint32_t* d_left = dev_cap ? static_cast<uint32_t*>(sub->d_left.p) : nullptr;
    uint32_t* d_right = want_right ? static_cast<uint32_t*>(sub->d_right.p) : nullptr;
    sl->staged = false;
    if (use_packed(idx)) {
      sl->staged = pick_staged(s, idx, host_key, start, nullptr, nullptr, n_rows, d_left != nullptr, key_id);
      if ((rc = launch_packed_any(sub, s, sl->staged, idx, dk, ds, de, n_rows, d_left, d_right, dev_cap)))
        return fail(E, rc, "%s", sub->err.msg.c_str());
    } else if (use_rank(idx)) {
      if ((rc = launch_rank_join(sub, idx, dk, ds, de, n_rows, d_left, d_right, dev_cap))) return fail(E, rc, "%s", sub->err.msg.c_str());
    } else {
      if ((rc = launch_count(sub, idx, dk, ds, de, n_rows))) return fail(E, rc, "%s", sub->err.msg.c_str());
      if (d_left && (rc = launch_write(sub, idx, ds, n_rows, d_left, d_right, dev_cap))) return fail(E, rc, "%s", sub->err.msg.c_str());
    }
    // counts as bytes: the pinned buffer still has room for the 4-byte counts, which follow at collect if a count overflows
    if (counts_u8 && (rc = launch_narrow_counts(sub, s->stream, static_cast<const uint32_t*>(sub->d_cnt.p), n_rows,
                                               static_cast<uint8_t*>(sub->d_cnt8.p),
                                               static_cast<unsigned long long*>(sub->d_scalar.p) + 4)))
      return fail(E, rc, "%s", sub->err.msg.c_str());
    s->launches += sub->launches;
    sub->launches = 0;
    SQ_CUDA(E, cudaEventRecord(sl->ev[2], s->stream));
    // ---- copy-out
    SQ_CUDA(E, cudaStreamWaitEvent(s->stream_out, sl->ev[2], 0));
    SQ_CUDA(E, cudaEventRecord(sl->ev[4], s->stream_out));
    SQ_CUDA(E, cudaMemcpyAsync(hs, sub->d_scalar.p, 40, cudaMemcpyDeviceToHost, s->stream_out));
    if (sl->h_counts && counts_u8) SQ_CUDA(E, cudaMemcpyAsync(sl->h_counts, sub->d_cnt8.p, n, cudaMemcpyDeviceToHost, s->stream_out));
    else if (sl->h_counts) SQ_CUDA(E, cudaMemcpyAsync(sl->h_counts, sub->d_cnt.p, n * 4, cudaMemcpyDeviceToHost, s->stream_out));
    SQ_CUDA(E, cudaEventRecord(sl->ev[3], s->stream_out));
    if (est) {
      SQ_CUDA(E, cudaMemcpyAsync(sl->h_left, d_left, est * 4, cudaMemcpyDeviceToHost, s->stream_out));
      if (want_right) SQ_CUDA(E, cudaMemcpyAsync(sl->h_right, d_right, est * 4, cudaMemcpyDeviceToHost, s->stream_out));
      sl->spec = est;
    }
    SQ_CUDA(E, cudaEventRecord(sl->ev[5], s->stream_out));
  }
  drain.armed = false;
  sl->busy = true;
  sl->ticket = s->next_ticket++;
  sl->key_bytes = key_id ? 4 : 8;
  *ticket_out = sl->ticket;
  return SQ_OK;
}

SQ_API int32_t sq_stream_submit(sq_stream* s, const sq_index* idx, const uint64_t* key_hash, const int32_t* start,
                                const int32_t* end, uint32_t n_rows, uint32_t flags, uint64_t* ticket_out) {
  return submit_tile(s, idx, key_hash, nullptr, start, end, n_rows, flags, ticket_out);
}

SQ_API int32_t sq_stream_submit_ids(sq_stream* s, const sq_index* idx, const uint32_t* key_id, const int32_t* start,
                                    const int32_t* end, uint32_t n_rows, uint32_t flags, uint64_t* ticket_out) {
  if (s && n_rows && !key_id) return fail(s->err, SQ_EINVAL, "null probe column");
  return submit_tile(s, idx, nullptr, key_id, start, end, n_rows, flags, ticket_out);
}

SQ_API int32_t sq_stream_set_key_dictionary(sq_stream* s, const uint64_t* key_hashes, uint32_t n_entries) {
  if (!s) return SQ_EINVAL;
  ErrorSlot& E = s->err;
  if (n_entries && !key_hashes) return fail(E, SQ_EINVAL, "null key dictionary");
  if (s->next_ticket != s->oldest_ticket) return fail(E, SQ_ESTATE, "tiles in flight still read the previous dictionary: collect them first");
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  int rc;
  if ((rc = ensure(E, s->d_dict, size_t(n_entries ? n_entries : 1) * 8, false))) return rc;
  if (n_entries) SQ_CUDA(E, cudaMemcpy(s->d_dict.p, key_hashes, size_t(n_entries) * 8, cudaMemcpyHostToDevice));
  s->dict_n = n_entries;
  return SQ_OK;
}

SQ_API int32_t sq_stream_collect(sq_stream* s, uint64_t ticket, sq_tile_out* out) {
  if (!s) return SQ_EINVAL;
  ErrorSlot& E = s->err;
  if (!out) return fail(E, SQ_EINVAL, "null output record");
  memset(out, 0, sizeof *out);
  if (s->slots.empty() || ticket != s->oldest_ticket || ticket >= s->next_ticket)
    return fail(E, SQ_ESTATE, "ticket %llu is not the oldest tile in flight (%llu); tiles are collected in submission order",
                (unsigned long long)ticket, (unsigned long long)s->oldest_ticket);
  sq_tile_slot* sl = s->slots[ticket % s->slots.size()];
  sq_stream* sub = sl->sub;
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  const size_t n = sl->n_rows;
  const bool count_only = (sl->flags & SQ_TILE_COUNT_ONLY) != 0;
  const bool want_right = !count_only && (sl->flags & SQ_TILE_RIGHT_IDX) != 0;
  uint64_t n_pairs = 0;
  uint32_t counts_width = 4;
  int rc = SQ_OK;
  auto bail = [&](int code) {  // the tile is gone either way: free its slot and its buffers
    drop_outputs(s, sl);
    sl->busy = false;
    s->oldest_ticket += 1;
    return code;
  };
  if (n) {
    cudaError_t ce = cudaEventSynchronize(sl->ev[3]);
    if (ce != cudaSuccess) return bail(fail(E, SQ_ECUDA, "tile %llu failed on the device: %s", (unsigned long long)ticket, cudaGetErrorString(ce)));
    auto* hs = static_cast<unsigned long long*>(sl->h_scalar.p);
    n_pairs = hs[0];
    if (sl->staged) staged_feedback(s, sl->n_rows, hs[2], hs[3]);
    const bool overflow = hs[1] != 0 || n_pairs > sl->dev_cap;
    uint64_t copied = sl->spec;
    if ((sl->flags & SQ_TILE_COUNTS_U8) && sl->h_counts) {
      counts_width = 1;
      if (hs[4]) {  // some row has more than 255 hits: the 4-byte counts after all (they are still in the slot's scratch)
        ce = cudaMemcpyAsync(sl->h_counts, sub->d_cnt.p, n * 4, cudaMemcpyDeviceToHost, s->stream_out);
        if (ce == cudaSuccess) ce = cudaEventRecord(sl->ev[5], s->stream_out);
        if (ce != cudaSuccess) return bail(fail(E, SQ_ECUDA, "D2H copy of the counts: %s", cudaGetErrorString(ce)));
        counts_width = 4;
      }
    }
    if (!count_only && n_pairs) {
      if (overflow) {  // the device buffers were too small for this tile: grow them, re-run only the write pass
        if ((rc = ensure(E, sub->d_left, n_pairs * 4, false))) return bail(rc);
        if (want_right && (rc = ensure(E, sub->d_right, n_pairs * 4, false))) return bail(rc);
        sub->n_pairs = n_pairs;
        if ((rc = tile_emit(sub, static_cast<uint32_t*>(sub->d_left.p), want_right ? static_cast<uint32_t*>(sub->d_right.p) : nullptr, n_pairs)))
          return bail(fail(E, rc, "%s", sub->err.msg.c_str()));
        s->launches += sub->launches;
        sub->launches = 0;
        cudaEventRecord(sl->ev[2], s->stream);
        cudaStreamWaitEvent(s->stream_out, sl->ev[2], 0);
        copied = 0;
        s->pipe_regrow += 1;
      }
      if (n_pairs > sl->h_cap) {  // ... and so were the pinned ones
        uint32_t* counts = sl->h_counts;
        sl->h_counts = nullptr;
        drop_outputs(s, sl);
        sl->h_counts = counts;
        if ((rc = pinned(s, n_pairs * 4, &sl->h_left))) return bail(rc);
        if (want_right && (rc = pinned(s, n_pairs * 4, &sl->h_right))) return bail(rc);
        sl->h_cap = n_pairs;
        copied = 0;
      }
      if (n_pairs > copied) {
        const auto* dl = static_cast<const uint32_t*>(sub->d_left.p);
        const auto* dr = static_cast<const uint32_t*>(sub->d_right.p);
        ce = cudaMemcpyAsync(sl->h_left + copied, dl + copied, (n_pairs - copied) * 4, cudaMemcpyDeviceToHost, s->stream_out);
        if (ce == cudaSuccess && want_right)
          ce = cudaMemcpyAsync(sl->h_right + copied, dr + copied, (n_pairs - copied) * 4, cudaMemcpyDeviceToHost, s->stream_out);
        if (ce == cudaSuccess) ce = cudaEventRecord(sl->ev[5], s->stream_out);
        if (ce != cudaSuccess) return bail(fail(E, SQ_ECUDA, "D2H copy of the pairs: %s", cudaGetErrorString(ce)));
      }
    }
    ce = cudaEventSynchronize(sl->ev[5]);
    if (ce != cudaSuccess) return bail(fail(E, SQ_ECUDA, "tile %llu copy-out: %s", (unsigned long long)ticket, cudaGetErrorString(ce)));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sl->ev[0], sl->ev[1]) == cudaSuccess) s->pipe_ms[0] += ms;
    if (cudaEventElapsedTime(&ms, sl->ev[1], sl->ev[2]) == cudaSuccess) s->pipe_ms[1] += ms;
    if (cudaEventElapsedTime(&ms, sl->ev[4], sl->ev[5]) == cudaSuccess) s->pipe_ms[2] += ms;
    cudaGetLastError();
    s->pipe_bytes[0] += n * (8 + sl->key_bytes);
    s->pipe_bytes[1] += 16 + (sl->h_counts ? n * ((sl->flags & SQ_TILE_COUNTS_U8) ? (counts_width == 4 ? 5 : 1) : 4) : 0) + (count_only ? 0 : (want_right ? 8 : 4) * (n_pairs > sl->spec ? n_pairs : sl->spec));
    s->pipe_tiles += 1;
    if (!count_only) s->pairs_per_row = double(n_pairs) / double(n);
  }
  if ((sl->flags & SQ_TILE_EXPAND_RIGHT) && !count_only && !want_right && n_pairs) {
    // right_idx from the counts, on the calling thread (IJ:1611-1618)
    if (!sl->h_counts) return bail(fail(E, SQ_EINVAL, "SQ_TILE_EXPAND_RIGHT needs the counts (SQ_TILE_NO_COUNTS was set)"));
    if ((rc = pinned(s, n_pairs * 4, &sl->h_right))) return bail(rc);
    expand_counts(sl->h_counts, sl->n_rows, sl->h_right, n_pairs);
  }
  out->n_pairs = n_pairs;
  out->n_rows = sl->n_rows;
  out->left_idx = sl->h_left;
  out->right_idx = sl->h_right;
  out->counts = sl->h_counts;
  out->counts_width = sl->h_counts ? counts_width : 0;
  sl->h_left = sl->h_right = sl->h_counts = nullptr;  // the caller owns them now (sq_host_free)
  sl->h_cap = 0;
  sl->busy = false;
  s->oldest_ticket += 1;
  return SQ_OK;
}

SQ_API int32_t sq_stream_in_flight(const sq_stream* s) { return s ? int32_t(s->next_ticket - s->oldest_ticket) : 0; }

SQ_API int32_t sq_stream_pipeline_stats(const sq_stream* s, double out8[8]) {
  if (!s || !out8) return SQ_EINVAL;
  out8[0] = s->pipe_ms[0];
  out8[1] = s->pipe_ms[1];
  out8[2] = s->pipe_ms[2];
  out8[3] = double(s->pipe_bytes[0]);
  out8[4] = double(s->pipe_bytes[1]);
  out8[5] = double(s->pipe_tiles);
  out8[6] = double(s->pipe_regrow);
  out8[7] = s->pairs_per_row;
  return SQ_OK;
}
