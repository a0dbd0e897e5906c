// Probe + emit over the PACKED build index (128-byte lines of 15 rows, see sq_internal.cuh) in ONE
// kernel: replaces the per-row loop of process_probe_batch (reference interval_join.rs:1586-1618:
// hash_map.get -> coitrees query -> pos_vect / rle_right -> index_right) for "narrow" indexes.
//
// Why this shape (measured on B200, tools/ubench_gather.cu, 1.6 GB buffer):
//   * random reads cost ~21 ps per L2 request regardless of size up to 32 B, 24.6 ps per 64 B and
//     41.6 ps per 128 B when ONE instruction of a lane group fetches the whole granule, but 39 ps /
//     78 ps when one thread fetches the same bytes with consecutive 16-byte loads: requests, not
//     bytes, are the currency, and a line must arrive as one cooperative request;
//   * the SoA walk (sq_probe.cu) needs ~7 distinct 64-byte granules and ~22 sector requests per probe
//     row (directory, starts, runmax, ends, rows), this kernel 1 directory word + ~1.6 lines.
//
//   phase 1  lane i owns probe row i of the warp: key hash -> key id -> segment meta -> ONE directory
//            word gives the line holding the last start <= probe end.
//   phase 2  8 lanes serve one probe row (4 rows per warp instruction): load the line, test the exact
//            predicate (start <= qe && end >= qs) on its 15 rows, step to the previous line while the
//            line's exmax (max end of all earlier rows) still reaches qs.  Hits go to a per-row stash
//            in shared memory (32 slots).
//   phase 3  CTA total -> chained scan with decoupled look-back across CTAs (tiles are handed out by a
//            ticket, so predecessors always run) -> output base of the CTA.
//   phase 4  rows with <= 32 hits are copied from the stash as ONE flattened list per warp (coalesced
//            stores of left_idx / right_idx); rows with more hits are re-walked by the whole warp, four
//            lines per step.
// Nothing per row is written to HBM except its hit count (= rle_right, interval_join.rs:1604) and its
// pairs; integer work bounded by HBM request rate, tensor cores do not apply.
#include <cstdlib>

#include "sq_internal.cuh"
#include "sq_packed_common.cuh"

namespace sq {


// kTiles consecutive tiles per CTA: the probe columns and directory words of ALL of them are requested up
// front, so only the first tile of a CTA waits for those two round trips (count-only launches use 2; see
// launch_packed_b for why emitting launches use 1).
#ifndef SQ_PACKED_THREADS_PER_SM
#define SQ_PACKED_THREADS_PER_SM 1024  // resident threads per SM the kernel is compiled for (register budget)
#endif
template <bool EMIT, bool WRITE_RIGHT, int kPBlock, int kTiles>
__global__ void __launch_bounds__(kPBlock, (EMIT && kTiles > 1 ? 768 : SQ_PACKED_THREADS_PER_SM) / kPBlock)
k_probe_packed(IndexView iv, const uint64_t* __restrict__ q_key, const int32_t* __restrict__ q_start,
               const int32_t* __restrict__ q_end, uint32_t n, uint32_t* __restrict__ cnt_out,
               unsigned long long* chain_state, unsigned int* ticket, unsigned long long* result,
               uint32_t* __restrict__ left_out, uint32_t* __restrict__ right_out, uint64_t capacity,
               uint32_t n_tiles, uint32_t backoff_ns) {
  constexpr int kPWarps = kPBlock / 32;
  constexpr int kStashRow = kPWarps * 32 * kStride;  // one tile's stash: every tile of the CTA keeps its own until the emit
  __shared__ uint32_t s_stash[EMIT ? kTiles * kStashRow : 1];
  __shared__ unsigned long long s_wtot[kTiles][kPWarps];
  __shared__ unsigned long long s_base[kTiles];
  __shared__ uint32_t s_bid;
  __shared__ __align__(8) uint8_t s_inv[kPWarps][32];
  __shared__ WalkShared s_walk[kPWarps];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  uint32_t bid = blockIdx.x;
  if (EMIT) {  // tiles in ticket order: every predecessor in the chained scan is already running
    if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    bid = s_bid;
  }

  // ---- phase 1 of every tile of this CTA: my probe rows -> their start lines ---------------------------
  int32_t t_qs[kTiles], t_qe[kTiles];
  uint64_t t_key[kTiles];
  StartLine t_sl[kTiles];
#pragma unroll
  for (int t = 0; t < kTiles; ++t) {
    const uint64_t i = (uint64_t(bid) * kTiles + t) * kPBlock + threadIdx.x;
    t_qs[t] = t_qe[t] = 0;
    t_key[t] = 0;
    if (i < n) {
      t_qs[t] = q_start[i];
      t_qe[t] = q_end[i];
      t_key[t] = q_key[i];
    }
  }
#pragma unroll
  for (int t = 0; t < kTiles; ++t) {
    const uint64_t i = (uint64_t(bid) * kTiles + t) * kPBlock + threadIdx.x;
    t_sl[t] = StartLine{0u, 0u, false};
    if (i < n) {
      const uint32_t id = ht_lookup(iv.ht_keys, iv.ht_ids, iv.ht_mask, iv.sentinel_id, t_key[t]);
      t_sl[t] = find_start_line(iv, id, t_qe[t]);
    }
  }

  // ---- phase 2 of every tile: walk in compacted rounds (sq_packed_common.cuh) --------------------------
  // All walks come before any look-back: the CTA publishes the totals of ALL its tiles at once, so a later
  // CTA never waits for this one's stores (walking tile 2b+1 only after emitting tile 2b serialised the grid).
  uint32_t t_cnt[kTiles], t_cincl[kTiles];
#pragma unroll
  for (int t = 0; t < kTiles; ++t) {
    const uint64_t i = (uint64_t(bid) * kTiles + t) * kPBlock + threadIdx.x;  // rows past n / tiles past n_tiles: not walking
    uint32_t* stash = s_stash + (EMIT ? t * kStashRow + warp * 32 * kStride : 0);
    bool walking = t_sl[t].act;
    uint32_t ln = t_sl[t].line;  // next line of my row
    uint32_t cnt = 0;            // hits of my row so far
    walk_rounds<EMIT>(iv, stash, s_walk[warp], t_qs[t], t_qe[t], t_sl[t].first, walking, ln, cnt);
    if (i < n) cnt_out[i] = cnt;  // rle_right (interval_join.rs:1604)
    t_cnt[t] = cnt;
    t_cincl[t] = warp_incl_sum(cnt);  // a warp emits < 2^32 pairs unless rows hit > 2^27 builds each
    const uint32_t wtot = __shfl_sync(0xffffffffu, t_cincl[t], 31);
    if (lane == 0) s_wtot[t][warp] = wtot;
  }
  __syncthreads();
  if (!EMIT) {  // count only: the grand total is order-free; one atomic per CTA (same-address atomics
                // serialise at ~2.5 ns each: one per warp would cost 1 ms per 12.5M rows by itself)
    if (threadIdx.x == 0) {
      unsigned long long tot = 0;
#pragma unroll
      for (int t = 0; t < kTiles; ++t)
#pragma unroll
        for (int w = 0; w < kPWarps; ++w) tot += s_wtot[t][w];
      if (tot) atomicAdd(result, tot);
    }
    return;
  }

  // ---- phase 3: tile totals -> chained scan ------------------------------------------------------------
  const uint32_t tile0 = bid * kTiles;
  if (warp == 0) {
    unsigned long long agg[kTiles];
#pragma unroll
    for (int t = 0; t < kTiles; ++t) {
      agg[t] = 0;
#pragma unroll
      for (int w = 0; w < kPWarps; ++w) agg[t] += s_wtot[t][w];
    }
    // later tiles of this CTA first, as aggregates: a successor's look-back passes over them at once
#pragma unroll
    for (int t = 1; t < kTiles; ++t)
      if (lane == 0 && tile0 + t < n_tiles) atomicExch(chain_state + tile0 + t, kFlagAgg | agg[t]);
    unsigned long long run = chain_lookback(chain_state, tile0, agg[0], backoff_ns);
    if (lane == 0) {
#pragma unroll
      for (int t = 0; t < kTiles; ++t) {
        if (tile0 + t >= n_tiles) break;
        s_base[t] = run;
        run += agg[t];
        if (t > 0) atomicExch(chain_state + tile0 + t, kFlagInc | run);
        if (tile0 + t == n_tiles - 1) result[0] = run;
        if (run > capacity) result[1] = 1;  // the caller's buffers are too small: report, write nothing here
      }
    }
  }
  __syncthreads();
  // ---- phase 4: ordered emit (stash -> flattened coalesced stores; rows with > 32 hits re-walked) -------
#pragma unroll
  for (int t = 0; t < kTiles; ++t) {
    if (tile0 + t >= n_tiles) break;  // CTA-uniform
    const uint32_t tile_first = (tile0 + t) * kPBlock + warp * 32;
    uint64_t base = s_base[t];
    unsigned long long cta_tot = 0;
#pragma unroll
    for (int w = 0; w < kPWarps; ++w) {
      if (w < warp) base += s_wtot[t][w];
      cta_tot += s_wtot[t][w];
    }
    const uint32_t wtot = uint32_t(s_wtot[t][warp]);
    if (wtot != 0 && s_base[t] + cta_tot <= capacity)
      emit_rows<WRITE_RIGHT>(iv, s_stash + t * kStashRow + warp * 32 * kStride, s_inv[warp], t_cnt[t], t_cincl[t] - t_cnt[t],
                             t_qs[t], t_qe[t], t_sl[t].line, t_sl[t].first, left_out + base,
                             WRITE_RIGHT ? right_out + base : nullptr, tile_first);
    __syncwarp();  // s_inv is reused by the next tile
  }
}

// ---------------------------------------------------------------------------------------------
bool use_packed(const sq_index* idx) {
  if (!idx->d_lines) return false;
  // option cuda_probe_layout: soa forces the SoA kernels, packed the packed-line kernel (tests, A/B runs)
  const int forced = idx->ctx->opt.probe_layout.load(std::memory_order_relaxed);
  if (forced == 2) return false;
  if (forced == 1) return true;
  // Measured on B200: the fused kernel wins when the index is far larger than L2 (cfg5, 100M rows:
  // 1.05 ms vs 1.83 ms per 12.5M probe rows) and loses on an L2-resident one (cfg2, 1M rows: 0.085
  // vs 0.067 ms), where the SoA kernels' loads are cache hits and the chained scan is pure overhead.
  // It also loses on deep overlap (cfg3 forced through it: 2.4 ms vs 1.3 ms): a row walks one line per
  // round and rows with more than 32 hits are re-walked one at a time, so the index must be shallow —
  // mean_back_lines is the build-time estimate of how many extra lines a probe walks.
  return idx->n_lines * 128ull > (64ull << 20) && idx->mean_back_lines <= 1.5f;
}

// Every CTA takes two consecutive tiles (count-only measured 0.69 vs 0.77 ms per 12.5M rows: the second tile's probe
// columns and directory word are already there).  An emitting CTA walks BOTH tiles before it joins the chained scan and
// publishes both totals together; emitting tile 2b before walking tile 2b+1 would make CTA b+1's look-back wait for
// CTA b's stores and serialise the grid (measured: 400 ms).  Option cuda_probe_tiles picks 1 or 2 for emitting launches.
template <int B>
static void launch_packed_b(sq_stream* s, const IndexView& iv, const uint64_t* d_key, const int32_t* d_start,
                            const int32_t* d_end, uint32_t n, uint32_t* cnt, unsigned long long* chain, unsigned int* ticket,
                            unsigned long long* result, uint32_t* d_left, uint32_t* d_right, uint64_t capacity) {
  const uint32_t n_tiles = (n + B - 1) / B;
  const uint32_t backoff = uint32_t(s->ctx->opt.lookback_backoff_ns.load(std::memory_order_relaxed));
  if (!d_left)
    k_probe_packed<false, false, B, 2><<<(n_tiles + 1) / 2, B, 0, s->stream>>>(iv, d_key, d_start, d_end, n, cnt, chain, ticket,
                                                                                result, nullptr, nullptr, 0, n_tiles, 0u);
  else if (B <= 128 && s->ctx->opt.probe_tiles.load(std::memory_order_relaxed) == 2) {  // two stashes of 256 rows: > 48 KB
    constexpr int B2 = B <= 128 ? B : 128;
    if (d_right)
      k_probe_packed<true, true, B2, 2><<<(n_tiles + 1) / 2, B2, 0, s->stream>>>(iv, d_key, d_start, d_end, n, cnt, chain, ticket,
                                                                                result, d_left, d_right, capacity, n_tiles, backoff);
    else
      k_probe_packed<true, false, B2, 2><<<(n_tiles + 1) / 2, B2, 0, s->stream>>>(iv, d_key, d_start, d_end, n, cnt, chain, ticket,
                                                                                 result, d_left, nullptr, capacity, n_tiles, backoff);
  } else if (d_right)
    k_probe_packed<true, true, B, 1><<<n_tiles, B, 0, s->stream>>>(iv, d_key, d_start, d_end, n, cnt, chain, ticket, result,
                                                                    d_left, d_right, capacity, n_tiles, backoff);
  else
    k_probe_packed<true, false, B, 1><<<n_tiles, B, 0, s->stream>>>(iv, d_key, d_start, d_end, n, cnt, chain, ticket, result,
                                                                     d_left, nullptr, capacity, n_tiles, backoff);
}

int launch_packed(sq_stream* s, const sq_index* idx, const uint64_t* d_key, const int32_t* d_start,
                  const int32_t* d_end, uint32_t n, uint32_t* d_left, uint32_t* d_right, uint64_t capacity) {
  ErrorSlot& E = s->err;
  const int block = s->ctx->opt.probe_block.load(std::memory_order_relaxed);
  const uint32_t n_tiles = (n + block - 1) / block;
  int rc;
  if ((rc = ensure(E, s->d_cnt, size_t(n) * 4, false))) return rc;
  if ((rc = ensure(E, s->d_tile, size_t(n_tiles) * 8 + 16, false))) return rc;
  if ((rc = ensure(E, s->d_scalar, 256, false))) return rc;
  auto* chain = static_cast<unsigned long long*>(s->d_tile.p);
  auto* ticket = reinterpret_cast<unsigned int*>(chain + n_tiles);
  auto* result = static_cast<unsigned long long*>(s->d_scalar.p);  // [0] n_pairs [1] overflow
  auto* cnt = static_cast<uint32_t*>(s->d_cnt.p);
  SQ_CUDA(E, cudaMemsetAsync(result, 0, 32, s->stream));
  const IndexView iv = idx->view();
  if (s->ctx->l2_persist_bytes && s->l2_window_idx != idx && idx->dir_bytes) {
    // optional (option cuda_l2_persist_mb): keep the line directory (8-byte entries, same entry count as the 4-byte row
    // directory), the one structure every probe row reads at a random place, in the persisting part of L2;
    // everything else streams through the rest
    cudaStreamAttrValue av{};
    size_t win = size_t(idx->dir_bytes) * 2;
    if (win > s->ctx->l2_window_max) win = s->ctx->l2_window_max;
    av.accessPolicyWindow.base_ptr = idx->d_dir_line;
    av.accessPolicyWindow.num_bytes = win;
    const double ratio = double(s->ctx->l2_persist_bytes) / double(win);
    av.accessPolicyWindow.hitRatio = float(ratio > 1.0 ? 1.0 : ratio);
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(s->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
    s->l2_window_idx = idx;
  }
  if (d_left) SQ_CUDA(E, cudaMemsetAsync(chain, 0, size_t(n_tiles) * 8 + 16, s->stream));
  if (block == 64) launch_packed_b<64>(s, iv, d_key, d_start, d_end, n, cnt, chain, ticket, result, d_left, d_right, capacity);
  else if (block == 256) launch_packed_b<256>(s, iv, d_key, d_start, d_end, n, cnt, chain, ticket, result, d_left, d_right, capacity);
  else launch_packed_b<128>(s, iv, d_key, d_start, d_end, n, cnt, chain, ticket, result, d_left, d_right, capacity);
  SQ_CUDA(E, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

}  // namespace sq
