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

constexpr int kPBlockDefault = 128;      // threads per CTA = probe rows per CTA

struct StartLine {
  uint32_t line;   // line holding the last row whose start can be <= qe
  uint32_t first;  // first line of the key segment
  bool act;        // false: no row of the build side can match
};

__device__ __forceinline__ StartLine find_start_line(const IndexView& iv, uint32_t id, int32_t qe) {
  StartLine r{0u, 0u, false};
  if (id == kNoKey) return r;  // key hash absent from the build side: no rows (interval_join.rs:965)
  const SegMeta m = iv.meta[id];
  if (qe < m.min_start) return r;
  const uint32_t off = uint32_t(qe) - uint32_t(m.min_start);
  const uint32_t b = m.shift >= 32 ? 0u : (off >> m.shift);
  // first row whose bin is > b: every row with start <= qe lies below it
  const uint32_t r_end = b >= m.nbins ? m.se : __ldg(iv.dir + m.dir_base + b + 1);
  r.line = m.line_base + (r_end - 1u - m.sb) / kLineRows;
  r.first = m.line_base;
  r.act = true;
  return r;
}

template <bool EMIT, bool WRITE_RIGHT, int kPBlock>
__global__ void __launch_bounds__(kPBlock, 1024 / kPBlock)
k_probe_packed(IndexView iv, const uint64_t* __restrict__ q_key, const int32_t* __restrict__ q_start,
               const int32_t* __restrict__ q_end, uint32_t n, uint32_t* __restrict__ cnt_out,
               unsigned long long* chain_state, unsigned int* ticket, unsigned long long* result,
               uint32_t* __restrict__ left_out, uint32_t* __restrict__ right_out, uint64_t capacity) {
  constexpr int kPWarps = kPBlock / 32;
  __shared__ uint32_t s_stash[EMIT ? kPWarps * 32 * kStride : 1];
  __shared__ unsigned long long s_wtot[kPWarps];
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_bid;
  __shared__ __align__(8) uint8_t s_inv[kPWarps][32];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  uint32_t bid = blockIdx.x;
  if (EMIT) {  // tiles in ticket order: every predecessor in the chained scan is already running
    if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    bid = s_bid;
  }
  const uint32_t i = bid * kPBlock + threadIdx.x;
  const uint32_t tile_first = bid * kPBlock + warp * 32;

  // ---- phase 1: my probe row -> start line ----------------------------------------------------------
  int32_t my_qs = 0, my_qe = 0;
  StartLine sl{0u, 0u, false};
  if (i < n) {
    my_qs = q_start[i];
    my_qe = q_end[i];
    const uint32_t id = ht_lookup(iv.ht_keys, iv.ht_ids, iv.ht_mask, iv.sentinel_id, q_key[i]);
    sl = find_start_line(iv, id, my_qe);
  }

  // ---- phase 2: walk in compacted rounds (sq_packed_common.cuh) -----------------------------------
  uint32_t* stash = s_stash + (EMIT ? warp * 32 * kStride : 0);
  bool walking = sl.act;
  uint32_t ln = sl.line;  // next line of my row
  uint32_t cnt = 0;       // hits of my row so far
  walk_rounds<EMIT>(iv, stash, s_inv[warp], my_qs, my_qe, sl.first, walking, ln, cnt);
  if (i < n) cnt_out[i] = cnt;  // rle_right (interval_join.rs:1604)

  const uint32_t cincl = warp_incl_sum(cnt);  // a warp emits < 2^32 pairs unless rows hit > 2^27 builds each
  const uint32_t wtot = __shfl_sync(0xffffffffu, cincl, 31);
  if (!EMIT) {  // count only: the grand total is order-free; one atomic per CTA (same-address atomics
                // serialise at ~2.5 ns each: one per warp would cost 1 ms per 12.5M rows by itself)
    __shared__ unsigned long long s_ctot;
    if (threadIdx.x == 0) s_ctot = 0;
    __syncthreads();
    if (lane == 0 && wtot) atomicAdd(&s_ctot, (unsigned long long)wtot);
    __syncthreads();
    if (threadIdx.x == 0 && s_ctot) atomicAdd(result, s_ctot);
    return;
  }

  // ---- phase 3: CTA total -> chained scan -------------------------------------------------------------
  if (lane == 0) s_wtot[warp] = wtot;
  __syncthreads();
  if (warp == 0) {
    unsigned long long agg = 0;
#pragma unroll
    for (int w = 0; w < kPWarps; ++w) agg += s_wtot[w];
    const unsigned long long excl = chain_lookback(chain_state, bid, agg);
    if (lane == 0) {
      s_base = excl;
      if (bid == gridDim.x - 1) result[0] = excl + agg;
      if (excl + agg > capacity) result[1] = 1;  // the caller's buffers are too small: report, write nothing here
    }
  }
  __syncthreads();
  if (wtot == 0) return;  // warp-uniform
  uint64_t base = s_base;
  unsigned long long cta_tot = 0;
#pragma unroll
  for (int w = 0; w < kPWarps; ++w) {
    if (w < warp) base += s_wtot[w];
    cta_tot += s_wtot[w];
  }
  if (s_base + cta_tot > capacity) return;  // CTA-uniform
  uint32_t* __restrict__ lout = left_out + base;
  uint32_t* __restrict__ rout = WRITE_RIGHT ? right_out + base : nullptr;
  const uint32_t coff = cincl - cnt;  // offset of my row's first pair inside the warp's run

  // ---- phase 4: ordered emit (stash -> flattened coalesced stores; rows with > 32 hits re-walked) ----
  emit_rows<WRITE_RIGHT>(iv, stash, s_inv[warp], cnt, coff, my_qs, my_qe, sl.line, sl.first, lout, rout, tile_first);
}

// ---------------------------------------------------------------------------------------------
bool use_packed(const sq_index* idx) {
  if (!idx->d_lines) return false;
  // experiment / test knobs, read per call: SQ_PACKED=0 forces the SoA kernels, =1 the packed kernel
  const char* e = getenv("SQ_PACKED");
  const int forced = e ? atoi(e) : -1;
  if (forced == 0) return false;
  if (forced == 1) return true;
  // Measured on B200: the fused kernel wins when the index is far larger than L2 (cfg5, 100M rows:
  // 1.15 ms vs 1.83 ms per 12.5M probe rows) and loses on an L2-resident one (cfg2, 1M rows: 0.085
  // vs 0.067 ms), where the SoA kernels' loads are cache hits and the chained scan is pure overhead.
  return idx->n_lines * 128ull > (64ull << 20);
}

template <int B>
static void launch_packed_b(sq_stream* s, const IndexView& iv, const uint64_t* d_key, const int32_t* d_start,
                            const int32_t* d_end, uint32_t n, uint32_t* cnt, unsigned long long* chain, unsigned int* ticket,
                            unsigned long long* result, uint32_t* d_left, uint32_t* d_right, uint64_t capacity) {
  const uint32_t n_tiles = (n + B - 1) / B;
  if (!d_left)
    k_probe_packed<false, false, B><<<n_tiles, B, 0, s->stream>>>(iv, d_key, d_start, d_end, n, cnt, chain, ticket, result,
                                                                   nullptr, nullptr, 0);
  else if (d_right)
    k_probe_packed<true, true, B><<<n_tiles, B, 0, s->stream>>>(iv, d_key, d_start, d_end, n, cnt, chain, ticket, result,
                                                                 d_left, d_right, capacity);
  else
    k_probe_packed<true, false, B><<<n_tiles, B, 0, s->stream>>>(iv, d_key, d_start, d_end, n, cnt, chain, ticket, result,
                                                                  d_left, nullptr, capacity);
}

int launch_packed(sq_stream* s, const sq_index* idx, const uint64_t* d_key, const int32_t* d_start,
                  const int32_t* d_end, uint32_t n, uint32_t* d_left, uint32_t* d_right, uint64_t capacity) {
  ErrorSlot& E = s->err;
  int block = kPBlockDefault;
  if (const char* e = getenv("SQ_PBLOCK")) {  // experiment knob
    const int v = atoi(e);
    if (v == 64 || v == 128 || v == 256) block = v;
  }
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
    // optional (SQ_L2_PERSIST_MB): keep the bin directory, the one structure every probe row reads at a
    // random place, in the persisting part of L2; everything else streams through the rest
    cudaStreamAttrValue av{};
    size_t win = size_t(idx->dir_bytes);
    if (win > s->ctx->l2_window_max) win = s->ctx->l2_window_max;
    av.accessPolicyWindow.base_ptr = idx->d_dir;
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
