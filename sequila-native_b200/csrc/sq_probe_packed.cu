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
#include "sq_probe_common.cuh"

namespace sq {

constexpr int kPBlockDefault = 128;      // threads per CTA = probe rows per CTA
constexpr uint32_t kSlots = 32;          // stash slots per probe row
constexpr uint32_t kStride = kSlots + 1; // padded row stride of the stash (bank spread)

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

// the two row slots this lane holds of a line: lane 0 = {header, row 0}, lanes 1..7 = {row 2k-1, row 2k}
struct Slots {
  uint32_t a_lo, a_id, b_lo, b_id;
};
__device__ __forceinline__ Slots slots_of(const uint4& d, int sub) {
  Slots s;
  s.a_lo = sub ? d.x : d.z;
  s.a_id = sub ? d.y : d.w;
  s.b_lo = d.z;
  s.b_id = sub ? d.w : kEmptyRow;
  return s;
}
__device__ __forceinline__ bool row_hits(uint32_t lo_word, uint32_t id, int32_t base, int32_t qs, int32_t qe) {
  const int32_t st = base + int32_t(lo_word & 0xFFFFu);
  const int32_t en = st + int32_t(lo_word >> 16);
  return id != kEmptyRow && st <= qe && en >= qs;
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
  const int g = lane >> 3, sub = lane & 7;  // lane group (one probe row at a time) and lane in group
  const int g0 = g * 8;

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

  // ---- phase 2: 4 probe rows per step, 8 lanes each; rounds of up to 8 independent line requests -----
  // Every round, the rows that still walk back are compacted (rank among walking rows -> step, group),
  // each fetches ONE more line, all steps of the warp at once: a warp waits for as many memory round
  // trips as its longest walk and executes only as many steps as it has (row, line) pairs.
  uint32_t* stash = s_stash + (EMIT ? warp * 32 * kStride : 0);
  bool walking = sl.act;
  uint32_t ln = sl.line;  // next line of my row
  uint32_t cnt = 0;       // hits of my row so far
  for (;;) {
    const unsigned A = __ballot_sync(0xffffffffu, walking);
    if (A == 0) break;
    const int n_walk = __popc(A);
    const int r = __popc(A & ((1u << lane) - 1u));  // my rank: served in step r >> 2 by group r & 3
    __syncwarp();
    if (walking) s_inv[warp][(r & 3) * 8 + (r >> 2)] = uint8_t(lane);
    __syncwarp();
    const unsigned long long srcs = *reinterpret_cast<const unsigned long long*>(&s_inv[warp][g0]);  // my group's 8 rows
    const int n_steps = (n_walk + 3) >> 2;
    uint4 v[8];
#pragma unroll
    for (int st = 0; st < 8; ++st) {
      const int p = int(srcs >> (8 * st)) & 31;
      const uint32_t lnp = __shfl_sync(0xffffffffu, ln, p);
      v[st] = (4 * st + g < n_walk) ? __ldg(iv.lines + size_t(lnp) * 8 + sub) : make_uint4(0u, 0u, 0u, kEmptyRow);
    }
    bool cont = false;
#pragma unroll
    for (int st = 0; st < 8; ++st) {
      if (st >= n_steps) break;  // warp-uniform
      const int p = int(srcs >> (8 * st)) & 31;
      const bool valid = 4 * st + g < n_walk;
      const int32_t qs = __shfl_sync(0xffffffffu, my_qs, p);
      const int32_t qe = __shfl_sync(0xffffffffu, my_qe, p);
      const uint32_t c0 = __shfl_sync(0xffffffffu, cnt, p);
      const uint4 d = v[st];
      const int32_t base = int32_t(__shfl_sync(0xffffffffu, d.x, g0));
      const int32_t exmax = int32_t(__shfl_sync(0xffffffffu, d.y, g0));
      const Slots s = slots_of(d, sub);
      const bool ha = valid && row_hits(s.a_lo, s.a_id, base, qs, qe);
      const bool hb = valid && row_hits(s.b_lo, s.b_id, base, qs, qe);
      const uint32_t ma = (__ballot_sync(0xffffffffu, ha) >> g0) & 0xFFu;
      const uint32_t mb = (__ballot_sync(0xffffffffu, hb) >> g0) & 0xFFu;
      if (EMIT) {
        const uint32_t below = (1u << sub) - 1u;
        const uint32_t pa = c0 + __popc(ma & below);
        const uint32_t pb = c0 + __popc(ma) + __popc(mb & below);
        if (ha && pa < kSlots) stash[p * kStride + pa] = s.a_id;
        if (hb && pb < kSlots) stash[p * kStride + pb] = s.b_id;
      }
      // hand the new count and "an earlier row still reaches qs" back to the owner lane
      const uint32_t c1 = __shfl_sync(0xffffffffu, c0 + __popc(ma) + __popc(mb), (r & 3) * 8);
      const unsigned reach = __ballot_sync(0xffffffffu, valid && exmax >= qs);
      if (walking && (r >> 2) == st) {
        cnt = c1;
        cont = (reach >> ((r & 3) * 8)) & 1u;
      }
    }
    walking = walking && cont && ln > sl.first;
    ln -= 1;
  }
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
    if (lane == 0) atomicExch(chain_state + bid, (bid == 0 ? kFlagInc : kFlagAgg) | agg);
    unsigned long long excl = 0;
    if (bid > 0) {
      int64_t look = int64_t(bid) - 1;
      for (;;) {
        const int64_t k = look - lane;
        unsigned long long x = kFlagInc;
        if (k >= 0) {
          do { x = *reinterpret_cast<volatile unsigned long long*>(chain_state + k); } while ((x >> 62) == 0);
        }
        const unsigned inc_mask = __ballot_sync(0xffffffffu, (x >> 62) == 2);
        const int first_inc = inc_mask ? (__ffs(inc_mask) - 1) : 32;
        unsigned long long y = (lane <= first_inc) ? (x & kValMask) : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) y += __shfl_xor_sync(0xffffffffu, y, d);
        excl += y;
        if (inc_mask) break;
        look -= 32;
      }
      if (lane == 0) atomicExch(chain_state + bid, kFlagInc | (excl + agg));
    }
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

  // ---- phase 4a: rows with <= 32 hits, one flattened list per warp out of the stash ------------------
  {
    const Flat f = flat_setup(cnt <= kSlots ? cnt : 0u, lane, s_inv[warp]);  // also orders the stash writes
    const uint32_t r_coff = __shfl_sync(0xffffffffu, coff, f.r_src);
    for (uint32_t t0 = 0; t0 < f.total; t0 += 32) {
      const uint32_t t = t0 + lane;
      const int r = flat_rank(f, t0, lane);
      const uint32_t k = t - __shfl_sync(0xffffffffu, f.r_excl, r);
      const uint32_t off = __shfl_sync(0xffffffffu, r_coff, r);
      const int src = __shfl_sync(0xffffffffu, f.r_src, r);
      if (t < f.total) {
        const uint32_t pos = off + k;
        lout[pos] = stash[src * kStride + k];
        if (WRITE_RIGHT) rout[pos] = tile_first + src;
      }
    }
  }
  // ---- phase 4b: rows with > 32 hits, the whole warp re-walks the row, four lines per step -----------
  unsigned big = __ballot_sync(0xffffffffu, cnt > kSlots);
  while (big) {
    const int p = __ffs(big) - 1;
    big &= big - 1;
    const int32_t qs = __shfl_sync(0xffffffffu, my_qs, p);
    const int32_t qe = __shfl_sync(0xffffffffu, my_qe, p);
    uint32_t ln = __shfl_sync(0xffffffffu, sl.line, p);
    const uint32_t first = __shfl_sync(0xffffffffu, sl.first, p);
    uint32_t run = __shfl_sync(0xffffffffu, coff, p);
    for (;;) {
      const bool has = ln - first >= uint32_t(g);  // group g takes line ln - g
      const uint4 d = has ? __ldg(iv.lines + size_t(ln - g) * 8 + sub) : make_uint4(0u, uint32_t(INT32_MIN), 0u, kEmptyRow);
      const int32_t lbase = int32_t(__shfl_sync(0xffffffffu, d.x, g0));
      const int32_t exmax = int32_t(__shfl_sync(0xffffffffu, d.y, g0));
      const bool more = has && exmax >= qs && (ln - g) > first;  // this line sends the walk one line further
      const unsigned mm = __ballot_sync(0xffffffffu, more);
      const unsigned m4 = (mm & 1u) | ((mm >> 7) & 2u) | ((mm >> 14) & 4u) | ((mm >> 21) & 8u);
      const bool live = has && ((m4 & ((1u << g) - 1u)) == ((1u << g) - 1u));  // every later line continued
      const Slots s = slots_of(d, sub);
      const bool ha = live && row_hits(s.a_lo, s.a_id, lbase, qs, qe);
      const bool hb = live && row_hits(s.b_lo, s.b_id, lbase, qs, qe);
      const unsigned ma = __ballot_sync(0xffffffffu, ha), mb = __ballot_sync(0xffffffffu, hb);
      const unsigned below = (1u << lane) - 1u;
      if (ha) {
        const uint32_t pos = run + __popc(ma & below);
        lout[pos] = s.a_id;
        if (WRITE_RIGHT) rout[pos] = tile_first + p;
      }
      if (hb) {
        const uint32_t pos = run + __popc(ma) + __popc(mb & below);
        lout[pos] = s.b_id;
        if (WRITE_RIGHT) rout[pos] = tile_first + p;
      }
      run += __popc(ma) + __popc(mb);
      if (m4 != 0xFu) break;
      ln -= 4;
    }
  }
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
