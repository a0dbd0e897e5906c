// Probe + emit of the `Cuda` interval join: replaces the per-row loop of process_probe_batch
// (reference interval_join.rs:1586-1618: hash_map.get -> coitrees query -> pos_vect / rle_right
// -> index_right) with three kernels over a tile of probe rows (count -> exclusive scan -> write):
//
//   k_probe_count  one thread per probe row: key hash -> key id -> segment meta; the upper bound
//                  hi of the candidate range comes from the segment's bin directory plus one round
//                  of sampled loads inside the bin; the lower bound lo from a speculative gallop
//                  backwards over the running max end.  Candidates are the contiguous rows [lo,hi).
//                  Each warp then walks its 32 rows' candidate ranges as ONE flattened list
//                  (coalesced reads, no lane idles on a short list); hits = end >= probe start.
//                  Saved per row: lo, nc, hit bitmask (nc <= 32), hit count; per CTA: pair total.
//   k_tile_scan    chained scan (decoupled look-back) of the CTA totals -> output offset per tile.
//   k_probe_write  the same flattened walk driven by the saved state: hits come from the bitmask,
//                  only row[] is read; a warp's hits are one contiguous run of the output
//                  (position = warp base + ballot rank) so left_idx / right_idx stores coalesce.
//
// count and write are separate kernels on purpose: a single fused kernel had every CTA wait at a
// barrier for its look-back (21-24 % of warp stalls in ncu, profiles/r01_v2_*) while holding its
// registers, and its write stage re-read index sectors that had left L2 in the meantime.
// Integer/byte work bounded by HBM (or by L1/L2 when the index is cache-resident); tensor cores
// do not apply.
#include <cstdlib>

#include "sq_internal.cuh"
#include "sq_probe_common.cuh"
#include "sq_soa_common.cuh"

namespace sq {

// ---------------------------------------------------------------------------------------------
// K1: search + count.  Per probe row it leaves in HBM everything the write kernel needs so that
// nothing is searched twice: lo, nc, the hit bitmask of its candidates when nc <= 32 (bit k <=>
// row lo+k is a hit) and the hit count (= rle_right, interval_join.rs:1604).  Per CTA: pair total.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kProbeBlock)
k_probe_count(IndexView iv, const uint64_t* __restrict__ q_key, const int32_t* __restrict__ q_start,
              const int32_t* __restrict__ q_end, uint32_t n, uint32_t* __restrict__ lo_out,
              uint32_t* __restrict__ nc_out, uint32_t* __restrict__ mask_out, uint32_t* __restrict__ cnt_out,
              unsigned long long* __restrict__ tile_total) {
  __shared__ unsigned long long s_tot[kWarpsPerBlock];
  __shared__ uint8_t s_inv[kWarpsPerBlock][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t i = blockIdx.x * kProbeBlock + threadIdx.x;

  Cand c{0u, 0u};
  int32_t qs = 0;
  if (i < n) {
    qs = q_start[i];
    const uint32_t id = ht_lookup(iv.ht_keys, iv.ht_ids, iv.ht_mask, iv.sentinel_id, q_key[i]);
    c = find_candidates(iv, id, qs, q_end[i]);
  }
  uint32_t cnt = 0, mask = 0;

  // ---- small rows, flattened ----------------------------------------------------------------------
  {
    const Flat f = flat_setup(c.nc <= kSmallMax ? c.nc : 0u, lane, s_inv[warp]);
    const uint32_t r_jbase = __shfl_sync(0xffffffffu, c.lo, f.r_src) - f.r_excl;  // candidate t -> row r_jbase + t
    const int32_t r_qs = __shfl_sync(0xffffffffu, qs, f.r_src);
    uint32_t r_mask = 0;  // hit mask of the row with my rank
    for (uint32_t t0 = 0; t0 < f.total; t0 += 32) {
      const uint32_t t = t0 + lane;
      const int r = flat_rank(f, t0, lane);
      const uint32_t j = __shfl_sync(0xffffffffu, r_jbase, r) + t;
      const int32_t pqs = __shfl_sync(0xffffffffu, r_qs, r);
      const bool hit = (t < f.total) && (__ldg(iv.end + j) >= pqs);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      // the share of my rank's row in this chunk: positions [r_excl, r_incl) clipped to the chunk
      const uint32_t a = max(f.r_excl, t0), b = min(f.r_incl, t0 + 32);
      if (f.r_incl != 0xffffffffu && a < b) r_mask |= ((m >> (a - t0)) & low_bits(b - a)) << (a - f.r_excl);
    }
    // hand the masks back from rank lanes to owner lanes
    const unsigned nz = __ballot_sync(0xffffffffu, c.nc != 0 && c.nc <= kSmallMax);
    const uint32_t got = __shfl_sync(0xffffffffu, r_mask, __popc(nz & ((1u << lane) - 1u)));
    if ((nz >> lane) & 1u) { mask = got; cnt = __popc(got); }
  }
  // ---- big rows, wide --------------------------------------------------------------------------------
  unsigned big = __ballot_sync(0xffffffffu, c.nc > kSmallMax);
  while (big) {
    const int p = __ffs(big) - 1;
    big &= big - 1;
    const uint32_t nc_p = __shfl_sync(0xffffffffu, c.nc, p);
    const uint32_t lo_p = __shfl_sync(0xffffffffu, c.lo, p);
    const int32_t qs_p = __shfl_sync(0xffffffffu, qs, p);
    uint32_t tot = 0;
    if (nc_p > kSkipMin) {  // a long range (a very long build interval keeps runmax high behind it): only blocks with a hit
      for_hit_blocks(iv, lo_p, lo_p + nc_p, qs_p, lane, [&](uint32_t j0) {
        const uint32_t j = j0 + lane;
        const bool hit = j >= lo_p && j < lo_p + nc_p && (__ldg(iv.end + j) >= qs_p);
        tot += __popc(__ballot_sync(0xffffffffu, hit));
      });
    } else {
      for (uint32_t k0 = 0; k0 < nc_p; k0 += 32) {
        const uint32_t k = k0 + lane;
        const bool hit = k < nc_p && (__ldg(iv.end + lo_p + k) >= qs_p);
        tot += __popc(__ballot_sync(0xffffffffu, hit));
      }
    }
    if (lane == p) cnt = tot;
  }

  if (i < n) {
    lo_out[i] = c.lo;
    nc_out[i] = c.nc;
    mask_out[i] = mask;
    cnt_out[i] = cnt;
  }
  unsigned long long wsum = cnt;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, d);
  if (lane == 0) s_tot[warp] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tot = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) tot += s_tot[w];
    tile_total[blockIdx.x] = tot;
  }
}

// ---------------------------------------------------------------------------------------------
// K2: exclusive scan of the per-CTA pair totals (in place) -> output offset of every probe tile,
// and the grand total.  Chained scan with decoupled look-back across CTAs of 1024 tiles each.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_tile_scan(unsigned long long* __restrict__ tile_total, uint32_t n_tiles, unsigned long long* chain_state,
            unsigned int* ticket, unsigned long long* result) {
  __shared__ unsigned long long s_w[32];
  __shared__ unsigned long long s_excl;
  __shared__ uint32_t s_bid;
  if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t bid = s_bid;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t i = bid * 1024 + threadIdx.x;
  const unsigned long long v = i < n_tiles ? tile_total[i] : 0ull;
  unsigned long long inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = s_w[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += o;
    }
    s_w[lane] = w;  // inclusive over warps
    const unsigned long long agg = __shfl_sync(0xffffffffu, w, 31);
    if (lane == 0) atomicExch(chain_state + bid, (bid == 0 ? kFlagInc : kFlagAgg) | agg);
    unsigned long long excl_base = 0;
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
        excl_base += y;
        if (inc_mask) break;
        look -= 32;
      }
      if (lane == 0) atomicExch(chain_state + bid, kFlagInc | (excl_base + agg));
    }
    if (lane == 0) {
      s_excl = excl_base;
      if (bid == gridDim.x - 1) result[0] = excl_base + agg;
    }
  }
  __syncthreads();
  if (i < n_tiles) tile_total[i] = s_excl + (warp ? s_w[warp - 1] : 0ull) + (inc - v);
}

// ---------------------------------------------------------------------------------------------
// K3: write.  The same walks, driven by the saved state; a small row takes its hits from the saved
// bitmask (end/runmax are not touched again), big rows re-test end[] against the probe start.
// Output offsets of consecutive probe rows are contiguous: row p's pairs start at warp base +
// exclusive prefix of the hit counts, and hit k of a row lands at that offset + popc(mask below
// k), so the stores of left_idx / right_idx fill dense runs of the output.
// ---------------------------------------------------------------------------------------------
template <bool WRITE_RIGHT>
__global__ void __launch_bounds__(kProbeBlock)
k_probe_write(IndexView iv, const int32_t* __restrict__ q_start, uint32_t n, const uint32_t* __restrict__ lo_in,
              const uint32_t* __restrict__ nc_in, const uint32_t* __restrict__ mask_in,
              const uint32_t* __restrict__ cnt_in, const unsigned long long* __restrict__ tile_base,
              unsigned long long* result, uint32_t* __restrict__ left_out, uint32_t* __restrict__ right_out,
              uint64_t capacity) {
  __shared__ unsigned long long s_wtot[kWarpsPerBlock];
  __shared__ uint8_t s_inv[kWarpsPerBlock][32];
  if (result[0] > capacity) {  // grid-uniform: the caller's buffer is too small, report it
    if (blockIdx.x == 0 && threadIdx.x == 0) result[1] = 1;
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t i = blockIdx.x * kProbeBlock + threadIdx.x;
  const uint32_t tile_first = blockIdx.x * kProbeBlock + warp * 32;

  uint32_t lo = 0, nc = 0, mask = 0, cnt = 0;
  if (i < n) {
    lo = lo_in[i];
    nc = nc_in[i];
    mask = mask_in[i];
    cnt = cnt_in[i];
  }
  const uint32_t cincl = warp_incl_sum(cnt);  // a warp emits < 2^32 pairs unless rows hit > 2^27 builds each
  const uint32_t wtot = __shfl_sync(0xffffffffu, cincl, 31);
  if (lane == 0) s_wtot[warp] = wtot;
  __syncthreads();
  if (wtot == 0) return;  // warp-uniform
  uint64_t base = tile_base[blockIdx.x];
  for (int w = 0; w < warp; ++w) base += s_wtot[w];
  uint32_t* __restrict__ lout = left_out + base;
  uint32_t* __restrict__ rout = WRITE_RIGHT ? right_out + base : nullptr;
  const uint32_t coff = cincl - cnt;  // offset of my row's first pair inside the warp's run

  // ---- small rows, flattened over the candidates up to each row's last hit --------------------------
  {
    const Flat f = flat_setup(mask ? 32u - uint32_t(__clz(mask)) : 0u, lane, s_inv[warp]);
    const uint32_t r_jbase = __shfl_sync(0xffffffffu, lo, f.r_src) - f.r_excl;
    const uint32_t r_mask = __shfl_sync(0xffffffffu, mask, f.r_src);
    const uint32_t r_coff = __shfl_sync(0xffffffffu, coff, f.r_src);
    for (uint32_t t0 = 0; t0 < f.total; t0 += 32) {
      const uint32_t t = t0 + lane;
      const int r = flat_rank(f, t0, lane);
      const uint32_t k = t - __shfl_sync(0xffffffffu, f.r_excl, r);  // candidate k of that row
      const uint32_t bits = __shfl_sync(0xffffffffu, r_mask, r);
      const uint32_t off = __shfl_sync(0xffffffffu, r_coff, r);
      const uint32_t j = __shfl_sync(0xffffffffu, r_jbase, r) + t;  // 32-bit wrap-around is intended
      const int src = __shfl_sync(0xffffffffu, f.r_src, r);
      if (t < f.total && ((bits >> (k & 31)) & 1u)) {
        const uint32_t pos = off + __popc(bits & low_bits(k & 31));
        lout[pos] = __ldg(iv.row + j);
        if (WRITE_RIGHT) rout[pos] = tile_first + src;
      }
    }
  }
  // ---- big rows, wide: re-test end[] ---------------------------------------------------------------------
  unsigned big = __ballot_sync(0xffffffffu, nc > kSmallMax);
  if (big == 0) return;
  int32_t qs = 0;
  if (i < n) qs = q_start[i];
  while (big) {
    const int p = __ffs(big) - 1;
    big &= big - 1;
    const uint32_t nc_p = __shfl_sync(0xffffffffu, nc, p);
    const uint32_t lo_p = __shfl_sync(0xffffffffu, lo, p);
    const int32_t qs_p = __shfl_sync(0xffffffffu, qs, p);
    uint32_t run = __shfl_sync(0xffffffffu, coff, p);
    auto block = [&](uint32_t j0) {  // rows j0 + lane of [lo_p, lo_p + nc_p)
      const uint32_t j = j0 + lane;
      const bool hit = j >= lo_p && j - lo_p < nc_p && (__ldg(iv.end + j) >= qs_p);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const uint32_t pos = run + __popc(m & ((1u << lane) - 1u));
        lout[pos] = __ldg(iv.row + j);
        if (WRITE_RIGHT) rout[pos] = tile_first + p;
      }
      run += __popc(m);
    };
    if (nc_p > kSkipMin) for_hit_blocks(iv, lo_p, lo_p + nc_p, qs_p, lane, block);
    else for (uint32_t k0 = 0; k0 < nc_p; k0 += 32) block(lo_p + k0);
  }
}


// ---------------------------------------------------------------------------------------------
// Algorithm::CoitreesNearest on the flat index (reference interval_join.rs:794-812, 909-956, 972-990,
// 1593-1602): ONE output row per probe row.  left = a build row overlapping the probe row if one
// exists (the reference reports the first its tree traversal visits, "an arbitrary one"; here: the
// overlapping row with the greatest sorted position), else the row `nearest()` picks, else NULL when
// the key hash never occurred on the build side.  `nearest()` looks at exactly two candidates of the
// segment ordered by (start, end, row): the last one with start < qe and the first one with
// start >= qe; the earlier candidate wins ties.  Rows with equal start are adjacent here (the segment
// is sorted by start), so each candidate is a min / max over one short run.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t first_start_greater(const IndexView& iv, const SegMeta& m, int32_t q) {
  if (q < m.min_start) return m.sb;
  const uint32_t off = uint32_t(q) - uint32_t(m.min_start);
  const uint32_t b = m.shift >= 32 ? 0u : (off >> m.shift);
  if (b >= m.nbins) return m.se;
  uint32_t a = __ldg(iv.dir + m.dir_base + b);
  uint32_t len = __ldg(iv.dir + m.dir_base + b + 1) - a;
  while (len) {
    const uint32_t half = len >> 1;
    if (__ldg(iv.start + a + half) <= q) { a += half + 1; len -= half + 1; } else len = half;
  }
  return a;
}

__global__ void __launch_bounds__(256)
k_probe_nearest(IndexView iv, const uint64_t* __restrict__ q_key, const int32_t* __restrict__ q_start,
                const int32_t* __restrict__ q_end, uint32_t n, uint32_t* __restrict__ left_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t qs = q_start[i], qe = q_end[i];
  const uint32_t id = ht_lookup(iv.ht_keys, iv.ht_ids, iv.ht_mask, iv.sentinel_id, q_key[i]);
  if (id == kNoKey) { left_out[i] = kEmptyRow; return; }  // NULL left side (interval_join.rs:1597-1598)
  const SegMeta m = iv.meta[id];
  // an overlap, if any: walk back from the last start <= qe while some earlier end still reaches qs
  const uint32_t hi = first_start_greater(iv, m, qe);
  if (hi > m.sb && __ldg(iv.runmax + hi - 1) >= qs) {  // some row with start <= qe reaches qs: the last such row
    const uint32_t j = last_hit_below(iv, m.sb, hi, qs);
    if (j < hi) { left_out[i] = __ldg(iv.row + j); return; }
  }
  // nearest(): `left` = first position with start >= qe in (start, end, row) order
  const uint32_t lb = qe == INT32_MIN ? m.sb : first_start_greater(iv, m, qe - 1);
  long long best_d = INT32_MAX;
  uint32_t best = kEmptyRow;
  auto consider = [&](int32_t st, int32_t en, uint32_t row) {
    long long d;
    if (qe < st) d = (long long)st - qe;
    else if (en < qs) d = (long long)qs - en;
    else d = 0;
    if (d < best_d) { best_d = d; best = row; }
  };
  auto first_of_run = [&](uint32_t j0) {  // min (end, row) over the run of rows with start == start[j0], j0 = its first row
    const int32_t st = __ldg(iv.start + j0);
    int32_t en = __ldg(iv.end + j0);
    uint32_t row = __ldg(iv.row + j0);
    for (uint32_t j = j0 + 1; j < m.se && __ldg(iv.start + j) == st; ++j) {
      const int32_t e2 = __ldg(iv.end + j);
      const uint32_t r2 = __ldg(iv.row + j);
      if (e2 < en || (e2 == en && r2 < row)) { en = e2; row = r2; }
    }
    consider(st, en, row);
  };
  if (lb > m.sb) {  // sorted[left - 1]: max (end, row) over the run ending at lb - 1
    const uint32_t j1 = lb - 1;
    const int32_t st = __ldg(iv.start + j1);
    int32_t en = __ldg(iv.end + j1);
    uint32_t row = __ldg(iv.row + j1);
    for (uint32_t j = j1; j > m.sb && __ldg(iv.start + j - 1) == st;) {
      --j;
      const int32_t e2 = __ldg(iv.end + j);
      const uint32_t r2 = __ldg(iv.row + j);
      if (e2 > en || (e2 == en && r2 > row)) { en = e2; row = r2; }
    }
    consider(st, en, row);
  } else {
    first_of_run(m.sb);  // left == 0: left.saturating_sub(1) is sorted[0] as well
  }
  if (lb < m.se) first_of_run(lb);  // sorted[left]
  left_out[i] = best;
}

int launch_nearest(sq_stream* s, const sq_index* idx, const uint64_t* d_key, const int32_t* d_start,
                   const int32_t* d_end, uint32_t n, uint32_t* d_left) {
  if (n == 0) return SQ_OK;
  k_probe_nearest<<<(n + 255) / 256, 256, 0, s->stream>>>(idx->view(), d_key, d_start, d_end, n, d_left);
  SQ_CUDA(s->err, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

// ---------------------------------------------------------------------------------------------
static int ensure_probe_state(sq_stream* s, uint32_t n, uint32_t n_tiles) {
  ErrorSlot& E = s->err;
  int rc;
  if ((rc = ensure(E, s->d_cnt, size_t(n) * 4, false))) return rc;
  if ((rc = ensure(E, s->d_state, size_t(n) * 12, false))) return rc;
  const size_t chain = (size_t(n_tiles) + 1023) / 1024;
  if ((rc = ensure(E, s->d_tile, (size_t(n_tiles) + chain) * 8, false))) return rc;
  if ((rc = ensure(E, s->d_scalar, 256, false))) return rc;
  return SQ_OK;
}

int launch_scan_u64(sq_stream* s, unsigned long long* d_vals, uint32_t n, unsigned long long* d_total) {
  ErrorSlot& E = s->err;
  const uint32_t n_chain = (n + 1023) / 1024;
  int rc;
  if ((rc = ensure(E, s->d_chain, size_t(n_chain) * 8 + 16, false))) return rc;
  auto* chain = static_cast<unsigned long long*>(s->d_chain.p);
  auto* ticket = reinterpret_cast<unsigned int*>(chain + n_chain);
  SQ_CUDA(E, cudaMemsetAsync(chain, 0, size_t(n_chain) * 8 + 16, s->stream));
  k_tile_scan<<<n_chain ? n_chain : 1, 1024, 0, s->stream>>>(d_vals, n, chain, ticket, d_total);
  SQ_CUDA(E, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

// K1 + K2 on the stream: per-row state, tile offsets and result[0] = n_pairs
int launch_count(sq_stream* s, const sq_index* idx, const uint64_t* d_key, const int32_t* d_start,
                 const int32_t* d_end, uint32_t n) {
  ErrorSlot& E = s->err;
  const uint32_t n_tiles = (n + kProbeBlock - 1) / kProbeBlock;
  const uint32_t n_chain = (n_tiles + 1023) / 1024;
  int rc;
  if ((rc = ensure_probe_state(s, n, n_tiles))) return rc;
  auto* tile = static_cast<unsigned long long*>(s->d_tile.p);
  auto* chain = tile + n_tiles;
  auto* result = static_cast<unsigned long long*>(s->d_scalar.p);  // [0] n_pairs [1] overflow [2] ticket
  auto* ticket = reinterpret_cast<unsigned int*>(result + 2);
  auto* st = static_cast<uint32_t*>(s->d_state.p);
  SQ_CUDA(E, cudaMemsetAsync(chain, 0, size_t(n_chain) * 8, s->stream));
  SQ_CUDA(E, cudaMemsetAsync(result, 0, 32, s->stream));
  k_probe_count<<<n_tiles, kProbeBlock, 0, s->stream>>>(idx->view(), d_key, d_start, d_end, n, st, st + n,
                                                        st + 2 * size_t(n), static_cast<uint32_t*>(s->d_cnt.p), tile);
  SQ_CUDA(E, cudaGetLastError());
  k_tile_scan<<<n_chain, 1024, 0, s->stream>>>(tile, n_tiles, chain, ticket, result);
  SQ_CUDA(E, cudaGetLastError());
  s->launches += 2;
  return SQ_OK;
}

// K3 on the stream; writes nothing and sets result[1] when n_pairs > capacity
int launch_write(sq_stream* s, const sq_index* idx, const int32_t* d_start, uint32_t n, uint32_t* d_left,
                 uint32_t* d_right, uint64_t capacity) {
  ErrorSlot& E = s->err;
  const uint32_t n_tiles = (n + kProbeBlock - 1) / kProbeBlock;
  auto* tile = static_cast<unsigned long long*>(s->d_tile.p);
  auto* result = static_cast<unsigned long long*>(s->d_scalar.p);
  auto* st = static_cast<uint32_t*>(s->d_state.p);
  const auto* cnt = static_cast<const uint32_t*>(s->d_cnt.p);
  if (d_right)
    k_probe_write<true><<<n_tiles, kProbeBlock, 0, s->stream>>>(idx->view(), d_start, n, st, st + n, st + 2 * size_t(n),
                                                               cnt, tile, result, d_left, d_right, capacity);
  else
    k_probe_write<false><<<n_tiles, kProbeBlock, 0, s->stream>>>(idx->view(), d_start, n, st, st + n, st + 2 * size_t(n),
                                                                cnt, tile, result, d_left, nullptr, capacity);
  SQ_CUDA(E, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

}  // namespace sq
