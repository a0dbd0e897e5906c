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

namespace sq {

// look-back word: [63:62] status, [61:0] value
constexpr uint64_t kFlagAgg = 1ull << 62;  // CTA aggregate available
constexpr uint64_t kFlagInc = 2ull << 62;  // inclusive prefix available
constexpr uint64_t kValMask = (1ull << 62) - 1;

struct Cand {
  uint32_t lo;  // first candidate (absolute position in the sorted arrays)
  uint32_t nc;  // number of candidates
};

// Both searches are written as a few rounds of INDEPENDENT loads (memory-level parallelism per
// thread) instead of dependent binary-search steps: every HBM/L2 round trip a probe row waits for
// is one of (1) the directory entry, (2) the four sampled starts that cover its bin, (3) the five
// speculative gallop points over runmax; the refinements that follow hit lines already in L1.
__device__ __forceinline__ Cand find_candidates(const IndexView& iv, uint32_t id, int32_t qs, int32_t qe) {
  Cand c{0u, 0u};
  if (id == kNoKey) return c;  // key hash absent from the build side: no rows (interval_join.rs:965)
  const SegMeta m = iv.meta[id];

  // ---- hi = first j in [sb, se) with start[j] > qe : bin directory, then a search inside the bin
  uint32_t a, len;
  if (qe < m.min_start) {
    a = m.sb; len = 0;
  } else {
    const uint32_t off = uint32_t(qe) - uint32_t(m.min_start);
    const uint32_t b = m.shift >= 32 ? 0u : (off >> m.shift);
    if (b >= m.nbins) {
      a = m.se; len = 0;
    } else {
      a = __ldg(iv.dir + m.dir_base + b);
      len = __ldg(iv.dir + m.dir_base + b + 1) - a;
    }
  }
  while (len > 32) {  // crowded bin (skewed data): narrow it the classic way first
    const uint32_t half = len >> 1;
    if (__ldg(iv.start + a + half) <= qe) { a += half + 1; len -= half + 1; } else len = half;
  }
  if (len) {
    // last element of each group of 8 (clamped): 4 independent loads that touch every sector of the bin
    const int32_t* sp = iv.start + a;
    int32_t s0 = __ldg(sp + min(7u, len - 1)), s1 = __ldg(sp + min(15u, len - 1));
    int32_t s2 = __ldg(sp + min(23u, len - 1)), s3 = __ldg(sp + min(31u, len - 1));
    uint32_t g = 0;  // number of whole groups that are <= qe
    g += (7u < len && s0 <= qe);
    g += (15u < len && s1 <= qe);
    g += (23u < len && s2 <= qe);
    g += (31u < len && s3 <= qe);
    // starts are sorted, so whole groups <= qe form a prefix: the answer is inside group g
    const uint32_t gb = 8u * g;
    uint32_t cnt = 0;
    if (gb < len) {
      const uint32_t gl = min(8u, len - gb);
      int32_t v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldg(sp + gb + min(uint32_t(k), gl - 1));
#pragma unroll
      for (int k = 0; k < 8; ++k) cnt += (uint32_t(k) < gl && v[k] <= qe);
    }
    a += gb + cnt;
  }
  const uint32_t hi = a;
  if (hi == m.sb) return c;

  // ---- lo = first j in [sb, hi) with runmax[j] >= qs (runmax non-decreasing): speculative gallop
  const uint32_t span = hi - m.sb;  // rows available below hi
  const int2* rp = iv.re + hi;      // rp[-d] = row hi-d
  const int32_t r1 = __ldg(&rp[-int(min(1u, span))].x);
  const int32_t r2 = __ldg(&rp[-int(min(2u, span))].x);
  const int32_t r4 = __ldg(&rp[-int(min(4u, span))].x);
  const int32_t r8 = __ldg(&rp[-int(min(8u, span))].x);
  const int32_t r16 = __ldg(&rp[-int(min(16u, span))].x);
  if (r1 < qs) return c;  // nothing reaches qs
  // [left, right]: right qualifies, everything below left does not
  uint32_t right, left;
  if (r2 < qs) { right = hi - 1; left = hi - 1; }
  else if (r4 < qs) { right = hi - min(2u, span); left = hi - min(4u, span) + 1; }
  else if (r8 < qs) { right = hi - min(4u, span); left = hi - min(8u, span) + 1; }
  else if (r16 < qs) { right = hi - min(8u, span); left = hi - min(16u, span) + 1; }
  else {
    right = hi - min(16u, span);
    left = m.sb;
    uint32_t step = 16;
    while (right - left >= step) {
      const uint32_t p = right - step;
      if (__ldg(&iv.re[p].x) >= qs) { right = p; step <<= 1; } else { left = p + 1; break; }
    }
  }
  if (left > right) left = right;  // clamped gallop points may coincide
  // first j in [left, right] with runmax[j] >= qs; right qualifies
  len = right - left;
  a = left;
  while (len) {
    const uint32_t half = len >> 1;
    if (__ldg(&iv.re[a + half].x) < qs) { a += half + 1; len -= half + 1; } else len = half;
  }
  c.lo = a;
  c.nc = hi - a;
  return c;
}

__device__ __forceinline__ uint32_t warp_incl_sum(uint32_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
    if (int(threadIdx.x & 31) >= d) v += o;
  }
  return v;
}

// Owner row (lane) of flattened candidate t: number of lanes whose inclusive prefix <= t.
__device__ __forceinline__ int owner_of(uint32_t incl, uint32_t t) {
  int p = 0;
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const uint32_t v = __shfl_sync(0xffffffffu, incl, p + s - 1);
    if (v <= t) p += s;
  }
  return p;
}

// bits [x, y) of a 32-bit mask, 0 <= x < y <= 32
__device__ __forceinline__ uint32_t bit_range(uint32_t x, uint32_t y) {
  const uint32_t hi = y >= 32 ? 0xffffffffu : ((1u << y) - 1u);
  return hi & ~((1u << x) - 1u);
}

__device__ __forceinline__ uint32_t low_bits(uint32_t nbits) {  // nbits in [0, 32]
  return nbits >= 32 ? 0xffffffffu : ((1u << nbits) - 1u);
}

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
  __shared__ uint32_t s_wtot[kWarpsPerBlock];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t i = blockIdx.x * kProbeBlock + threadIdx.x;

  Cand c{0u, 0u};
  int32_t qs = 0;
  if (i < n) {
    qs = q_start[i];
    const uint32_t id = ht_lookup(iv.ht_keys, iv.ht_ids, iv.ht_mask, iv.sentinel_id, q_key[i]);
    c = find_candidates(iv, id, qs, q_end[i]);
  }

  // flattened walk over the warp's candidates: lane t looks at candidate t of the concatenation
  const uint32_t incl = warp_incl_sum(c.nc);
  const uint32_t excl = incl - c.nc;
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  const uint32_t jbase = c.lo - excl;  // row of candidate t (owned by this lane) = jbase + t
  uint32_t cnt = 0, mask = 0;
  for (uint32_t t0 = 0; t0 < total; t0 += 32) {
    const uint32_t t = t0 + lane;
    const int p = owner_of(incl, t);
    const uint32_t j = __shfl_sync(0xffffffffu, jbase, p) + t;
    const int32_t pqs = __shfl_sync(0xffffffffu, qs, p);
    const bool hit = (t < total) && (__ldg(&iv.re[j].y) >= pqs);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    // my own row's share of this chunk: flattened positions [excl, incl) clipped to the chunk
    const uint32_t a = max(excl, t0), b = min(incl, t0 + 32);
    if (a < b) {
      const uint32_t bits = (m >> (a - t0)) & low_bits(b - a);
      cnt += __popc(bits);
      if (c.nc <= 32) mask |= bits << (a - excl);
    }
  }
  if (i < n) {
    lo_out[i] = c.lo;
    nc_out[i] = c.nc;
    mask_out[i] = mask;
    cnt_out[i] = cnt;
  }
  uint32_t wsum = cnt;  // <= 32 * n_rows_build fits u32 only per lane; sum per CTA in u64 below
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, d);
  // a warp's total can exceed 2^32 only if 32 rows each hit > 2^27 build rows; keep u64 across warps
  if (lane == 0) s_wtot[warp] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tot = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) tot += s_wtot[w];
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
// K3: write.  Same flattened walk, driven by the saved candidate ranges; a probe row with <= 32
// candidates takes its hits from the saved bitmask (the index arrays end/runmax are not touched
// again), wider rows re-test end[] against the probe start.  Output offsets of consecutive probe
// rows are contiguous, so a warp's hits are one contiguous run: position = warp base + ballot
// rank, and the stores of left_idx / right_idx are fully coalesced.
// ---------------------------------------------------------------------------------------------
template <bool WRITE_RIGHT>
__global__ void __launch_bounds__(kProbeBlock)
k_probe_write(IndexView iv, const int32_t* __restrict__ q_start, uint32_t n, const uint32_t* __restrict__ lo_in,
              const uint32_t* __restrict__ nc_in, const uint32_t* __restrict__ mask_in,
              const uint32_t* __restrict__ cnt_in, const unsigned long long* __restrict__ tile_base,
              unsigned long long* result, uint32_t* __restrict__ left_out, uint32_t* __restrict__ right_out,
              uint64_t capacity) {
  __shared__ uint32_t s_wtot[kWarpsPerBlock];
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
  uint32_t wsum = cnt;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, d);
  if (lane == 0) s_wtot[warp] = wsum;
  __syncthreads();
  if (wsum == 0) return;  // warp-uniform
  uint64_t base = tile_base[blockIdx.x];
  for (int w = 0; w < warp; ++w) base += s_wtot[w];

  const unsigned wide = __ballot_sync(0xffffffffu, nc > 32);
  int32_t qs = 0;
  if (wide && i < n) qs = q_start[i];
  const uint32_t incl = warp_incl_sum(nc);
  const uint32_t excl = incl - nc;
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  for (uint32_t t0 = 0; t0 < total; t0 += 32) {
    const uint32_t t = t0 + lane;
    const int p = owner_of(incl, t);
    const uint32_t k = t - __shfl_sync(0xffffffffu, excl, p);  // candidate k of row p
    const uint32_t j = __shfl_sync(0xffffffffu, lo, p) + k;
    const uint32_t bits = __shfl_sync(0xffffffffu, mask, p);
    bool hit = (t < total) && ((bits >> (k & 31)) & 1u);
    if (wide) {  // warp-uniform
      const int32_t pqs = __shfl_sync(0xffffffffu, qs, p);
      if ((wide >> p) & 1u) hit = (t < total) && (__ldg(&iv.re[j].y) >= pqs);
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const uint64_t pos = base + __popc(m & ((1u << lane) - 1u));
      left_out[pos] = __ldg(iv.row + j);
      if (WRITE_RIGHT) right_out[pos] = tile_first + p;
    }
    base += __popc(m);
  }
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
