// Probe + emit of the `Cuda` interval join: replaces the per-row loop of process_probe_batch
// (reference interval_join.rs:1586-1618: hash_map.get -> coitrees query -> pos_vect / rle_right
// -> index_right) with ONE fused kernel over a tile of probe rows, k_probe_join:
//
//   search   one thread per probe row: key hash -> key id -> segment meta; the upper bound hi of
//            the candidate range comes from the segment's bin directory (one load) plus a short
//            search inside one cache line of start[]; the lower bound lo by galloping backwards
//            from hi over the running max end.  Candidates are the contiguous rows [lo, hi).
//   count    each warp walks the 32 rows' candidate ranges as one flattened list (coalesced
//            reads of re[], no lane idles on a short list); hits = end >= probe start.
//   scan     the CTA's pair total goes through a decoupled look-back, so every CTA learns the
//            output offset of its first pair inside the same pass (count -> exclusive scan ->
//            write without a second kernel or per-row offsets in HBM).
//   write    same flattened walk (re[] now in L1/L2); output offsets of consecutive probe rows
//            are contiguous, so a warp's hits are one contiguous run: position = warp base +
//            ballot rank.  Stores of left_idx / right_idx are fully coalesced.
//
// The write stage runs only if the tile's pairs fit the caller's capacity; otherwise the kernel
// has still produced the exact pair count and per-row counts and the caller re-runs it with a
// large enough buffer (two-phase protocol of the C ABI).  Integer/byte work bounded by HBM (or
// by L2 when the index fits there); tensor cores do not apply.
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

template <bool WRITE_RIGHT>
__global__ void __launch_bounds__(kProbeBlock)
k_probe_join(IndexView iv, const uint64_t* __restrict__ q_key, const int32_t* __restrict__ q_start,
             const int32_t* __restrict__ q_end, uint32_t n, uint32_t* __restrict__ cnt_out,
             unsigned long long* tile_state, unsigned int* ticket, unsigned long long* result,
             uint32_t* __restrict__ left_out, uint32_t* __restrict__ right_out, uint64_t capacity) {
  __shared__ uint64_t s_wtot[kWarpsPerBlock];
  __shared__ uint64_t s_base;
  __shared__ uint32_t s_bid;

  if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);  // CTAs take tiles in start order
  __syncthreads();
  const uint32_t bid = s_bid;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t i = bid * kProbeBlock + threadIdx.x;
  const uint32_t tile_first = bid * kProbeBlock + warp * 32;

  // ---- search -------------------------------------------------------------------------------
  Cand c{0u, 0u};
  int32_t qs = 0;
  if (i < n) {
    qs = q_start[i];
    const uint32_t id = ht_lookup(iv.ht_keys, iv.ht_ids, iv.ht_mask, iv.sentinel_id, q_key[i]);
    c = find_candidates(iv, id, qs, q_end[i]);
  }

  // ---- count: flattened walk over the warp's candidates ----------------------------------------
  const uint32_t incl = warp_incl_sum(c.nc);
  const uint32_t excl = incl - c.nc;
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  uint32_t cnt = 0, wcount = 0;
  for (uint32_t t0 = 0; t0 < total; t0 += 32) {
    const uint32_t t = t0 + lane;
    const int p = owner_of(incl, t);
    const uint32_t j = __shfl_sync(0xffffffffu, c.lo, p) + (t - __shfl_sync(0xffffffffu, excl, p));
    const int32_t pqs = __shfl_sync(0xffffffffu, qs, p);
    const bool hit = (t < total) && (__ldg(&iv.re[j].y) >= pqs);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    // my own row's share of this chunk: flattened positions [excl, incl) clipped to the chunk
    const uint32_t a = max(excl, t0), b = min(incl, t0 + 32);
    if (a < b) cnt += __popc(m & bit_range(a - t0, b - t0));
    wcount += __popc(m);
  }
  if (cnt_out && i < n) cnt_out[i] = cnt;

  // ---- scan: CTA total -> decoupled look-back -> exclusive base of this CTA ---------------------
  if (lane == 0) s_wtot[warp] = wcount;
  __syncthreads();
  uint64_t agg = 0;
#pragma unroll
  for (int w = 0; w < kWarpsPerBlock; ++w) agg += s_wtot[w];
  if (warp == 0) {
    if (lane == 0)
      atomicExch(tile_state + bid, (unsigned long long)((bid == 0 ? kFlagInc : kFlagAgg) | agg));
    uint64_t excl_base = 0;
    if (bid > 0) {
      int64_t look = int64_t(bid) - 1;
      for (;;) {
        const int64_t k = look - lane;
        uint64_t w = kFlagInc;  // tiles before 0 behave as a finished prefix of 0
        if (k >= 0) {
          do {
            w = *reinterpret_cast<volatile unsigned long long*>(tile_state + k);
          } while ((w >> 62) == 0);
        }
        const unsigned inc_mask = __ballot_sync(0xffffffffu, (w >> 62) == 2);
        const int first_inc = inc_mask ? (__ffs(inc_mask) - 1) : 32;
        uint64_t v = (lane <= first_inc) ? (w & kValMask) : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        excl_base += v;
        if (inc_mask) break;
        look -= 32;
      }
      if (lane == 0) atomicExch(tile_state + bid, (unsigned long long)(kFlagInc | (excl_base + agg)));
    }
    if (lane == 0) {
      s_base = excl_base;
      if (bid == gridDim.x - 1) result[0] = excl_base + agg;
    }
  }
  if (left_out == nullptr) return;  // count-only pass
  __syncthreads();

  // ---- write -----------------------------------------------------------------------------------
  const uint64_t cta_base = s_base;
  if (cta_base + agg > capacity) {  // CTA-uniform: the caller's buffer is too small, report it
    if (threadIdx.x == 0) result[1] = 1;
    return;
  }
  if (wcount == 0) return;  // warp-uniform
  uint64_t base = cta_base;
  for (int w = 0; w < warp; ++w) base += s_wtot[w];
  for (uint32_t t0 = 0; t0 < total; t0 += 32) {
    const uint32_t t = t0 + lane;
    const int p = owner_of(incl, t);
    const uint32_t j = __shfl_sync(0xffffffffu, c.lo, p) + (t - __shfl_sync(0xffffffffu, excl, p));
    const int32_t pqs = __shfl_sync(0xffffffffu, qs, p);
    const bool hit = (t < total) && (__ldg(&iv.re[j].y) >= pqs);
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
int launch_join(sq_stream* s, const sq_index* idx, const uint64_t* d_key, const int32_t* d_start,
                const int32_t* d_end, uint32_t n, uint32_t* d_left, uint32_t* d_right, uint64_t capacity) {
  ErrorSlot& E = s->err;
  const uint32_t n_tiles = (n + kProbeBlock - 1) / kProbeBlock;
  int rc;
  if ((rc = ensure(E, s->d_cnt, size_t(n) * 4, false))) return rc;
  if ((rc = ensure(E, s->d_tile, size_t(n_tiles) * 8, false))) return rc;
  if ((rc = ensure(E, s->d_scalar, 256, false))) return rc;
  auto* tile_state = static_cast<unsigned long long*>(s->d_tile.p);
  auto* result = static_cast<unsigned long long*>(s->d_scalar.p);  // [0] n_pairs [1] overflow [2] ticket
  auto* ticket = reinterpret_cast<unsigned int*>(result + 2);
  SQ_CUDA(E, cudaMemsetAsync(tile_state, 0, size_t(n_tiles) * 8, s->stream));
  SQ_CUDA(E, cudaMemsetAsync(result, 0, 32, s->stream));
  if (capacity == 0) d_left = nullptr;
  // Keep the bin directory (one random 4-byte read per probe row, ~n_rows/4 bytes) resident in the
  // L2 set-aside: the streaming index/probe/output traffic would otherwise evict it.
  if (s->l2_window_idx != idx && s->ctx->l2_persist_bytes && idx->dir_bytes) {
    cudaStreamAttrValue av{};
    size_t bytes = idx->dir_bytes;
    if (bytes > s->ctx->l2_window_max) bytes = s->ctx->l2_window_max;
    av.accessPolicyWindow.base_ptr = idx->d_dir;
    av.accessPolicyWindow.num_bytes = bytes;
    av.accessPolicyWindow.hitRatio = bytes <= s->ctx->l2_persist_bytes ? 1.0f : float(s->ctx->l2_persist_bytes) / float(bytes);
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(s->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
    s->l2_window_idx = idx;
  }
  auto* cnt = static_cast<uint32_t*>(s->d_cnt.p);
  if (d_left && d_right)
    k_probe_join<true><<<n_tiles, kProbeBlock, 0, s->stream>>>(idx->view(), d_key, d_start, d_end, n, cnt, tile_state,
                                                              ticket, result, d_left, d_right, capacity);
  else
    k_probe_join<false><<<n_tiles, kProbeBlock, 0, s->stream>>>(idx->view(), d_key, d_start, d_end, n, cnt, tile_state,
                                                               ticket, result, d_left, nullptr, capacity);
  SQ_CUDA(E, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

}  // namespace sq
