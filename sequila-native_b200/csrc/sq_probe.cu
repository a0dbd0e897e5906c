// Probe + emit of the `Cuda` interval join: replaces the per-row loop of process_probe_batch
// (reference interval_join.rs:1586-1618: hash_map.get -> coitrees query -> pos_vect / rle_right
// -> index_right) with two kernels over a tile of probe rows:
//
//   k_probe_count   one thread per probe row: key hash -> id, two binary searches in the key's
//                   segment give the contiguous candidate range [lo, hi); each warp then walks
//                   the 32 rows' candidate ranges as ONE flattened list (coalesced reads of end[],
//                   no lane idles on a short list) and counts hits per row.  The CTA total goes
//                   through a decoupled look-back so every CTA learns the output offset of its
//                   first pair in the same pass (no separate scan kernel, no per-row offsets).
//   k_probe_write   same flattened walk; because output offsets of consecutive probe rows are
//                   contiguous, a warp's hits form one contiguous run of the output: position =
//                   warp base + ballot rank.  Stores of left_idx/right_idx are fully coalesced.
//
// Both kernels are integer/byte work bounded by HBM (or by L2 latency when the index is small);
// tensor cores do not apply.
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

// hi = first j in [sb,se) with start[j] > qe ; lo = first j in [sb,hi) with runmax[j] >= qs
__device__ __forceinline__ Cand find_candidates(const IndexView& iv, uint32_t id, int32_t qs, int32_t qe) {
  Cand c{0u, 0u};
  if (id == kNoKey) return c;
  const uint32_t sb = __ldg(iv.seg_off + id), se = __ldg(iv.seg_off + id + 1);
  uint32_t a = sb, len = se - sb;
  while (len) {
    const uint32_t half = len >> 1;
    if (__ldg(iv.start + a + half) <= qe) { a += half + 1; len -= half + 1; } else len = half;
  }
  const uint32_t hi = a;
  a = sb; len = hi - sb;
  while (len) {
    const uint32_t half = len >> 1;
    if (__ldg(iv.runmax + a + half) < qs) { a += half + 1; len -= half + 1; } else len = half;
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

__global__ void __launch_bounds__(kProbeBlock)
k_probe_count(IndexView iv, const uint64_t* __restrict__ q_key, const int32_t* __restrict__ q_start,
              const int32_t* __restrict__ q_end, uint32_t n, uint32_t* __restrict__ lo_out,
              uint32_t* __restrict__ nc_out, uint32_t* __restrict__ cnt_out,
              unsigned long long* tile_state, uint64_t* __restrict__ tile_base,
              unsigned int* ticket, unsigned long long* n_pairs_out) {
  __shared__ uint32_t s_cnt[kProbeBlock];
  __shared__ uint64_t s_wtot[kWarpsPerBlock];
  __shared__ uint32_t s_bid;

  if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);  // CTAs take tiles in start order
  __syncthreads();
  const uint32_t bid = s_bid;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t i = bid * kProbeBlock + threadIdx.x;

  Cand c{0u, 0u};
  int32_t qs = 0;
  if (i < n) {
    qs = q_start[i];
    const uint32_t id = ht_lookup(iv.ht_keys, iv.ht_ids, iv.ht_mask, iv.sentinel_id, q_key[i]);
    c = find_candidates(iv, id, qs, q_end[i]);
  }

  // flattened walk over the warp's candidates, counting hits per owner row
  uint32_t* wcnt = s_cnt + warp * 32;
  wcnt[lane] = 0;
  const uint32_t incl = warp_incl_sum(c.nc);
  const uint32_t excl = incl - c.nc;
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  __syncwarp();
  for (uint32_t t0 = 0; t0 < total; t0 += 32) {
    const uint32_t t = t0 + lane;
    const int p = owner_of(incl, t);
    const uint32_t j = __shfl_sync(0xffffffffu, c.lo, p) + (t - __shfl_sync(0xffffffffu, excl, p));
    const int32_t pqs = __shfl_sync(0xffffffffu, qs, p);
    const bool hit = (t < total) && (__ldg(iv.end + j) >= pqs);
    const unsigned peers = __match_any_sync(0xffffffffu, hit ? p : 32 + lane);
    if (hit && (__ffs(peers) - 1) == lane) wcnt[p] += __popc(peers);
    __syncwarp();
  }
  const uint32_t cnt = wcnt[lane];
  if (i < n) {
    lo_out[i] = c.lo;
    nc_out[i] = c.nc;
    cnt_out[i] = cnt;
  }

  // CTA total -> decoupled look-back -> exclusive base of this CTA
  uint64_t wsum = cnt;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, d);
  if (lane == 0) s_wtot[warp] = wsum;
  __syncthreads();
  if (warp == 0) {
    uint64_t agg = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) agg += s_wtot[w];
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
      tile_base[bid] = excl_base;
      if (bid == gridDim.x - 1) *n_pairs_out = excl_base + agg;
    }
  }
}

template <bool WRITE_RIGHT>
__global__ void __launch_bounds__(kProbeBlock)
k_probe_write(IndexView iv, const int32_t* __restrict__ q_start, uint32_t n,
              const uint32_t* __restrict__ lo_in, const uint32_t* __restrict__ nc_in,
              const uint32_t* __restrict__ cnt_in, const uint64_t* __restrict__ tile_base,
              uint32_t* __restrict__ left_out, uint32_t* __restrict__ right_out) {
  __shared__ uint64_t s_wtot[kWarpsPerBlock];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t i = blockIdx.x * kProbeBlock + threadIdx.x;
  const uint32_t tile_first = blockIdx.x * kProbeBlock + warp * 32;

  uint32_t lo = 0, nc = 0, cnt = 0;
  int32_t qs = 0;
  if (i < n) {
    lo = lo_in[i];
    nc = nc_in[i];
    cnt = cnt_in[i];
    qs = q_start[i];
  }
  uint64_t wsum = cnt;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, d);
  if (lane == 0) s_wtot[warp] = wsum;
  __syncthreads();
  if (wsum == 0) return;  // warp-uniform
  uint64_t base = tile_base[blockIdx.x];
  for (int w = 0; w < warp; ++w) base += s_wtot[w];

  const uint32_t incl = warp_incl_sum(nc);
  const uint32_t excl = incl - nc;
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  for (uint32_t t0 = 0; t0 < total; t0 += 32) {
    const uint32_t t = t0 + lane;
    const int p = owner_of(incl, t);
    const uint32_t j = __shfl_sync(0xffffffffu, lo, p) + (t - __shfl_sync(0xffffffffu, excl, p));
    const int32_t pqs = __shfl_sync(0xffffffffu, qs, p);
    const bool hit = (t < total) && (__ldg(iv.end + j) >= pqs);
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
int launch_count(sq_stream* s, const sq_index* idx, const uint64_t* d_key, const int32_t* d_start,
                 const int32_t* d_end, uint32_t n) {
  ErrorSlot& E = s->err;
  const uint32_t n_tiles = (n + kProbeBlock - 1) / kProbeBlock;
  int rc;
  if ((rc = ensure(E, s->d_lo, size_t(n) * 4, false))) return rc;
  if ((rc = ensure(E, s->d_ncand, size_t(n) * 4, false))) return rc;
  if ((rc = ensure(E, s->d_cnt, size_t(n) * 4, false))) return rc;
  if ((rc = ensure(E, s->d_tile, size_t(n_tiles) * 16, false))) return rc;
  if ((rc = ensure(E, s->d_scalar, 256, false))) return rc;
  auto* tile_state = static_cast<unsigned long long*>(s->d_tile.p);
  auto* tile_base = reinterpret_cast<uint64_t*>(tile_state + n_tiles);
  auto* n_pairs = static_cast<unsigned long long*>(s->d_scalar.p);
  auto* ticket = reinterpret_cast<unsigned int*>(n_pairs + 1);
  SQ_CUDA(E, cudaMemsetAsync(tile_state, 0, size_t(n_tiles) * 8, s->stream));
  SQ_CUDA(E, cudaMemsetAsync(s->d_scalar.p, 0, 16, s->stream));
  k_probe_count<<<n_tiles, kProbeBlock, 0, s->stream>>>(
      idx->view(), d_key, d_start, d_end, n, static_cast<uint32_t*>(s->d_lo.p),
      static_cast<uint32_t*>(s->d_ncand.p), static_cast<uint32_t*>(s->d_cnt.p), tile_state, tile_base,
      ticket, n_pairs);
  SQ_CUDA(E, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

int launch_write(sq_stream* s, uint32_t* d_left, uint32_t* d_right) {
  ErrorSlot& E = s->err;
  const uint32_t n = s->n_rows;
  const uint32_t n_tiles = (n + kProbeBlock - 1) / kProbeBlock;
  auto* tile_state = static_cast<unsigned long long*>(s->d_tile.p);
  auto* tile_base = reinterpret_cast<const uint64_t*>(tile_state + n_tiles);
  const auto* lo = static_cast<const uint32_t*>(s->d_lo.p);
  const auto* nc = static_cast<const uint32_t*>(s->d_ncand.p);
  const auto* cnt = static_cast<const uint32_t*>(s->d_cnt.p);
  if (d_right)
    k_probe_write<true><<<n_tiles, kProbeBlock, 0, s->stream>>>(s->idx->view(), s->d_q_start, n, lo, nc, cnt,
                                                               tile_base, d_left, d_right);
  else
    k_probe_write<false><<<n_tiles, kProbeBlock, 0, s->stream>>>(s->idx->view(), s->d_q_start, n, lo, nc, cnt,
                                                                tile_base, d_left, nullptr);
  SQ_CUDA(E, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

}  // namespace sq
