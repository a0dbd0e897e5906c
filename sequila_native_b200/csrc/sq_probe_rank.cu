// Probe + emit over the SoA arrays in ONE kernel, for indexes that carry the rank structure over the ends
// (cache-resident, wide or deep indexes: cfg2 / cfg3 / cfg4).  Replaces the per-row loop of process_probe_batch
// (reference interval_join.rs:1586-1618: hash_map.get -> coitrees query -> pos_vect / rle_right -> index_right).
//
// The count -> scan -> write chain of sq_probe.cu walks every candidate TWICE (k_probe_count to count, k_probe_write to
// write; ncu on cfg3: the count kernel costs more than the write one, at 5 % DRAM utilisation), because a probe row's
// output offset needs the counts of all rows before it.  With every build row well-formed (start <= end) the count needs
// no walk: hits = |{start <= qe}| - |{end < qs}| inside the key segment, two directory-guided
// upper bounds — the first is the candidate range's `hi` the walk needs anyway, the second runs over the same rows'
// ends sorted per segment (IndexView::send, built with the index).  So:
//
//   phase 1  thread = probe row: key id, candidate range [lo, hi) (find_candidates), count = (hi - sb) - rank of qs in
//            the sorted ends.  Probe rows with qs > qe + 1 (inverted rows: the rank identity does not hold) count their
//            candidates one by one.
//   phase 2  CTA total -> chained scan with decoupled look-back (tiles in ticket order) -> output base.  Count-only
//            launches stop here: no candidate was touched.
//   phase 3  the ONE walk: rows with <= 32 candidates flattened over the warp (coalesced loads of end[] / row[], no lane
//            idles on a short list), a hit's position = row offset + hits of the same row in earlier lanes / chunks;
//            rows with more candidates one at a time, 32 candidates per step; left_idx / right_idx stores fill the
//            warp's contiguous run of the output.
// Integer work on L2-resident arrays, bounded by issue rate and HBM writes; tensor cores do not apply.
#include "sq_internal.cuh"
#include "sq_probe_common.cuh"
#include "sq_soa_common.cuh"

namespace sq {

constexpr int kRB = 256;  // probe rows per CTA
constexpr int kRWarps = kRB / 32;
// Rows with up to kFlatMax candidates are walked flattened (no lane idles on a short list; ~70 warp instructions per 32
// candidates for finding each candidate's row and its position); longer rows take the whole warp, one row after the
// other, 32 candidates per step (~15 instructions per step, but the steps of a warp's rows are dependent load -> ballot
// round trips in sequence).  Measured on cfg3 (24 candidates per row): flattening up to 32 candidates 1.51 ms, up to 8
// candidates 1.75 ms.  Taking four steps at a time (all end[] loads, then the ballots, then all row[] loads, then the stores:
// two L2 round trips per group instead of per step) was measured too: cfg4 0.48-0.51 ms instead of 0.38-0.40 ms, cfg3 1.53
// instead of 1.51 ms — 61 registers instead of 48 and the guards of partial groups cost more than the shorter chains save.
#ifndef SQ_RANK_FLAT_MAX
#define SQ_RANK_FLAT_MAX 32
#endif
constexpr uint32_t kFlatMax = SQ_RANK_FLAT_MAX;

// chain_lookback lives in sq_packed_common.cuh together with packed-line helpers; the same protocol, restated here
// for a CTA's first warp (status word: [63:62] flag, [61:0] value)
__device__ __forceinline__ unsigned long long rank_lookback(unsigned long long* chain_state, uint32_t tile,
                                                            unsigned long long agg, uint32_t backoff_ns) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) atomicExch(chain_state + tile, (tile == 0 ? kFlagInc : kFlagAgg) | agg);
  unsigned long long excl = 0;
  if (tile > 0) {
    int64_t look = int64_t(tile) - 1;
    for (;;) {
      const int64_t k = look - lane;
      unsigned long long x = kFlagInc;
      if (k >= 0)
        while (((x = *reinterpret_cast<volatile unsigned long long*>(chain_state + k)) >> 62) == 0) __nanosleep(backoff_ns);
      const unsigned inc_mask = __ballot_sync(0xffffffffu, (x >> 62) == 2);
      const int first_inc = inc_mask ? (__ffs(inc_mask) - 1) : 32;
      unsigned long long y = (lane <= first_inc) ? (x & kValMask) : 0;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) y += __shfl_xor_sync(0xffffffffu, y, d);
      excl += y;
      if (inc_mask) break;
      look -= 32;
    }
    if (lane == 0) atomicExch(chain_state + tile, kFlagInc | (excl + agg));
  }
  return excl;
}

// COMPACT: the searches in their few-instruction form (cache-resident indexes, see sq_soa_common.cuh)
template <bool EMIT, bool WRITE_RIGHT, bool COMPACT>
__global__ void __launch_bounds__(kRB, 4)
k_probe_rank(IndexView iv, const uint64_t* __restrict__ q_key, const int32_t* __restrict__ q_start,
             const int32_t* __restrict__ q_end, uint32_t n, uint32_t* __restrict__ cnt_out,
             unsigned long long* chain_state, unsigned int* ticket, unsigned long long* result,
             uint32_t* __restrict__ left_out, uint32_t* __restrict__ right_out, uint64_t capacity, uint32_t n_tiles,
             uint32_t backoff_ns) {
  __shared__ unsigned long long s_wtot[kRWarps];
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_bid;
  __shared__ uint8_t s_inv[kRWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t tile = blockIdx.x;
  if (EMIT) {  // tiles in ticket order: every predecessor in the chained scan is already running
    if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    tile = s_bid;
  }
  const uint64_t i = uint64_t(tile) * kRB + threadIdx.x;

  // ---- phase 1: candidate range and count, no candidate touched ------------------------------------------------
  Cand c{0u, 0u};
  int32_t qs = 0;
  uint32_t cnt = 0;
  if (i < n) {
    qs = q_start[i];
    const int32_t qe = q_end[i];
    const uint32_t id = ht_lookup(iv.ht_keys, iv.ht_ids, iv.ht_mask, iv.sentinel_id, q_key[i]);
    c = find_candidates<COMPACT>(iv, id, qs, qe);
    if (c.nc) {
      if ((long long)qs <= (long long)qe + 1) {
        const SegMeta m = iv.meta[id];
        // rows of the segment with end < qs = first position of the sorted ends with end > qs - 1
        const uint32_t below = qs == INT32_MIN ? 0u : upper_bound_dir<COMPACT>(iv.send, iv.edir, iv.emeta[id], qs - 1) - m.sb;
        cnt = (c.lo + c.nc - m.sb) - below;
      } else {  // inverted probe row: {end < qs} is no subset of {start <= qe}; count the candidates
        for (uint32_t j = c.lo + c.nc;;) {
          const uint32_t f = last_hit_below(iv, c.lo, j, qs);  // == j: no further hit below j
          if (f == j) break;
          ++cnt;
          j = f;
        }
      }
    }
    cnt_out[i] = cnt;  // rle_right (interval_join.rs:1604)
  }
  const uint32_t cincl = warp_incl_sum(cnt);  // a warp emits < 2^32 pairs unless rows hit > 2^27 builds each
  const uint32_t wtot = __shfl_sync(0xffffffffu, cincl, 31);
  if (lane == 0) s_wtot[warp] = wtot;
  __syncthreads();
  unsigned long long cta_tot = 0, before = 0;
#pragma unroll
  for (int w = 0; w < kRWarps; ++w) {
    if (w < warp) before += s_wtot[w];
    cta_tot += s_wtot[w];
  }
  if (!EMIT) {
    if (threadIdx.x == 0 && cta_tot) atomicAdd(result, cta_tot);
    return;
  }

  // ---- phase 2: chained scan over tiles ---------------------------------------------------------------------------
  if (warp == 0) {
    const unsigned long long excl = rank_lookback(chain_state, tile, cta_tot, backoff_ns);
    if (lane == 0) {
      s_base = excl;
      if (tile == n_tiles - 1) result[0] = excl + cta_tot;
      if (excl + cta_tot > capacity) result[1] = 1;  // the caller's buffers are too small: report, write nothing here
    }
  }
  __syncthreads();
  if (wtot == 0 || s_base + cta_tot > capacity) return;  // warp-uniform / CTA-uniform
  const uint64_t base = s_base + before;
  uint32_t* __restrict__ lout = left_out + base;
  uint32_t* __restrict__ rout = WRITE_RIGHT ? right_out + base : nullptr;
  const uint32_t tile_first = tile * kRB + warp * 32;
  const uint32_t coff = cincl - cnt;  // offset of my row's first pair inside the warp's run

  // ---- phase 3: the walk.  Small rows flattened: candidate t of the warp's concatenated list -> lane t ---------
  {
    const bool small = cnt != 0 && c.nc <= kFlatMax;
    const Flat f = flat_setup(small ? c.nc : 0u, lane, s_inv[warp]);
    const uint32_t r_jbase = __shfl_sync(0xffffffffu, c.lo, f.r_src) - f.r_excl;  // candidate t -> row r_jbase + t
    const int32_t r_qs = __shfl_sync(0xffffffffu, qs, f.r_src);
    const uint32_t r_coff = __shfl_sync(0xffffffffu, coff, f.r_src);
    uint32_t r_done = 0;  // (rank lane) hits of my rank's row in earlier chunks
    for (uint32_t t0 = 0; t0 < f.total; t0 += 32) {
      const uint32_t t = t0 + lane;
      const int r = flat_rank(f, t0, lane);
      const uint32_t j = __shfl_sync(0xffffffffu, r_jbase, r) + t;  // 32-bit wrap-around is intended
      const int32_t pqs = __shfl_sync(0xffffffffu, r_qs, r);
      const uint32_t off = __shfl_sync(0xffffffffu, r_coff, r);
      const uint32_t done = __shfl_sync(0xffffffffu, r_done, r);
      const uint32_t excl_r = __shfl_sync(0xffffffffu, f.r_excl, r);
      const int src = __shfl_sync(0xffffffffu, f.r_src, r);
      const bool hit = (t < f.total) && (__ldg(iv.end + j) >= pqs);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const uint32_t a = excl_r > t0 ? excl_r - t0 : 0u;  // my row's first lane in this chunk
        const uint32_t pos = off + done + __popc(m & ((1u << lane) - 1u) & ~low_bits(a));
        lout[pos] = __ldg(iv.row + j);
        if (WRITE_RIGHT) rout[pos] = tile_first + src;
      }
      // rank lanes: the share of my rank's row in this chunk
      const uint32_t a = max(f.r_excl, t0), b = min(f.r_incl, t0 + 32);
      if (f.r_incl != 0xffffffffu && a < b) r_done += __popc((m >> (a - t0)) & low_bits(b - a));
    }
  }
  // ---- big rows, the whole warp on one row, 32 candidates per step ------------------------------------------------
  unsigned big = __ballot_sync(0xffffffffu, cnt != 0 && c.nc > kFlatMax);
  while (big) {
    const int p = __ffs(big) - 1;
    big &= big - 1;
    const uint32_t nc_p = __shfl_sync(0xffffffffu, c.nc, p);
    const uint32_t lo_p = __shfl_sync(0xffffffffu, c.lo, p);
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
    if (nc_p > kSkipMin) for_hit_blocks(iv, lo_p, lo_p + nc_p, qs_p, lane, block);  // long range: only blocks with a hit
    else for (uint32_t k0 = 0; k0 < nc_p; k0 += 32) block(lo_p + k0);
  }
}

// Measured on B200 (profiles/r02_subcfg_rank.json): the rank kernel replaces the count walk by a second directory-guided
// search.  It wins where candidate lists are long — cfg4 (~100 rows deep): join 0.38 vs 0.50 ms, count(1) 0.12 vs 0.19 ms —
// and loses where they are short and the searches expensive — cfg3 (~20 deep, starts clustered into crowded directory
// bins): 1.51 vs 1.32 ms; cfg2 (depth ~0) is launch-bound either way.  The build samples the overlap depth.
bool use_rank(const sq_index* idx) {
  if (!idx->d_send || use_packed(idx)) return false;
  const int opt = idx->ctx->opt.rank_count.load(std::memory_order_relaxed);
  if (opt == 0) return false;
  if (opt == 2) return true;
  return idx->mean_depth >= 48.f;
}

int launch_rank_join(sq_stream* s, const sq_index* idx, const uint64_t* d_key, const int32_t* d_start,
                     const int32_t* d_end, uint32_t n, uint32_t* d_left, uint32_t* d_right, uint64_t capacity) {
  ErrorSlot& E = s->err;
  const uint32_t n_tiles = (n + kRB - 1) / kRB;
  int rc;
  if ((rc = ensure(E, s->d_cnt, size_t(n) * 4, false))) return rc;
  if ((rc = ensure(E, s->d_tile, size_t(n_tiles) * 8 + 16, false))) return rc;
  if ((rc = ensure(E, s->d_scalar, 256, false))) return rc;
  auto* chain = static_cast<unsigned long long*>(s->d_tile.p);
  auto* ticket = reinterpret_cast<unsigned int*>(chain + n_tiles);
  auto* result = static_cast<unsigned long long*>(s->d_scalar.p);  // [0] n_pairs [1] overflow
  auto* cnt = static_cast<uint32_t*>(s->d_cnt.p);
  SQ_CUDA(E, cudaMemsetAsync(result, 0, 32, s->stream));
  if (d_left) SQ_CUDA(E, cudaMemsetAsync(chain, 0, size_t(n_tiles) * 8 + 16, s->stream));
  const IndexView iv = idx->view();
  const uint32_t backoff = uint32_t(s->ctx->opt.lookback_backoff_ns.load(std::memory_order_relaxed));
  // arrays the searches touch: start, runmax, sorted ends, two directories — in L2 (126 MB) with room for the streams?
  const bool compact = idx->n_rows * 14ull <= (48ull << 20);
#define SQ_RANK_LAUNCH(E_, R_, C_)                                                                                       \
  k_probe_rank<E_, R_, C_><<<n_tiles, kRB, 0, s->stream>>>(iv, d_key, d_start, d_end, n, cnt, chain, ticket, result, \
                                                          E_ ? d_left : nullptr, R_ ? d_right : nullptr, E_ ? capacity : 0, n_tiles, E_ ? backoff : 0u)
  if (!d_left) { if (compact) SQ_RANK_LAUNCH(false, false, true); else SQ_RANK_LAUNCH(false, false, false); }
  else if (d_right) { if (compact) SQ_RANK_LAUNCH(true, true, true); else SQ_RANK_LAUNCH(true, true, false); }
  else { if (compact) SQ_RANK_LAUNCH(true, false, true); else SQ_RANK_LAUNCH(true, false, false); }
#undef SQ_RANK_LAUNCH
  SQ_CUDA(E, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

}  // namespace sq
