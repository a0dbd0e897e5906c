// Searches over the SoA arrays of the flat index, shared by the probe kernels of sq_probe.cu (count -> scan -> write)
// and sq_probe_rank.cu (rank-difference count fused with the write).
#pragma once
#include "sq_internal.cuh"

namespace sq {

// first j in [m.sb, m.se) with vals[j] > q, for a segment-sorted array `vals` with its bin directory `dir`
// (one directory load, then one round of independent loads inside the bin)
// compact: a plain binary search inside the bin — a quarter of the instructions, dependent loads; for cache-resident
// indexes, where the loads are L1 / L2 hits and the kernels are bound by their instruction stream (cfg3: the thread-per-row
// searches cost 38 warp instructions per probe row in the latency-oriented form)
template <bool COMPACT = false>
__device__ __forceinline__ uint32_t upper_bound_dir(const int32_t* __restrict__ vals, const uint32_t* __restrict__ dir,
                                                    const SegMeta& m, int32_t qe) {
  uint32_t a, len;
  if (qe < m.min_start) {
    a = m.sb; len = 0;
  } else {
    const uint32_t off = uint32_t(qe) - uint32_t(m.min_start);
    const uint32_t b = m.shift >= 32 ? 0u : (off >> m.shift);
    if (b >= m.nbins) {
      a = m.se; len = 0;
    } else {
      a = __ldg(dir + m.dir_base + b);
      len = __ldg(dir + m.dir_base + b + 1) - a;
    }
  }
  if constexpr (COMPACT) {
    while (len) {
      const uint32_t half = len >> 1;
      if (__ldg(vals + a + half) <= qe) { a += half + 1; len -= half + 1; } else len = half;
    }
    return a;
  }
  while (len > 32) {  // crowded bin (skewed data): narrow it the classic way first
    const uint32_t half = len >> 1;
    if (__ldg(vals + a + half) <= qe) { a += half + 1; len -= half + 1; } else len = half;
  }
  if (len) {
    // last element of each group of 8 (clamped): 4 independent loads that touch every sector of the bin
    const int32_t* sp = vals + a;
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
  return a;
}

// Walks of long candidate ranges.  Calls leaf(j0) — by the whole warp — for every aligned block [j0, j0 + 32) of sorted
// rows that intersects [lo, hi) and whose maximum end reaches qs, in ascending order; blocks without a hit are skipped
// 32768, 1024 or 32 rows at a time.  A range of 4M rows behind one chromosome-long interval costs a few steps per
// level instead of 125,000 steps of 32 rows.
constexpr uint32_t kSkipMin = 2048;  // shorter candidate ranges are walked row by row
template <typename Leaf>
__device__ __forceinline__ void for_hit_blocks(const IndexView& iv, uint32_t lo, uint32_t hi, int32_t qs, int lane, Leaf&& leaf) {
  const uint32_t last = hi - 1;
  for (uint32_t c3 = lo >> 15; c3 <= (last >> 15); c3 += 32) {
    const uint32_t b3 = c3 + lane;
    unsigned m3 = __ballot_sync(0xffffffffu, b3 <= (last >> 15) && __ldg(iv.bmax3 + b3) >= qs);
    while (m3) {
      const uint32_t B3 = c3 + uint32_t(__ffs(int(m3)) - 1);
      m3 &= m3 - 1;
      const uint32_t b2 = B3 * 32 + lane;
      unsigned m2 = __ballot_sync(0xffffffffu, b2 >= (lo >> 10) && b2 <= (last >> 10) && __ldg(iv.bmax2 + b2) >= qs);
      while (m2) {
        const uint32_t B2 = B3 * 32 + uint32_t(__ffs(int(m2)) - 1);
        m2 &= m2 - 1;
        const uint32_t b1 = B2 * 32 + lane;
        unsigned m1 = __ballot_sync(0xffffffffu, b1 >= (lo >> 5) && b1 <= (last >> 5) && __ldg(iv.bmax1 + b1) >= qs);
        while (m1) {
          const uint32_t B1 = B2 * 32 + uint32_t(__ffs(int(m1)) - 1);
          m1 &= m1 - 1;
          leaf(B1 * 32);
        }
      }
    }
  }
}

// thread-level: the last row j in [lo, hi) with end[j] >= qs, or hi when there is none
__device__ __forceinline__ uint32_t last_hit_below(const IndexView& iv, uint32_t lo, uint32_t hi, int32_t qs) {
  uint32_t j = hi;
  while (j > lo) {
    if ((j & 31u) == 0 && j - lo >= 32) {  // at a block boundary: whole blocks below j without a hit are skipped
      if ((j & 1023u) == 0 && j - lo >= 1024) {
        if ((j & 32767u) == 0 && j - lo >= 32768 && __ldg(iv.bmax3 + (j >> 15) - 1) < qs) { j -= 32768; continue; }
        if (__ldg(iv.bmax2 + (j >> 10) - 1) < qs) { j -= 1024; continue; }
      }
      if (__ldg(iv.bmax1 + (j >> 5) - 1) < qs) { j -= 32; continue; }
    }
    --j;
    if (__ldg(iv.end + j) >= qs) return j;
  }
  return hi;
}

struct Cand {
  uint32_t lo;  // first candidate (absolute position in the sorted arrays)
  uint32_t nc;  // number of candidates
};

// Both searches are written as a few rounds of INDEPENDENT loads (memory-level parallelism per
// thread) instead of dependent binary-search steps: every HBM/L2 round trip a probe row waits for
// is one of (1) the directory entry, (2) the four sampled starts that cover its bin, (3) the five
// speculative gallop points over runmax; the refinements that follow hit lines already in L1.
template <bool COMPACT = false>
__device__ __forceinline__ Cand find_candidates(const IndexView& iv, uint32_t id, int32_t qs, int32_t qe) {
  Cand c{0u, 0u};
  if (id == kNoKey) return c;  // key hash absent from the build side: no rows (interval_join.rs:965)
  const SegMeta m = iv.meta[id];

  const uint32_t hi = upper_bound_dir<COMPACT>(iv.start, iv.dir, m, qe);
  if (hi == m.sb) return c;
  if constexpr (COMPACT) {
    // lo = first j in [sb, hi) with runmax[j] >= qs: gallop back from hi (candidate lists are short), then bisect
    if (__ldg(iv.runmax + hi - 1) < qs) return c;
    uint32_t right = hi - 1, step = 1;  // right qualifies
    uint32_t left = m.sb;
    while (right - left >= step) {
      const uint32_t p = right - step;
      if (__ldg(iv.runmax + p) >= qs) { right = p; step <<= 1; } else { left = p + 1; break; }
    }
    while (left < right) {
      const uint32_t mid = (left + right) >> 1;
      if (__ldg(iv.runmax + mid) < qs) left = mid + 1; else right = mid;
    }
    c.lo = right;
    c.nc = hi - right;
    return c;
  }

  // ---- lo = first j in [sb, hi) with runmax[j] >= qs (runmax non-decreasing): speculative gallop
  const uint32_t span = hi - m.sb;  // rows available below hi
  const int32_t* rp = iv.runmax + hi;  // rp[-d] = row hi-d
  const int32_t r1 = __ldg(rp - int(min(1u, span)));
  const int32_t r2 = __ldg(rp - int(min(2u, span)));
  const int32_t r4 = __ldg(rp - int(min(4u, span)));
  const int32_t r8 = __ldg(rp - int(min(8u, span)));
  const int32_t r16 = __ldg(rp - int(min(16u, span)));
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
      if (__ldg(iv.runmax + p) >= qs) { right = p; step <<= 1; } else { left = p + 1; break; }
    }
  }
  if (left > right) left = right;  // clamped gallop points may coincide
  // first j in [left, right] with runmax[j] >= qs; right qualifies
  uint32_t len = right - left;
  uint32_t a = left;
  while (len) {
    const uint32_t half = len >> 1;
    if (__ldg(iv.runmax + a + half) < qs) { a += half + 1; len -= half + 1; } else len = half;
  }
  c.lo = a;
  c.nc = hi - a;
  return c;
}

}  // namespace sq
