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
