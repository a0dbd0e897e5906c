// Device pieces shared by the probe kernels over the packed lines (sq_probe_packed.cu: one tile per CTA;
// sq_probe_pipe.cu: persistent CTAs with software-pipelined loads): decoding a line, the register-path
// walk in compacted rounds, the chained scan across tiles and the ordered emit of a warp's rows.
#pragma once
#include "sq_internal.cuh"
#include "sq_probe_common.cuh"

namespace sq {

constexpr uint32_t kSlots = 32;          // stash slots per probe row
constexpr uint32_t kStride = kSlots + 1; // padded row stride of the stash (bank spread)

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
  // directory entry b + 1 = first row whose bin is > b: every row with start <= qe lies below it, and
  // dir_line gives the line of the last such row directly
  const uint2 e = __ldg(iv.dir_line + m.dir_base + (b >= m.nbins ? m.nbins : b + 1u));
  r.first = m.line_base;
  // e.y = first start of line e.x: when it lies past qe the line holds no candidate and the walk starts a line earlier
#ifdef SQ_NO_LINE_SKIP  // A/B builds only
  r.line = e.x;
#else
  r.line = e.x - ((qe < int32_t(e.y) && e.x > r.first) ? 1u : 0u);
#endif
  r.act = true;
  return r;
}

// ---------------------------------------------------------------------------------------------
// Walk in rounds.  Lane L owns probe row L of the warp: `walking` = the row still has a line to visit,
// `ln` = that line, `cnt` = its hits so far.  Every round the walking rows are compacted (rank among
// walking rows -> step, group), each fetches ONE line with a cooperative 8-lane load, all steps of the warp
// at once (8 independent requests per lane in flight): a warp waits for as many memory round trips as its
// longest walk and executes only as many steps as it has (row, line) pairs.
// The kernel is bound by the LSU/MIO pipe as much as by memory latency (ncu: lsu data pipe 53 % busy, the
// highest unit), so what a step needs of its row travels through ONE 16-byte shared-memory record instead
// of a shuffle per field, and the line numbers of a round through two vector loads per group.
// ---------------------------------------------------------------------------------------------
struct __align__(16) RowState {
  int32_t qs, qe;
  uint32_t cnt_reach;  // [30:0] hits so far, [31] the last line said "an earlier row still reaches qs"
  uint32_t pad;
};
#ifndef SQ_WALK_DUMMY
#define SQ_WALK_DUMMY 1
#endif
struct __align__(16) WalkShared {
  RowState row[33];   // indexed by owner lane; [32] = a record no build row can hit (slots no row uses point at it)
  uint32_t line[32];  // line to fetch, indexed by slot = group * 8 + step
  uint8_t inv[32];    // owner lane, indexed by slot
};

// Variant for count-only launches.  Row records live in SLOT order (slot = group * 8 + step) for the duration of a round: a step reads its record at
// a fixed offset from the lane's group base (no owner decode, no address arithmetic) and the record names its
// owner lane; slots past the last walking row hold a record no build row can hit (qs = INT32_MAX, qe = INT32_MIN)
// and fetch line 0, so a step needs no "is this slot in use" predicate either.  Measured against the owner-order
// walk below (same box, ms per 12.5M probe rows): count only 0.642 vs 0.662; emitting 1.067 vs 1.028 (the stash
// address now depends on the record just loaded), position-sorted emitting 0.893 vs 0.890 — hence count only.
template <bool EMIT>
__device__ __forceinline__ void walk_rounds_slot_order(const IndexView& iv, uint32_t* stash, WalkShared& ws, int32_t my_qs,
                                            int32_t my_qe, uint32_t first, bool& walking, uint32_t& ln, uint32_t& cnt) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 3, sub = lane & 7, g0 = g * 8;
  const uint4* __restrict__ my_lines = iv.lines + sub;
  RowState* grp = ws.row + g0;  // the eight records of my group, one per step
  for (;;) {
    const unsigned A = __ballot_sync(0xffffffffu, walking);
    if (A == 0) break;
    const int n_walk = __popc(A);
    const int r = __popc(A & ((1u << lane) - 1u));  // my rank: served in step r >> 2 by group r & 3
    const int slot = (r & 3) * 8 + (r >> 2);
    __syncwarp();
    ws.row[lane] = RowState{INT32_MAX, INT32_MIN, 0u, uint32_t(lane)};
    ws.line[lane] = 0u;
    __syncwarp();
    if (walking) {
      ws.row[slot] = RowState{my_qs, my_qe, cnt, uint32_t(lane)};
      ws.line[slot] = ln;
    }
    __syncwarp();
    const uint4 la = *reinterpret_cast<const uint4*>(ws.line + g0), lb = *reinterpret_cast<const uint4*>(ws.line + g0 + 4);
    const uint32_t lines8[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
    const int n_steps = (n_walk + 3) >> 2;
    uint4 v[8];
#pragma unroll
    for (int st = 0; st < 8; ++st)
      if (st < n_steps)  // warp-uniform: later rounds have few steps
        v[st] = __ldg(my_lines + size_t(lines8[st]) * 8);
#pragma unroll
    for (int st = 0; st < 8; ++st) {
      if (st >= n_steps) break;  // warp-uniform
      const RowState rs = grp[st];
      const uint32_t c0 = rs.cnt_reach;
      const uint4 d = v[st];
      const int32_t base = int32_t(__shfl_sync(0xffffffffu, d.x, g0));
      const Slots s = slots_of(d, sub);
      const bool ha = row_hits(s.a_lo, s.a_id, base, rs.qs, rs.qe);
      const bool hb = row_hits(s.b_lo, s.b_id, base, rs.qs, rs.qe);
      const uint32_t ma = (__ballot_sync(0xffffffffu, ha) >> g0) & 0xFFu;
      const uint32_t mb = (__ballot_sync(0xffffffffu, hb) >> g0) & 0xFFu;
      if (EMIT) {
        const uint32_t below = (1u << sub) - 1u;
        const uint32_t pa = c0 + __popc(ma & below);
        const uint32_t pb = c0 + __popc(ma) + __popc(mb & below);
        uint32_t* mine = stash + rs.pad * kStride;
        if (ha && pa < kSlots) mine[pa] = s.a_id;
        if (hb && pb < kSlots) mine[pb] = s.b_id;
      }
      // the group's first lane holds the line header: new count and "an earlier row still reaches qs"
      if (sub == 0) grp[st].cnt_reach = (c0 + __popc(ma) + __popc(mb)) | (int32_t(d.y) >= rs.qs ? 0x80000000u : 0u);
    }
    __syncwarp();
    if (walking) {
      const uint32_t cr = ws.row[slot].cnt_reach;
      cnt = cr & 0x7FFFFFFFu;
      walking = (cr >> 31) != 0u && ln > first;
      ln -= 1;
    }
  }
}

template <bool EMIT>
__device__ __forceinline__ void walk_rounds(const IndexView& iv, uint32_t* stash, WalkShared& ws, int32_t my_qs,
                                            int32_t my_qe, uint32_t first, bool& walking, uint32_t& ln, uint32_t& cnt) {
  if (!EMIT) {
    walk_rounds_slot_order<false>(iv, stash, ws, my_qs, my_qe, first, walking, ln, cnt);
    return;
  }
  const int lane = threadIdx.x & 31;
  const int g = lane >> 3, sub = lane & 7, g0 = g * 8;
  const uint4* __restrict__ my_lines = iv.lines + sub;
  ws.row[lane] = RowState{my_qs, my_qe, 0u, 0u};
#if SQ_WALK_DUMMY
  if (lane == 0) ws.row[32] = RowState{INT32_MAX, INT32_MIN, 0u, 0u};
#endif
  for (;;) {
    const unsigned A = __ballot_sync(0xffffffffu, walking);
    if (A == 0) break;
    const int n_walk = __popc(A);
    const int r = __popc(A & ((1u << lane) - 1u));  // my rank: served in step r >> 2 by group r & 3
    __syncwarp();
    ws.line[lane] = 0u;  // slots past the last walking row fetch line 0 (always there) and are ignored
#if SQ_WALK_DUMMY
    ws.inv[lane] = 32;   // ... and test it against the record nothing hits: a step needs no "slot in use" predicate
#endif
    __syncwarp();
    if (walking) {
      const int slot = (r & 3) * 8 + (r >> 2);
      ws.inv[slot] = uint8_t(lane);
      ws.line[slot] = ln;
    }
    __syncwarp();
    const unsigned long long srcs = *reinterpret_cast<const unsigned long long*>(ws.inv + g0);  // my group's 8 rows
    const uint4 la = *reinterpret_cast<const uint4*>(ws.line + g0), lb = *reinterpret_cast<const uint4*>(ws.line + g0 + 4);
    const uint32_t lines8[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
    const int n_steps = (n_walk + 3) >> 2;
    uint4 v[8];
#pragma unroll
    for (int st = 0; st < 8; ++st)
      if (st < n_steps)  // warp-uniform: later rounds have few steps
        v[st] = __ldg(my_lines + size_t(lines8[st]) * 8);
#pragma unroll
    for (int st = 0; st < 8; ++st) {
      if (st >= n_steps) break;  // warp-uniform
#if SQ_WALK_DUMMY
      const int p = int(srcs >> (8 * st)) & 63;
      const bool valid = true;
#else
      const int p = int(srcs >> (8 * st)) & 31;
      const bool valid = 4 * st + g < n_walk;
#endif
      const RowState rs = ws.row[p];
      const uint32_t c0 = rs.cnt_reach & 0x7FFFFFFFu;
      const uint4 d = v[st];
      const int32_t base = int32_t(__shfl_sync(0xffffffffu, d.x, g0));
      const Slots s = slots_of(d, sub);
      const bool ha = valid && row_hits(s.a_lo, s.a_id, base, rs.qs, rs.qe);
      const bool hb = valid && row_hits(s.b_lo, s.b_id, base, rs.qs, rs.qe);
      const uint32_t ma = (__ballot_sync(0xffffffffu, ha) >> g0) & 0xFFu;
      const uint32_t mb = (__ballot_sync(0xffffffffu, hb) >> g0) & 0xFFu;
      if (EMIT) {
        const uint32_t below = (1u << sub) - 1u;
        const uint32_t pa = c0 + __popc(ma & below);
        const uint32_t pb = c0 + __popc(ma) + __popc(mb & below);
        if (ha && pa < kSlots) stash[p * kStride + pa] = s.a_id;
        if (hb && pb < kSlots) stash[p * kStride + pb] = s.b_id;
      }
      // the group's first lane holds the line header: new count and "an earlier row still reaches qs"
      if (sub == 0 && valid)
        ws.row[p].cnt_reach = (c0 + __popc(ma) + __popc(mb)) | (int32_t(d.y) >= rs.qs ? 0x80000000u : 0u);
    }
    __syncwarp();
    if (walking) {
      const uint32_t cr = ws.row[lane].cnt_reach;
      cnt = cr & 0x7FFFFFFFu;
      walking = (cr >> 31) != 0u && ln > first;
      ln -= 1;
    }
  }
}



// ---------------------------------------------------------------------------------------------
// Chained scan with decoupled look-back over tiles (one 64-bit word per tile: [63:62] status).  Called
// by ONE full warp of the tile's CTA: publishes the tile's aggregate, sums its predecessors' (32 words
// per step, stopping at the first inclusive prefix), publishes the inclusive prefix and returns the
// exclusive one.  Predecessor tiles must already be running (ticket order / resident persistent grid).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long chain_lookback(unsigned long long* chain_state, uint32_t tile,
                                                              unsigned long long agg, uint32_t backoff_ns = 64) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) atomicExch(chain_state + tile, (tile == 0 ? kFlagInc : kFlagAgg) | agg);
  unsigned long long excl = 0;
  if (tile > 0) {
    int64_t look = int64_t(tile) - 1;
    for (;;) {
      const int64_t k = look - lane;
      unsigned long long x = kFlagInc;
      if (k >= 0) {
        // back off while a predecessor has not published yet: a spinning warp takes issue slots away from
        // the warps of the other CTAs on this SM that are doing the predecessors' work
        while (((x = *reinterpret_cast<volatile unsigned long long*>(chain_state + k)) >> 62) == 0) __nanosleep(backoff_ns);
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
    if (lane == 0) atomicExch(chain_state + tile, kFlagInc | (excl + agg));
  }
  return excl;
}

// ---------------------------------------------------------------------------------------------
// Ordered emit of one warp's 32 rows.  Row L's pairs go to lout/rout[coff_L ...] (coff = exclusive prefix
// of the hit counts inside the warp).  Rows with <= kSlots hits are copied from the stash as ONE flattened
// list (coalesced stores); rows with more hits are re-walked by the whole warp, four lines per step, from
// their start line `line0`.
// ---------------------------------------------------------------------------------------------
template <bool WRITE_RIGHT>
__device__ __forceinline__ void emit_rows(const IndexView& iv, const uint32_t* stash, uint8_t* inv, uint32_t cnt,
                                          uint32_t coff, int32_t my_qs, int32_t my_qe, uint32_t line0, uint32_t first,
                                          uint32_t* __restrict__ lout, uint32_t* __restrict__ rout, uint32_t tile_first) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 3, sub = lane & 7, g0 = g * 8;
  {
    const Flat f = flat_setup(cnt <= kSlots ? cnt : 0u, lane, inv);  // also orders the stash writes
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
  unsigned big = __ballot_sync(0xffffffffu, cnt > kSlots);
  while (big) {
    const int p = __ffs(big) - 1;
    big &= big - 1;
    const int32_t qs = __shfl_sync(0xffffffffu, my_qs, p);
    const int32_t qe = __shfl_sync(0xffffffffu, my_qe, p);
    uint32_t ln = __shfl_sync(0xffffffffu, line0, p);
    const uint32_t fst = __shfl_sync(0xffffffffu, first, p);
    uint32_t run = __shfl_sync(0xffffffffu, coff, p);
    for (;;) {
      const bool has = ln - fst >= uint32_t(g);  // group g takes line ln - g
      const uint4 d = has ? __ldg(iv.lines + size_t(ln - g) * 8 + sub) : make_uint4(0u, uint32_t(INT32_MIN), 0u, kEmptyRow);
      const int32_t lbase = int32_t(__shfl_sync(0xffffffffu, d.x, g0));
      const int32_t exmax = int32_t(__shfl_sync(0xffffffffu, d.y, g0));
      const bool more = has && exmax >= qs && (ln - g) > fst;  // this line sends the walk one line further
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

}  // namespace sq
