// Probe + emit for POSITION-LOCAL probe tiles (BAM / BED order, or any order in which neighbouring probe rows
// land near each other in the index): the build tile a CTA needs is staged in shared memory by ONE bulk copy
// of the TMA engine and every thread then serves its own probe row out of shared memory.  Replaces the per-row
// loop of process_probe_batch (reference interval_join.rs:1586-1618: hash_map.get -> coitrees query -> pos_vect /
// rle_right -> index_right) like sq_probe_packed.cu does, for tiles where SURVEY.md Appendix D's strategy B applies.
//
// Why a second kernel: k_probe_packed spends 8 lanes on every (probe row, 128-byte line) pair because a random
// line must arrive as one cooperative request — 43 warp instructions per probe row, and with position-sorted
// probes (every line an L2 hit) it still runs at 0.81 ms per 12.5M rows: bound by its instruction stream
// (DESIGN.md section 4).  When the 128 probe rows of a CTA touch ONE contiguous range of lines, that range can be
// fetched once per CTA instead (cp.async.bulk global -> shared, completion on an mbarrier: SASS UBLKCP), decoded once
// per build row into (end, running max end) pairs, and searched by one THREAD per probe row: the warp-level cost of
// a probe row drops to roughly a third, and DRAM sees one streaming read of the lines instead of one request per
// (row, line).
//
//   phase 1  thread = probe row: key hash -> key id -> segment -> one directory entry = the line holding the last
//            start <= probe end (find_start_line, shared with k_probe_packed).
//   phase 2  CTA: min / max of those lines; [min - 2, max] fits the staging buffer -> one thread arms an mbarrier
//            and issues ONE cp.async.bulk for the whole range; otherwise the CTA takes the global path (below).
//   phase 3  decode: thread = line slot (16 per line: header + 15 rows); end = start + width, and an inclusive
//            prefix max over the 16-lane group seeded with the header's exmax turns the line-granular "an earlier
//            row still reaches qs" into a per-row running max — the same array the SoA layout keeps in HBM
//            (runmax), rebuilt in shared memory from the packed line.
//   phase 4  thread = probe row: 4-step upper bound inside its start line (16-bit start offsets), then a backward
//            scan over (end, runmax) pairs until runmax < probe start — exactly the candidate range [lo, hi) of
//            the flat index, no row tested twice.  A scan that runs off the staged range (a long overlap chain)
//            continues line by line from global memory.
//   phase 5  CTA total -> chained scan with decoupled look-back (tiles in ticket order) -> output base.
//   phase 6  hits go through a shared-memory window in output order and leave as coalesced stores of left_idx /
//            right_idx (a warp's rows own adjacent output runs; direct per-thread stores would cost one L2
//            transaction per pair).
// A CTA whose probe rows do not share a stageable range (random probe order, sparse probes over a dense index,
// a tile that straddles many contigs) walks lines from global memory thread by thread — correct for any input
// but slow; result[2] counts such CTAs so the host can send the next tiles of that stream to k_probe_packed
// (sq_api.cu: staged_feedback).  Integer work bounded by HBM bandwidth; tensor cores do not apply.
#include "sq_internal.cuh"
#include "sq_packed_common.cuh"

namespace sq {

constexpr int kSB = 128;                // probe rows per CTA = threads per CTA
constexpr int kSWarps = kSB / 32;
constexpr uint32_t kOutWin = 1024;      // pairs per output window in shared memory
constexpr uint32_t kHaloLines = 2;      // lines staged below the lowest start line (walk-back room)
constexpr uint32_t kCapLines = 96;      // staging capacity in lines (12 KB raw + 12 KB decoded)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

// one packed line from global memory, thread per row: count (and optionally report) the rows hit by (qs, qe)
template <typename F>
__device__ __forceinline__ uint32_t scan_global_line(const uint4* __restrict__ lp, int32_t qs, int32_t qe, int32_t& exmax, F&& push) {
  const uint4 h = __ldg(lp);
  const int32_t base = int32_t(h.x);
  exmax = int32_t(h.y);
  uint32_t c = 0;
  if (row_hits(h.z, h.w, base, qs, qe)) { push(h.w); ++c; }
#pragma unroll
  for (int k = 1; k < 8; ++k) {
    const uint4 d = __ldg(lp + k);
    if (row_hits(d.x, d.y, base, qs, qe)) { push(d.y); ++c; }
    if (row_hits(d.z, d.w, base, qs, qe)) { push(d.w); ++c; }
  }
  return c;
}

// walk lines ln, ln - 1, ... of the key segment that begins at `first` while an earlier row still reaches qs
template <typename F>
__device__ __forceinline__ uint32_t walk_global(const IndexView& iv, uint32_t ln, uint32_t first, int32_t qs, int32_t qe, F&& push) {
  uint32_t c = 0;
  for (;;) {
    int32_t exmax;
    c += scan_global_line(iv.lines + size_t(ln) * 8, qs, qe, exmax, push);
    if (!(exmax >= qs && ln > first)) break;
    --ln;
  }
  return c;
}

template <bool EMIT, bool WRITE_RIGHT>
__global__ void __launch_bounds__(kSB, 7)
k_probe_staged(IndexView iv, const uint64_t* __restrict__ q_key, const int32_t* __restrict__ q_start,
               const int32_t* __restrict__ q_end, uint32_t n, uint32_t* __restrict__ cnt_out,
               unsigned long long* chain_state, unsigned int* ticket, unsigned long long* result,
               uint32_t* __restrict__ left_out, uint32_t* __restrict__ right_out, uint64_t capacity,
               uint32_t n_tiles, uint32_t backoff_ns) {
  // raw lines as (lo, id) slot pairs: slot p = line * 16 + s; s = 0 is the header {base_start, exmax}, s = 1..15 row s-1
  __shared__ __align__(128) uint2 s_raw[kCapLines * 16];
  __shared__ __align__(16) int2 s_er[kCapLines * 16];  // decoded {end, running max of end} per slot (INT32_MIN end: no row)
  __shared__ uint8_t s_nrows[kCapLines];                // rows of each staged line
  __shared__ __align__(16) uint32_t s_out[EMIT ? kOutWin + 4 : 1];  // output window: left_idx (shifted by the output's misalignment)
  __shared__ __align__(16) uint8_t s_own[EMIT ? kOutWin + 4 : 1];   // ... and the probe row (thread) that owns each pair
  __shared__ __align__(8) unsigned long long s_mbar;
  __shared__ uint32_t s_min[kSWarps], s_max[kSWarps];
  __shared__ uint32_t s_stage_lo, s_nl, s_mode, s_bid;
  __shared__ unsigned long long s_wtot[kSWarps];
  __shared__ unsigned long long s_base;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t mbar = smem_u32(&s_mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t tile = blockIdx.x;
  if (EMIT) {  // tiles in ticket order: every predecessor in the chained scan is already running
    if (tid == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    tile = s_bid;
  }

  // ---- phase 1: my probe row -> its start line ------------------------------------------------------------
  const uint64_t i = uint64_t(tile) * kSB + tid;
  int32_t qs = 0, qe = 0;
  StartLine sl{0u, 0u, false};
  if (i < n) {
    qs = q_start[i];
    qe = q_end[i];
    const uint32_t id = ht_lookup(iv.ht_keys, iv.ht_ids, iv.ht_mask, iv.sentinel_id, q_key[i]);
    sl = find_start_line(iv, id, qe);
  }

  // ---- phase 2: the CTA's line range; one bulk copy stages it ------------------------------------------------
  {
    const uint32_t mn = __reduce_min_sync(0xffffffffu, sl.act ? sl.line : 0xFFFFFFFFu);
    const uint32_t mx = __reduce_max_sync(0xffffffffu, sl.act ? sl.line : 0u);
    if (lane == 0) { s_min[warp] = mn; s_max[warp] = mx; }
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll
    for (int w = 0; w < kSWarps; ++w) { mn = min(mn, s_min[w]); mx = max(mx, s_max[w]); }
    uint32_t mode = 0, lo = 0, nl = 0;  // 0: no row of this tile can match anything
    if (mn != 0xFFFFFFFFu) {
      lo = mn >= kHaloLines ? mn - kHaloLines : 0u;
      nl = mx - lo + 1u;
      if (nl <= kCapLines) {
        mode = 1;
        atomicAdd(result + 3, (unsigned long long)nl);  // staged lines, summed: the host learns how dense the tiles are
        const uint32_t bytes = nl * 128u;
        const uint4* src = iv.lines + size_t(lo) * 8;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(s_raw)), "l"(src), "r"(bytes), "r"(mbar) : "memory");
      } else {
        mode = 2;
        atomicAdd(result + 2, 1ull);  // tells the host that this stream's tiles are not position-local
      }
    }
    s_mode = mode;
    s_stage_lo = lo;
    s_nl = nl;
  }
  __syncthreads();
  const uint32_t mode = s_mode, stage_lo = s_stage_lo, nl = s_nl;

  // ---- phase 3: decode the staged lines ------------------------------------------------------------------------
  if (mode == 1) {
    uint32_t done;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(mbar), "r"(0u) : "memory");
    } while (!done);
    const uint32_t n_slots = nl * 16u;
    const int half = lane & 16;
    for (uint32_t p0 = uint32_t(warp) * 32u; p0 < n_slots; p0 += kSB) {
      const uint32_t p = p0 + lane;
      const bool inb = p < n_slots;
      const uint2 v = inb ? s_raw[p] : make_uint2(0u, kEmptyRow);
      const int s = int(p & 15u);
      const int32_t base = int32_t(__shfl_sync(0xffffffffu, v.x, half));
      const bool valid = inb && s > 0 && v.y != kEmptyRow;
      const int32_t en = valid ? base + int32_t(v.x & 0xFFFFu) + int32_t(v.x >> 16) : INT32_MIN;
      int32_t rm = (s == 0) ? int32_t(v.y) : en;  // the header carries exmax: max end of every earlier row of the segment
#pragma unroll
      for (int d = 1; d < 16; d <<= 1) {
        const int32_t o = __shfl_up_sync(0xffffffffu, rm, d, 16);
        if (s >= d) rm = max(rm, o);
      }
      const unsigned vm = __ballot_sync(0xffffffffu, valid);
      if (inb) {
        s_er[p] = make_int2(en, rm);
        if (s == 0) s_nrows[p >> 4] = uint8_t(__popc((vm >> half) & 0xFFFFu));
      }
    }
  }
  __syncthreads();

  // ---- phase 4: my row's hits ----------------------------------------------------------------------------------
  uint32_t cnt = 0;
  uint32_t mask = 0;            // bit k <=> staged slot j_hi - k is a hit (the first 32 scanned slots)
  int j_hi = 0, j_stop = 0;     // staged slots (j_stop, j_hi] were scanned
  bool cont = false;            // the walk continues in global memory ...
  uint32_t g_line = 0;          // ... from this line
  auto nothing = [](uint32_t) {};
  if (sl.act) {
    if (mode == 1) {
      uint32_t L = sl.line - stage_lo;
      const uint32_t Lmin = sl.first > stage_lo ? sl.first - stage_lo : 0u;
      int r;  // rows of line L with start <= qe (their 16-bit start offsets ascend)
      for (;;) {
        const int nr = int(s_nrows[L]);
        const long long d = (long long)qe - (long long)int32_t(s_raw[L * 16].x);
        const int target = d < 0 ? -1 : (d > 65535 ? 65535 : int(d));
        r = 0;
#pragma unroll
        for (int st = 8; st > 0; st >>= 1)
          if (r + st <= nr && int(s_raw[L * 16 + r + st].x & 0xFFFFu) <= target) r += st;
        // the directory names the line of the last row of the probe's BIN; when the bin's tail spans whole (short)
        // lines past qe, the last start <= qe lives further down: every row below the scan's entry point must satisfy
        // start <= qe, so step down until a line has such a row
        if (r > 0 || L <= Lmin) break;
        --L;
      }
      int j = int(L * 16) + r;
      const int jmin = int(Lmin * 16u);
      j_hi = j;
      uint32_t bit = 1u;  // shifts out after 32 slots: deeper hits are found again by the emit's rescan
      while (j >= jmin) {
        const int2 e = s_er[j];
        if (e.y < qs) break;  // no row at or below j reaches qs
        if (e.x >= qs) { ++cnt; mask |= bit; }
        bit <<= 1;
        --j;
      }
      j_stop = j;
      if (j < 0 && sl.first < stage_lo) {  // ran off the staged range while earlier rows still reach qs
        cont = true;
        g_line = stage_lo - 1u;
      }
    } else {
      cont = true;
      g_line = sl.line;
    }
    if (cont) cnt += walk_global(iv, g_line, sl.first, qs, qe, nothing);
  }
  if (i < n) cnt_out[i] = cnt;  // rle_right (interval_join.rs:1604)

  const uint32_t cincl = warp_incl_sum(cnt);
  if (lane == 31) s_wtot[warp] = cincl;
  __syncthreads();
  unsigned long long cta_tot = 0;
  uint32_t my_off = cincl - cnt;  // my first pair inside the CTA's output run
#pragma unroll
  for (int w = 0; w < kSWarps; ++w) {
    if (w < warp) my_off += uint32_t(s_wtot[w]);
    cta_tot += s_wtot[w];
  }
  if (!EMIT) {  // count only: the grand total is order-free
    if (tid == 0 && cta_tot) atomicAdd(result, cta_tot);
    return;
  }

  // ---- phase 5: chained scan over tiles -------------------------------------------------------------------------
  if (warp == 0) {
    const unsigned long long excl = chain_lookback(chain_state, tile, cta_tot, backoff_ns);
    if (lane == 0) {
      s_base = excl;
      if (tile == n_tiles - 1) result[0] = excl + cta_tot;
      if (excl + cta_tot > capacity) result[1] = 1;  // the caller's buffers are too small: report, write nothing here
    }
  }
  __syncthreads();
  const unsigned long long base = s_base;
  if (cta_tot == 0 || base + cta_tot > capacity) return;  // CTA-uniform

  // ---- phase 6: hits -> shared-memory window in output order -> coalesced stores ---------------------------------
  uint32_t* __restrict__ lout = left_out + base;
  uint32_t* __restrict__ rout = WRITE_RIGHT ? right_out + base : nullptr;
  const uint32_t tile_first = tile * kSB;
  const uint32_t T = uint32_t(cta_tot);
  // The window is laid out with the output's own misalignment (slot = position + a, a = output index mod 4), so its
  // body leaves as 16-byte stores: 2 instructions per 4 pairs instead of 8.
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(left_out) | (WRITE_RIGHT ? reinterpret_cast<uintptr_t>(right_out) : 0)) & 15u) == 0;
  for (uint32_t w0 = 0; w0 < T; w0 += kOutWin) {
    const uint32_t a = vec_ok ? uint32_t((base + w0) & 3ull) : 0u;
    if (cnt && my_off < w0 + kOutWin && my_off + cnt > w0) {
      uint32_t k = my_off - w0;  // window position of my next pair (wraps below zero for pairs of earlier windows)
      if (mode == 1) {
        if (my_off >= w0 && k + cnt <= kOutWin && j_hi - j_stop <= 32 && !cont) {
          // the common case: every hit is in the mask and in this window
          uint32_t mk = mask;
          uint32_t o = a + k;
          while (mk) {
            const int b = __ffs(int(mk)) - 1;
            mk &= mk - 1u;
            s_out[o] = s_raw[j_hi - b].y;
            s_own[o] = uint8_t(tid);
            ++o;
          }
        } else {
          for (int j = j_hi; j > j_stop; --j)
            if (s_er[j].x >= qs) {
              if (k < kOutWin) { s_out[a + k] = s_raw[j].y; s_own[a + k] = uint8_t(tid); }
              ++k;
            }
        }
      }
      if (cont) {
        auto push = [&](uint32_t id) {
          if (k < kOutWin) { s_out[a + k] = id; s_own[a + k] = uint8_t(tid); }
          ++k;
        };
        walk_global(iv, g_line, sl.first, qs, qe, push);
      }
    }
    __syncthreads();
    const uint32_t m = min(kOutWin, T - w0);
    if (vec_ok) {
      // window slots [a, a + m) <-> output elements G0 + slot, G0 = (base + w0) & ~3 (16-byte aligned)
      uint32_t* __restrict__ lg = left_out + ((base + w0) & ~3ull);
      uint32_t* __restrict__ rg = WRITE_RIGHT ? right_out + ((base + w0) & ~3ull) : nullptr;
      const uint32_t lo4 = (a + 3u) >> 2, hi4 = (a + m) >> 2;  // whole 4-slot groups
      for (uint32_t v = lo4 + tid; v < hi4; v += kSB) {
        reinterpret_cast<uint4*>(lg)[v] = reinterpret_cast<const uint4*>(s_out)[v];
        if (WRITE_RIGHT) {
          const uchar4 o = reinterpret_cast<const uchar4*>(s_own)[v];
          reinterpret_cast<uint4*>(rg)[v] = make_uint4(tile_first + o.x, tile_first + o.y, tile_first + o.z, tile_first + o.w);
        }
      }
      // ragged head and tail (at most 3 slots each)
      if (tid < 8) {
        const uint32_t head_end = min(lo4 * 4u, a + m), tail_begin = max(hi4 * 4u, head_end);
        const uint32_t sidx = tid < 4 ? a + tid : tail_begin + (tid - 4);
        const bool on = tid < 4 ? sidx < head_end : sidx < a + m;
        if (on) {
          lg[sidx] = s_out[sidx];
          if (WRITE_RIGHT) rg[sidx] = tile_first + s_own[sidx];
        }
      }
    } else {
      for (uint32_t t = tid; t < m; t += kSB) {
        lout[w0 + t] = s_out[t];
        if (WRITE_RIGHT) rout[w0 + t] = tile_first + s_own[t];
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
int launch_staged(sq_stream* s, const sq_index* idx, const uint64_t* d_key, const int32_t* d_start,
                  const int32_t* d_end, uint32_t n, uint32_t* d_left, uint32_t* d_right, uint64_t capacity) {
  ErrorSlot& E = s->err;
  const uint32_t n_tiles = (n + kSB - 1) / kSB;
  int rc;
  if ((rc = ensure(E, s->d_cnt, size_t(n) * 4, false))) return rc;
  if ((rc = ensure(E, s->d_tile, size_t(n_tiles) * 8 + 16, false))) return rc;
  if ((rc = ensure(E, s->d_scalar, 256, false))) return rc;
  auto* chain = static_cast<unsigned long long*>(s->d_tile.p);
  auto* ticket = reinterpret_cast<unsigned int*>(chain + n_tiles);
  auto* result = static_cast<unsigned long long*>(s->d_scalar.p);  // [0] n_pairs [1] overflow [2] CTAs on the global path
  auto* cnt = static_cast<uint32_t*>(s->d_cnt.p);
  SQ_CUDA(E, cudaMemsetAsync(result, 0, 32, s->stream));
  if (d_left) SQ_CUDA(E, cudaMemsetAsync(chain, 0, size_t(n_tiles) * 8 + 16, s->stream));
  const IndexView iv = idx->view();
  const uint32_t backoff = uint32_t(s->ctx->opt.lookback_backoff_ns.load(std::memory_order_relaxed));
  if (!d_left)
    k_probe_staged<false, false><<<n_tiles, kSB, 0, s->stream>>>(iv, d_key, d_start, d_end, n, cnt, chain, ticket, result, nullptr,
                                                                 nullptr, 0, n_tiles, 0u);
  else if (d_right)
    k_probe_staged<true, true><<<n_tiles, kSB, 0, s->stream>>>(iv, d_key, d_start, d_end, n, cnt, chain, ticket, result, d_left,
                                                               d_right, capacity, n_tiles, backoff);
  else
    k_probe_staged<true, false><<<n_tiles, kSB, 0, s->stream>>>(iv, d_key, d_start, d_end, n, cnt, chain, ticket, result, d_left,
                                                                nullptr, capacity, n_tiles, backoff);
  SQ_CUDA(E, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

uint32_t staged_tile_rows() { return kSB; }

// ---------------------------------------------------------------------------------------------
// Is a device tile position-ordered?  m <= 1024 adjacent row pairs spread over the tile.
__global__ void __launch_bounds__(1024) k_order_sample(const uint64_t* __restrict__ q_key, const int32_t* __restrict__ q_start,
                                                       uint32_t m, uint32_t stride, uint32_t* __restrict__ out2) {
  bool ok = false;
  if (threadIdx.x < m) {
    const size_t a = size_t(threadIdx.x) * stride;
    ok = q_key[a] == q_key[a + 1] && q_start[a] <= q_start[a + 1];
  }
  const int c = __syncthreads_count(ok);
  if (threadIdx.x == 0) { out2[0] = uint32_t(c); out2[1] = m; }
}

int sample_order_device(sq_stream* s, const uint64_t* d_key, const int32_t* d_start, uint32_t n, uint32_t out2[2]) {
  ErrorSlot& E = s->err;
  int rc;
  if ((rc = ensure(E, s->d_scalar, 256, false))) return rc;
  if ((rc = ensure(E, s->h_scalar, 256, true))) return rc;
  auto* d = reinterpret_cast<uint32_t*>(static_cast<char*>(s->d_scalar.p) + 128);
  auto* h = reinterpret_cast<uint32_t*>(static_cast<char*>(s->h_scalar.p) + 128);
  const uint32_t m = n - 1 < 1024u ? n - 1 : 1024u;
  const uint32_t stride = (n - 1) / m;
  k_order_sample<<<1, 1024, 0, s->stream>>>(d_key, d_start, m, stride, d);
  SQ_CUDA(E, cudaGetLastError());
  SQ_CUDA(E, cudaMemcpyAsync(h, d, 8, cudaMemcpyDeviceToHost, s->stream));
  SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  out2[0] = h[0];
  out2[1] = h[1];
  s->launches += 1;
  return SQ_OK;
}

}  // namespace sq
