// Scan side of the join (include/sequila_scan.h): delimited text (BED / CSV) -> the device columns the
// join consumes, parsed on the device.  Takes over what DataFusion's CSV reader + create_hashes +
// evaluate_as_i32 do on the host in the reference (queries/q1-coitrees.sql:6-14, interval_join.rs:1037,
// 1211, 1661-1672).
//
//   k_scan_count   every thread owns 64 bytes of text: newline mask (4 bytes per instruction), row
//                  starts = bytes behind a newline that do not open an empty / comment row; per 16 KB
//                  tile the number of row starts
//   (exclusive scan of the tile counts: launch_scan_u64)
//   k_scan_parse   the tile's text is staged in shared memory on the way to the same row starts, which a
//                  block scan compacts into a row list; the CTA's threads take the rows round-robin
//                  (neighbouring lanes = neighbouring rows: shared-memory bytes in, coalesced column
//                  stores out), parse field by field (key bytes -> key hash, BIGINT fields -> checked
//                  Int32) and note (hash -> first text offset) of the key in a per-CTA shared-memory
//                  table that is flushed to a small global table once per CTA
//   k_scan_ids     dictionary id per row (ids = order of first occurrence, assigned on the host from
//                  the table, a few thousand entries at most for genomes)
// Byte work bounded by HBM: the text is read twice (count, parse), 16-20 bytes per row are written.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "sequila_scan.h"
#include "sq_internal.cuh"
#include "sq_keyhash.h"
#include "sq_scan_row.h"

#define SQ_API extern "C" __attribute__((visibility("default")))

struct sq_scan {
  sq_ctx* ctx = nullptr;
  uint64_t n_rows = 0;
  uint64_t* d_key = nullptr;
  int32_t* d_start = nullptr;
  int32_t* d_end = nullptr;
  uint32_t* d_ids = nullptr;
  uint64_t bytes = 0;
  std::vector<std::string> dict;
  std::vector<uint64_t> dict_hash;
  float ms[3] = {0.f, 0.f, 0.f};
};

namespace sq {

constexpr int kScanBlock = 256;
constexpr int kScanChunk = 64;  // bytes of text per thread
constexpr uint32_t kScanTile = uint32_t(kScanBlock) * kScanChunk;  // 16 KB of text per CTA
constexpr unsigned long long kNoError = ~0ull;
constexpr int kLocalDict = 64;  // slots of the per-CTA key table

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t* s_w, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < kScanBlock / 32 ? s_w[lane] : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += o;
    }
    s_w[lane] = w;
  }
  __syncthreads();
  *total = s_w[31];
  return (warp ? s_w[warp - 1] : 0u) + inc - v;
}

__device__ __forceinline__ uint32_t newline_bits(const uint4& a, const uint4& b) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t eq = __vcmpeq4(w[k], 0x0a0a0a0au) & 0x01010101u;  // one flag bit per matching byte
    m |= ((eq * 0x01020408u) >> 24) << (4 * k);                      // the four flags side by side
  }
  return m;
}

// The thread's 64 bytes of text (bytes behind the end of the text read as '\n'): the two row-start masks of its
// halves; `keep`, when given, receives the bytes (the tile staged in shared memory).
__device__ __forceinline__ void chunk_row_starts(const uint8_t* __restrict__ text, uint64_t n, uint64_t q0, uint8_t comment,
                                                 uint4* keep, uint32_t* m_lo, uint32_t* m_hi) {
  uint4 v[4];
  if (q0 + kScanChunk <= n) {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __ldg(reinterpret_cast<const uint4*>(text + q0) + k);
  } else {
    uint8_t* bytes = reinterpret_cast<uint8_t*>(v);
    for (int k = 0; k < kScanChunk; ++k) bytes[k] = q0 + k < n ? __ldg(text + q0 + k) : uint8_t('\n');
  }
  if (keep) {
#pragma unroll
    for (int k = 0; k < 4; ++k) keep[k] = v[k];
  }
  *m_lo = *m_hi = 0u;
  if (q0 >= n) return;
  const TextSrc src{text, n};
  const uint32_t nl_lo = newline_bits(v[0], v[1]), nl_hi = newline_bits(v[2], v[3]);
  const uint64_t left = n - q0;
  const bool prev_nl = q0 == 0 || __ldg(text + q0 - 1) == '\n';
  *m_lo = row_starts_from_newlines(src, q0, nl_lo, prev_nl, left >= 32 ? 32u : uint32_t(left), comment);
  if (left > 32) *m_hi = row_starts_from_newlines(src, q0 + 32, nl_hi, (nl_lo >> 31) != 0u, left >= 64 ? 32u : uint32_t(left - 32), comment);
}

__global__ void __launch_bounds__(kScanBlock) k_scan_count(const uint8_t* __restrict__ text, uint64_t n, uint8_t comment,
                                                           unsigned long long* __restrict__ tile_rows) {
  __shared__ uint32_t s_w[32];
  const uint64_t q0 = uint64_t(blockIdx.x) * kScanTile + uint64_t(threadIdx.x) * kScanChunk;
  uint32_t m_lo, m_hi;
  chunk_row_starts(text, n, q0, comment, nullptr, &m_lo, &m_hi);
  uint32_t total;
  block_excl_scan_u32(__popc(m_lo) + __popc(m_hi), s_w, &total);
  if (threadIdx.x == 0) tile_rows[blockIdx.x] = total;
}

// hash -> first text offset of the key field (<< 16 | length), open addressing, slot `cap` = the key hash
// that equals the empty marker
struct DictTable {
  unsigned long long* keys;
  unsigned long long* vals;
  uint32_t mask;
  unsigned int* n_distinct;
  unsigned int* overflow;
};

__device__ __forceinline__ void dict_note(const DictTable& t, uint64_t key, unsigned long long v) {
  uint32_t slot;
  if (key == kEmptyKey) {
    slot = t.mask + 1u;
  } else {
    slot = uint32_t(mix64(key)) & t.mask;
    uint32_t step = 0;
    for (;; ++step) {
      if (step > t.mask) { *t.overflow = 1u; return; }
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(t.keys + slot);
      if (cur == kEmptyKey) {
        cur = atomicCAS(t.keys + slot, (unsigned long long)kEmptyKey, (unsigned long long)key);
        if (cur == kEmptyKey) {
          if (atomicAdd(t.n_distinct, 1u) + 1u > (t.mask + 1u) / 2u) *t.overflow = 1u;
          break;
        }
      }
      if (cur == key) break;
      slot = (slot + 1u) & t.mask;
    }
  }
  if (*reinterpret_cast<volatile unsigned long long*>(t.vals + slot) > v) atomicMin(t.vals + slot, v);
}

__device__ __forceinline__ uint32_t dict_find(const DictTable& t, uint64_t key) {
  if (key == kEmptyKey) return uint32_t(t.vals[t.mask + 1u]);
  uint32_t slot = uint32_t(mix64(key)) & t.mask;
  for (uint32_t step = 0; step <= t.mask; ++step) {
    const unsigned long long cur = t.keys[slot];
    if (cur == key) return uint32_t(t.vals[slot]);
    if (cur == kEmptyKey) break;
    slot = (slot + 1u) & t.mask;
  }
  return 0xFFFFFFFFu;
}

// per-CTA key table in shared memory: the rows of one tile carry a handful of distinct keys (one, in a sorted
// BED file), so the global table sees a few operations per CTA instead of one per row — every row of the text
// reading the same 24 table words made those L2 lines the hot spot of the kernel
__device__ __forceinline__ bool local_note(unsigned long long* s_keys, unsigned long long* s_vals, uint64_t key,
                                           unsigned long long v) {
  if (key == kEmptyKey) return false;
  uint32_t slot = uint32_t(key >> 17) & (kLocalDict - 1);
  for (int step = 0; step < kLocalDict; ++step) {
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(s_keys + slot);
    if (cur == kEmptyKey) cur = atomicCAS(s_keys + slot, (unsigned long long)kEmptyKey, (unsigned long long)key);
    if (cur == kEmptyKey || cur == key) {
      if (*reinterpret_cast<volatile unsigned long long*>(s_vals + slot) > v) atomicMin(s_vals + slot, v);
      return true;
    }
    slot = (slot + 1u) & (kLocalDict - 1);
  }
  return false;  // more distinct keys in this tile than slots: the caller goes to the global table
}

// Locate + parse: the tile's text goes to shared memory on the way to the newline masks, the row starts of the
// tile are compacted into a list (block scan), and the CTA's threads take the rows of that list round-robin:
// neighbouring lanes parse neighbouring rows (shared-memory bytes, coalesced column stores), whatever the
// row lengths.
__global__ void __launch_bounds__(kScanBlock)
k_scan_parse(const uint8_t* __restrict__ text, uint64_t n, ScanOpts o, const unsigned long long* __restrict__ tile_base,
             uint64_t n_rows, uint64_t* __restrict__ key_out, int32_t* __restrict__ start_out,
             int32_t* __restrict__ end_out, DictTable dict, unsigned long long* err_off) {
  __shared__ uint4 s_text[kScanTile / 16];
  __shared__ uint16_t s_row[kScanTile / 2];  // a row is at least one byte and its newline
  __shared__ uint32_t s_w[32];
  __shared__ unsigned long long s_dkeys[kLocalDict], s_dvals[kLocalDict];
  const uint64_t tile0 = uint64_t(blockIdx.x) * kScanTile;
  const uint64_t q0 = tile0 + uint64_t(threadIdx.x) * kScanChunk;
  if (threadIdx.x < kLocalDict) {
    s_dkeys[threadIdx.x] = kEmptyKey;
    s_dvals[threadIdx.x] = ~0ull;
  }
  uint32_t m_lo, m_hi;
  chunk_row_starts(text, n, q0, o.comment, s_text + threadIdx.x * (kScanChunk / 16), &m_lo, &m_hi);
  uint32_t total;
  uint32_t at = block_excl_scan_u32(__popc(m_lo) + __popc(m_hi), s_w, &total);
  while (m_lo) {
    s_row[at++] = uint16_t(threadIdx.x * kScanChunk + (__ffs(m_lo) - 1));
    m_lo &= m_lo - 1;
  }
  while (m_hi) {
    s_row[at++] = uint16_t(threadIdx.x * kScanChunk + 32 + (__ffs(m_hi) - 1));
    m_hi &= m_hi - 1;
  }
  __syncthreads();
  const TileSrc src{reinterpret_cast<const uint8_t*>(s_text), uint32_t(__cvta_generic_to_shared(s_text)), tile0, kScanTile,
                    TextSrc{text, n}};
  const uint64_t first = tile_base[blockIdx.x];
  for (uint32_t k = threadIdx.x; k < total; k += kScanBlock) {
    uint64_t r = first + k;
    if (o.has_header) {
      if (r == 0) continue;  // the header row
      r -= 1;
    }
    if (r >= n_rows) continue;
    const uint64_t q = tile0 + s_row[k];
    // every row but the tile's last ends (newline included) in front of the next row start, i.e. inside the tile
    const RowResult res = k + 1 < total ? parse_row(SmemSrc{src.tile_saddr, tile0}, uint32_t(s_row[k]), o)
                                        : parse_row(src, uint32_t(s_row[k]), o);
    if (res.err != kRowOk) {
      atomicMin(err_off, (unsigned long long)q);  // offsets grow with the row number: min = first bad row
      continue;
    }
    key_out[r] = res.key;
    start_out[r] = res.start;
    end_out[r] = res.end;
    if (o.col_key >= 0) {
      const unsigned long long v = ((unsigned long long)res.key_off << 16) | res.key_len;
      if (!local_note(s_dkeys, s_dvals, res.key, v)) dict_note(dict, res.key, v);
    }
  }
  if (o.col_key < 0) return;
  __syncthreads();
  if (threadIdx.x < kLocalDict && s_dkeys[threadIdx.x] != kEmptyKey) dict_note(dict, s_dkeys[threadIdx.x], s_dvals[threadIdx.x]);
}

__global__ void __launch_bounds__(256) k_scan_ids(const uint64_t* __restrict__ key, uint64_t n_rows, DictTable dict,
                                                  uint32_t* __restrict__ ids) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_rows; i += stride) ids[i] = dict_find(dict, key[i]);
}

#endif  // __CUDACC__

namespace {

struct DeviceTemp {  // frees on scope exit
  void* p = nullptr;
  ~DeviceTemp() { if (p) cudaFree(p); }
};

void free_scan(sq_scan* sc) {
  if (!sc) return;
  cudaFree(sc->d_key);
  cudaFree(sc->d_start);
  cudaFree(sc->d_end);
  cudaFree(sc->d_ids);
  delete sc;
}

std::string printable(const uint8_t* p, size_t n) {
  std::string s;
  for (size_t i = 0; i < n && i < 120; ++i) {
    const uint8_t c = p[i];
    if (c == '\n') break;
    if (c == '\t') s += "\\t";
    else if (c == '\r') s += "\\r";
    else if (c < 32 || c > 126) s += '?';
    else s += char(c);
  }
  return s;
}

// the first failing row: copy it back, parse it again on the host, word the error
int diagnose(sq_stream* s, const uint8_t* d_text, uint64_t n, uint64_t off, const ScanOpts& o) {
  ErrorSlot& E = s->err;
  const uint64_t len = std::min<uint64_t>(n - off, 1ull << 16);
  std::vector<uint8_t> win(len);
  SQ_CUDA(E, cudaMemcpyAsync(win.data(), d_text + off, len, cudaMemcpyDeviceToHost, s->stream));
  SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  const RowResult r = parse_row(TextSrc{win.data(), len}, 0, o);
  const std::string row = printable(win.data(), len);
  switch (r.err) {
    case kRowCastStart:
    case kRowCastEnd:  // exact text of the reference's failure (interval_join.rs:1959-1965)
      return fail(E, SQ_ECAST, "Arrow error: Cast error: Can't cast value %lld to type Int32", (long long)r.bad_value);
    case kRowFewFields:
      return fail(E, SQ_EPARSE, "row at byte %llu has %d field(s), the table needs field %d: '%s'", (unsigned long long)off,
                  r.fields, std::max(o.col_key, std::max(o.col_start, o.col_end)), row.c_str());
    case kRowBadStart:
    case kRowBadEnd:
      return fail(E, SQ_EPARSE, "row at byte %llu: field %d is not a BIGINT: '%s'", (unsigned long long)off,
                  r.err == kRowBadStart ? o.col_start : o.col_end, row.c_str());
    case kRowQuoted:
      return fail(E, SQ_EPARSE, "row at byte %llu: quoted fields are not supported: '%s'", (unsigned long long)off, row.c_str());
    case kRowKeyTooLong:
      return fail(E, SQ_EPARSE, "row at byte %llu: key field longer than 65534 bytes", (unsigned long long)off);
    default:
      return fail(E, SQ_EPARSE, "row at byte %llu could not be parsed: '%s'", (unsigned long long)off, row.c_str());
  }
}

int scan_device(sq_stream* s, const uint8_t* d_text, const uint8_t* h_text, uint64_t n, const sq_scan_options* opt,
                cudaEvent_t ev_begin, sq_scan** out) {
  ErrorSlot& E = s->err;
  ScanOpts o;
  o.delim = opt->delimiter;
  o.comment = opt->comment;
  o.has_header = opt->has_header != 0;
  o.col_key = opt->col_key;
  o.col_start = opt->col_start;
  o.col_end = opt->col_end;
  o.start_minus = opt->start_minus;
  o.end_minus = opt->end_minus;
  if (o.col_start < 0 || o.col_end < 0) return fail(E, SQ_EINVAL, "col_start and col_end must name a field");
  if (o.delim == '\n' || o.delim == '\r' || o.delim == 0) return fail(E, SQ_EINVAL, "bad delimiter");
  if (o.col_key >= 0 && (o.col_key == o.col_start || o.col_key == o.col_end))
    return fail(E, SQ_EINVAL, "the key column cannot also be an interval column");
  if (reinterpret_cast<uintptr_t>(d_text) & 15u) return fail(E, SQ_EINVAL, "device text must be 16-byte aligned");

  // [0,1] locate (count + tile scan)  [2,3] parse  [4,5] id kernel: kernels only, allocations stay outside
  cudaEvent_t ev[6];
  for (auto& e : ev) SQ_CUDA(E, cudaEventCreate(&e));
  struct EvGuard { cudaEvent_t* e; ~EvGuard() { for (int i = 0; i < 6; ++i) cudaEventDestroy(e[i]); } } evg{ev};

  auto* sc = new sq_scan();
  sc->ctx = s->ctx;
  struct ScanGuard { sq_scan* p; ~ScanGuard() { if (p) free_scan(p); } } guard{sc};

  const uint64_t n_tiles64 = (n + kScanTile - 1) / kScanTile;
  if (n_tiles64 > 0x7FFFFFFFull) return fail(E, SQ_EINVAL, "text too large for one scan (%llu bytes)", (unsigned long long)n);
  const uint32_t n_tiles = uint32_t(n_tiles64);
  if (ev_begin) SQ_CUDA(E, cudaEventRecord(ev[0], s->stream));
  uint64_t total_rows = 0;
  DeviceTemp tiles, tot;
  if (n_tiles) {
    SQ_CUDA(E, cudaMalloc(&tiles.p, size_t(n_tiles) * 8));
    SQ_CUDA(E, cudaMalloc(&tot.p, 64));
    auto* d_tiles = static_cast<unsigned long long*>(tiles.p);
    auto* d_tot = static_cast<unsigned long long*>(tot.p);
    float h2d = 0.f;
    if (ev_begin) {  // the text's host-to-device copy is over here
      SQ_CUDA(E, cudaEventSynchronize(ev[0]));
      cudaEventElapsedTime(&h2d, ev_begin, ev[0]);
    }
    sc->ms[0] = h2d;
    SQ_CUDA(E, cudaEventRecord(ev[0], s->stream));
    k_scan_count<<<n_tiles, kScanBlock, 0, s->stream>>>(d_text, n, o.comment, d_tiles);
    SQ_CUDA(E, cudaGetLastError());
    s->launches += 1;
    int rc;
    if ((rc = launch_scan_u64(s, d_tiles, n_tiles, d_tot))) return rc;
    SQ_CUDA(E, cudaEventRecord(ev[1], s->stream));
    SQ_CUDA(E, cudaMemcpyAsync(&total_rows, d_tot, 8, cudaMemcpyDeviceToHost, s->stream));
    SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  }
  const uint64_t n_rows = total_rows - ((o.has_header && total_rows) ? 1 : 0);
  if (n_rows >= 0xFFFFFFFFull) return fail(E, SQ_EINVAL, "%llu rows: a table side must stay below 2^32 - 1 rows", (unsigned long long)n_rows);
  sc->n_rows = n_rows;
  if (n_rows) {
    SQ_CUDA(E, cudaMalloc(&sc->d_key, n_rows * 8));
    SQ_CUDA(E, cudaMalloc(&sc->d_start, n_rows * 4));
    SQ_CUDA(E, cudaMalloc(&sc->d_end, n_rows * 4));
    sc->bytes = n_rows * 16;
    if (o.col_key >= 0) {
      SQ_CUDA(E, cudaMalloc(&sc->d_ids, n_rows * 4));
      sc->bytes += n_rows * 4;
    }
  }

  std::vector<unsigned long long> h_keys, h_vals;
  DeviceTemp tab, flags;
  DictTable dict{};
  SQ_CUDA(E, cudaMalloc(&flags.p, 64));
  auto* d_flags = static_cast<unsigned long long*>(flags.p);  // [0] first bad offset, [1] n_distinct (u32), [2] overflow (u32)
  // option cuda_scan_dict_capacity: tests start small to exercise the growth path
  uint32_t cap = uint32_t(s->ctx->opt.scan_dict_capacity.load(std::memory_order_relaxed));
  for (; n_rows;) {
    if (tab.p) { cudaFree(tab.p); tab.p = nullptr; }
    SQ_CUDA(E, cudaMalloc(&tab.p, size_t(cap + 1) * 16));
    SQ_CUDA(E, cudaMemsetAsync(tab.p, 0xFF, size_t(cap + 1) * 16, s->stream));
    SQ_CUDA(E, cudaMemsetAsync(d_flags, 0xFF, 8, s->stream));
    SQ_CUDA(E, cudaMemsetAsync(d_flags + 1, 0, 16, s->stream));
    dict.keys = static_cast<unsigned long long*>(tab.p);
    dict.vals = dict.keys + (cap + 1);
    dict.mask = cap - 1;
    dict.n_distinct = reinterpret_cast<unsigned int*>(d_flags + 1);
    dict.overflow = reinterpret_cast<unsigned int*>(d_flags + 2);
    SQ_CUDA(E, cudaEventRecord(ev[2], s->stream));
    k_scan_parse<<<n_tiles, kScanBlock, 0, s->stream>>>(d_text, n, o, static_cast<unsigned long long*>(tiles.p), n_rows,
                                                         sc->d_key, sc->d_start, sc->d_end, dict, d_flags);
    SQ_CUDA(E, cudaGetLastError());
    SQ_CUDA(E, cudaEventRecord(ev[3], s->stream));
    s->launches += 1;
    unsigned long long h_flags[3];
    SQ_CUDA(E, cudaMemcpyAsync(h_flags, d_flags, 24, cudaMemcpyDeviceToHost, s->stream));
    SQ_CUDA(E, cudaStreamSynchronize(s->stream));
    if (h_flags[0] != kNoError) return diagnose(s, d_text, n, h_flags[0], o);
    if (uint32_t(h_flags[2]) == 0) break;
    if (cap >= (1u << 28)) return fail(E, SQ_EINVAL, "more than 2^27 distinct keys");
    cap <<= 4;  // the key column has more distinct values than the table holds: again with a larger one
  }
  bool ids_timed = false;
  if (n_rows && o.col_key >= 0) {
    h_keys.resize(cap + 1);
    h_vals.resize(cap + 1);
    SQ_CUDA(E, cudaMemcpyAsync(h_keys.data(), dict.keys, size_t(cap + 1) * 8, cudaMemcpyDeviceToHost, s->stream));
    SQ_CUDA(E, cudaMemcpyAsync(h_vals.data(), dict.vals, size_t(cap + 1) * 8, cudaMemcpyDeviceToHost, s->stream));
    SQ_CUDA(E, cudaStreamSynchronize(s->stream));
    struct Ent { unsigned long long val; uint32_t slot; };
    std::vector<Ent> ents;
    for (uint32_t i = 0; i <= cap; ++i)
      if (h_vals[i] != ~0ull) ents.push_back({h_vals[i], i});
    std::sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.val < b.val; });  // first occurrence order
    sc->dict.reserve(ents.size());
    for (uint32_t id = 0; id < ents.size(); ++id) {
      const uint64_t off = ents[id].val >> 16;
      const uint32_t len = uint32_t(ents[id].val & 0xFFFFull);
      std::string str(len, '\0');
      if (len) {
        if (h_text) memcpy(&str[0], h_text + off, len);
        else SQ_CUDA(E, cudaMemcpy(&str[0], d_text + off, len, cudaMemcpyDeviceToHost));
      }
      sc->dict.push_back(std::move(str));
      sc->dict_hash.push_back(ents[id].slot == cap ? kEmptyKey : h_keys[ents[id].slot]);
      h_vals[ents[id].slot] = id;
    }
    SQ_CUDA(E, cudaMemcpyAsync(dict.vals, h_vals.data(), size_t(cap + 1) * 8, cudaMemcpyHostToDevice, s->stream));
    const uint64_t want = (n_rows + 255) / 256;
    const uint64_t gcap = uint64_t(s->ctx->sm_count) * 16;
    SQ_CUDA(E, cudaEventRecord(ev[4], s->stream));
    k_scan_ids<<<int(std::min(want, gcap)), 256, 0, s->stream>>>(sc->d_key, n_rows, dict, sc->d_ids);
    SQ_CUDA(E, cudaGetLastError());
    SQ_CUDA(E, cudaEventRecord(ev[5], s->stream));
    s->launches += 1;
    ids_timed = true;
  }
  SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  if (n_rows) {
    float locate = 0.f, parse = 0.f;
    cudaEventElapsedTime(&locate, ev[0], ev[1]);
    cudaEventElapsedTime(&parse, ev[2], ev[3]);
    sc->ms[1] = locate + parse;
    if (ids_timed) cudaEventElapsedTime(&sc->ms[2], ev[4], ev[5]);
  }
  guard.p = nullptr;
  *out = sc;
  return SQ_OK;
}

// Pageable host text -> device.  A plain cudaMemcpyAsync from pageable memory is staged by the driver on one
// thread (measured 11 GB/s on the B200 box: 43 of the 46 ms a 477 MB scan took); here kThreads host threads
// copy 4 MB chunks into a ring of pinned buffers (two per thread) and queue the DMA of each chunk themselves,
// so the PCIe copy runs beside the host-side memcpy of the next chunks.
int copy_text_to_device(sq_stream* s, uint8_t* d_text, const uint8_t* text, uint64_t n) {
  ErrorSlot& E = s->err;
  constexpr uint64_t kChunk = 4ull << 20;
  constexpr int kThreads = 4, kSlots = 2;
  if (n < 4 * kChunk) {
    if (n) SQ_CUDA(E, cudaMemcpyAsync(d_text, text, n, cudaMemcpyHostToDevice, s->stream));
    return SQ_OK;
  }
  int rc;
  if ((rc = ensure(E, s->h_scan, kChunk * kThreads * kSlots, true))) return rc;
  auto* ring = static_cast<uint8_t*>(s->h_scan.p);
  cudaEvent_t done[kThreads][kSlots] = {};
  struct Events {  // destroyed on every way out
    cudaEvent_t (&ev)[kThreads][kSlots];
    ~Events() { for (auto& t : ev) for (auto& e : t) if (e) cudaEventDestroy(e); }
  } events{done};
  for (auto& t : done)
    for (auto& e : t) SQ_CUDA(E, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  const uint64_t n_chunks = (n + kChunk - 1) / kChunk;
  std::atomic<int> failed{0};
  const int device = s->ctx->device;
  auto worker = [&](int t) {
    if (cudaSetDevice(device) != cudaSuccess) { failed = 1; return; }
    uint64_t round = 0;
    for (uint64_t c = uint64_t(t); c < n_chunks && !failed; c += kThreads, ++round) {
      const int slot = int(round % kSlots);
      uint8_t* pin = ring + (uint64_t(t) * kSlots + slot) * kChunk;
      if (round >= kSlots && cudaEventSynchronize(done[t][slot]) != cudaSuccess) { failed = 1; return; }
      const uint64_t off = c * kChunk, len = std::min(kChunk, n - off);
      memcpy(pin, text + off, len);
      if (cudaMemcpyAsync(d_text + off, pin, len, cudaMemcpyHostToDevice, s->stream) != cudaSuccess ||
          cudaEventRecord(done[t][slot], s->stream) != cudaSuccess) { failed = 1; return; }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < kThreads; ++t) th.emplace_back(worker, t);
  worker(0);
  for (auto& x : th) x.join();
  // the ring is reused by the next call: every DMA out of it must be over before this one returns
  cudaError_t e = cudaStreamSynchronize(s->stream);
  if (failed || e != cudaSuccess) {
    cudaGetLastError();
    return fail(E, SQ_ECUDA, "host-to-device copy of the text failed: %s", cudaGetErrorString(e));
  }
  return SQ_OK;
}

}  // namespace
}  // namespace sq

using namespace sq;

SQ_API int32_t sq_scan_text_device(sq_stream* s, const uint8_t* d_text, uint64_t n_bytes, const sq_scan_options* opt,
                                   sq_scan** out) {
  if (!s) return SQ_EINVAL;
  if (!opt || !out || (n_bytes && !d_text)) return fail(s->err, SQ_EINVAL, "null argument");
  SQ_CUDA(s->err, cudaSetDevice(s->ctx->device));
  return scan_device(s, d_text, nullptr, n_bytes, opt, nullptr, out);
}

SQ_API int32_t sq_scan_text(sq_stream* s, const uint8_t* text, uint64_t n_bytes, const sq_scan_options* opt, sq_scan** out) {
  if (!s) return SQ_EINVAL;
  ErrorSlot& E = s->err;
  if (!opt || !out || (n_bytes && !text)) return fail(E, SQ_EINVAL, "null argument");
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  DeviceTemp d;
  cudaEvent_t ev0;
  SQ_CUDA(E, cudaEventCreate(&ev0));
  struct G { cudaEvent_t e; ~G() { cudaEventDestroy(e); } } g{ev0};
  SQ_CUDA(E, cudaMalloc(&d.p, n_bytes + 64));
  SQ_CUDA(E, cudaEventRecord(ev0, s->stream));
  int rc;
  if ((rc = copy_text_to_device(s, static_cast<uint8_t*>(d.p), text, n_bytes))) return rc;
  return scan_device(s, static_cast<const uint8_t*>(d.p), text, n_bytes, opt, ev0, out);
}

SQ_API uint64_t sq_scan_rows(const sq_scan* sc) { return sc ? sc->n_rows : 0; }
SQ_API uint64_t sq_scan_bytes(const sq_scan* sc) { return sc ? sc->bytes : 0; }
SQ_API const uint64_t* sq_scan_key_hash_device(const sq_scan* sc) { return sc ? sc->d_key : nullptr; }
SQ_API const int32_t* sq_scan_start_device(const sq_scan* sc) { return sc ? sc->d_start : nullptr; }
SQ_API const int32_t* sq_scan_end_device(const sq_scan* sc) { return sc ? sc->d_end : nullptr; }
SQ_API const uint32_t* sq_scan_key_ids_device(const sq_scan* sc) { return sc ? sc->d_ids : nullptr; }
SQ_API uint32_t sq_scan_dict_size(const sq_scan* sc) { return sc ? uint32_t(sc->dict.size()) : 0u; }

SQ_API int32_t sq_scan_dict_entry(const sq_scan* sc, uint32_t id, const uint8_t** bytes, uint32_t* len, uint64_t* key_hash) {
  if (!sc || id >= sc->dict.size()) return SQ_EINVAL;
  if (bytes) *bytes = reinterpret_cast<const uint8_t*>(sc->dict[id].data());
  if (len) *len = uint32_t(sc->dict[id].size());
  if (key_hash) *key_hash = sc->dict_hash[id];
  return SQ_OK;
}

SQ_API int32_t sq_scan_fetch(sq_stream* s, const sq_scan* sc, uint64_t* key_hash, int32_t* start, int32_t* end,
                             uint32_t* key_ids) {
  if (!s) return SQ_EINVAL;
  ErrorSlot& E = s->err;
  if (!sc) return fail(E, SQ_EINVAL, "null scan");
  if (key_ids && !sc->d_ids && sc->n_rows) return fail(E, SQ_ESTATE, "the scan has no key column");
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  const uint64_t n = sc->n_rows;
  if (n) {
    if (key_hash) SQ_CUDA(E, cudaMemcpyAsync(key_hash, sc->d_key, n * 8, cudaMemcpyDeviceToHost, s->stream));
    if (start) SQ_CUDA(E, cudaMemcpyAsync(start, sc->d_start, n * 4, cudaMemcpyDeviceToHost, s->stream));
    if (end) SQ_CUDA(E, cudaMemcpyAsync(end, sc->d_end, n * 4, cudaMemcpyDeviceToHost, s->stream));
    if (key_ids) SQ_CUDA(E, cudaMemcpyAsync(key_ids, sc->d_ids, n * 4, cudaMemcpyDeviceToHost, s->stream));
    SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  }
  return SQ_OK;
}

SQ_API int32_t sq_scan_timing(const sq_scan* sc, float out3[3]) {
  if (!sc || !out3) return SQ_EINVAL;
  for (int i = 0; i < 3; ++i) out3[i] = sc->ms[i];
  return SQ_OK;
}

SQ_API void sq_scan_free(sq_scan* sc) {
  if (!sc) return;
  cudaSetDevice(sc->ctx->device);
  free_scan(sc);
}
