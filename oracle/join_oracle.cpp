// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
//
// CPU restatement of sequila-native's interval-join hot path with the default
// `Coitrees` algorithm, on the arguments the C ABI in include/sequila_cuda.h
// receives.  Reference: sequila/sequila-core/src/physical_planner/joins/interval_join.rs
// (abbreviated IJ).
//
//   orc_index_build   IJ:662-683   bucket rows by key *hash only* (IJ:1042-1048),
//                                  one coitrees tree per key (IJ:769-793)
//   orc_probe         IJ:1582-1618 per probe row: map lookup (miss => no rows,
//                                  IJ:965), tree.query, push `pos as u32`,
//                                  rle_right -> index_right expansion
//   orc_probe_counts  IJ:1604      the rle_right vector itself
//   orc_gather_i32    IJ:1620-1632 arrow::compute::take of one 4-byte column
//   orc_brute         independent O(Nb*Np) check of the predicate IV:95-137 / CT/nosimd.rs:647-649
//   orc_nearest       IJ:794-812 (build: tree + intervals sorted by (first, last)), IJ:909-956
//                     (`nearest`), IJ:972-990 (`get`: first overlap the tree reports, else nearest),
//                     IJ:1593-1602 (one output row per probe row, NULL left side on a key miss)
//   orc_time_probe    the same probe loop dealt to T threads in 8192-row
//                     batches over one shared index (= PartitionMode::CollectLeft
//                     with target_partitions = T, IJ:473-487), for the CPU baseline
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstdlib>
#include <thread>
#include <unordered_map>

#include "coitrees_port.hpp"

namespace {

using Clock = std::chrono::steady_clock;

struct IndexBase {
  virtual ~IndexBase() = default;
  virtual void probe_row(uint64_t key, int32_t first, int32_t last, std::vector<uint32_t>& out) const = 0;
  virtual size_t bytes() const = 0;
  size_t n_keys = 0;
  double build_seconds = 0;
};

template <int LANES>
struct Index final : IndexBase {
  std::unordered_map<uint64_t, orc::Tree<LANES>> trees;

  void probe_row(uint64_t key, int32_t first, int32_t last, std::vector<uint32_t>& out) const override {
    auto it = trees.find(key);
    if (it == trees.end()) return;  // contig absent on the build side => no rows (IJ:965)
    it->second.query(first, last, [&](uint64_t pos) { out.push_back(uint32_t(pos)); });  // IJ:1590
  }
  size_t bytes() const override {
    size_t b = 0;
    for (auto& kv : trees) b += kv.second.bytes();
    return b;
  }
};

template <int LANES>
IndexBase* build_index(const uint64_t* key, const int32_t* start, const int32_t* end, uint64_t n) {
  auto t0 = Clock::now();
  std::unordered_map<uint64_t, std::vector<orc::Interval>> buckets;  // IJ:662
  buckets.reserve(16);
  for (uint64_t i = 0; i < n; ++i) {
    auto ins = buckets.try_emplace(key[i]);
    if (ins.second) ins.first->second.reserve(4096);  // IJ:1046
    ins.first->second.push_back(orc::Interval{start[i], end[i], i});
  }
  auto* idx = new Index<LANES>();
  idx->trees.reserve(buckets.size());
  for (auto& kv : buckets) idx->trees.emplace(kv.first, orc::Tree<LANES>(std::move(kv.second)));
  idx->n_keys = idx->trees.size();
  idx->build_seconds = std::chrono::duration<double>(Clock::now() - t0).count();
  return idx;
}

}  // namespace

extern "C" {

// variant: 1 = scalar nodes (nosimd.rs), 8 = eight-interval chunk nodes (avx.rs)
void* orc_index_build(const uint64_t* key, const int32_t* start, const int32_t* end, uint64_t n,
                      int32_t variant) {
  if (variant == 8) return build_index<8>(key, start, end, n);
  return build_index<1>(key, start, end, n);
}

void orc_index_free(void* h) { delete static_cast<IndexBase*>(h); }
uint64_t orc_index_bytes(const void* h) { return static_cast<const IndexBase*>(h)->bytes(); }
uint64_t orc_index_keys(const void* h) { return static_cast<const IndexBase*>(h)->n_keys; }
double orc_index_build_seconds(const void* h) { return static_cast<const IndexBase*>(h)->build_seconds; }

// Full-mode probe of one batch.  *left_out / *right_out are malloc'ed (free with orc_free).
// right_out is non-decreasing (probe order preserved); counts_out (nullable) = rle_right.
int64_t orc_probe(const void* h, const uint64_t* key, const int32_t* start, const int32_t* end,
                  uint64_t n, uint32_t** left_out, uint32_t** right_out, uint32_t* counts_out) {
  const auto* idx = static_cast<const IndexBase*>(h);
  std::vector<uint32_t> left, hits;
  std::vector<uint32_t> rle(n);
  hits.reserve(256);  // IJ:1584
  for (uint64_t i = 0; i < n; ++i) {
    hits.clear();
    idx->probe_row(key[i], start[i], end[i], hits);
    rle[i] = uint32_t(hits.size());
    left.insert(left.end(), hits.begin(), hits.end());
  }
  const size_t np = left.size();
  auto* l = static_cast<uint32_t*>(std::malloc(std::max<size_t>(np, 1) * sizeof(uint32_t)));
  auto* r = static_cast<uint32_t*>(std::malloc(std::max<size_t>(np, 1) * sizeof(uint32_t)));
  std::memcpy(l, left.data(), np * sizeof(uint32_t));
  size_t o = 0;
  for (uint64_t i = 0; i < n; ++i)  // IJ:1611-1618
    for (uint32_t k = 0; k < rle[i]; ++k) r[o++] = uint32_t(i);
  if (counts_out) std::memcpy(counts_out, rle.data(), n * sizeof(uint32_t));
  *left_out = l;
  *right_out = r;
  return int64_t(np);
}

int64_t orc_probe_counts(const void* h, const uint64_t* key, const int32_t* start, const int32_t* end,
                         uint64_t n, uint32_t* counts_out) {
  const auto* idx = static_cast<const IndexBase*>(h);
  std::vector<uint32_t> hits;
  int64_t total = 0;
  for (uint64_t i = 0; i < n; ++i) {
    hits.clear();
    idx->probe_row(key[i], start[i], end[i], hits);
    counts_out[i] = uint32_t(hits.size());
    total += int64_t(hits.size());
  }
  return total;
}


// Algorithm::CoitreesNearest.  left_out[i] = build row chosen for probe row i, 0xFFFFFFFF = NULL.
// overlap_out[i] (nullable) = 1 when the row was an overlap (then WHICH overlapping row is reported
// is the tree's traversal order, unspecified by the reference: "return an arbitrary one", IJ:976).
void orc_nearest(const uint64_t* bkey, const int32_t* bstart, const int32_t* bend, uint64_t nb,
                 const uint64_t* pkey, const int32_t* pstart, const int32_t* pend, uint64_t np,
                 uint32_t* left_out, uint8_t* overlap_out) {
  struct PerKey {
    orc::Tree<8> tree;
    std::vector<orc::Interval> sorted;
  };
  std::unordered_map<uint64_t, std::vector<orc::Interval>> buckets;
  for (uint64_t i = 0; i < nb; ++i) buckets[bkey[i]].push_back(orc::Interval{bstart[i], bend[i], i});
  std::unordered_map<uint64_t, PerKey> map;
  for (auto& kv : buckets) {
    std::vector<orc::Interval> sorted = kv.second;  // IJ:803-806: stable sort by (first, last)
    std::stable_sort(sorted.begin(), sorted.end(), [](const orc::Interval& a, const orc::Interval& b) {
      return a.first != b.first ? a.first < b.first : a.last < b.last;
    });
    map.emplace(kv.first, PerKey{orc::Tree<8>(std::vector<orc::Interval>(kv.second)), std::move(sorted)});
  }
  for (uint64_t i = 0; i < np; ++i) {
    left_out[i] = 0xFFFFFFFFu;
    if (overlap_out) overlap_out[i] = 0;
    auto it = map.find(pkey[i]);
    if (it == map.end()) continue;  // IJ:974: no tree for the key => f never called => NULL (IJ:1597-1598)
    const int32_t start = pstart[i], end = pend[i];
    int seen = 0;
    it->second.tree.query(start, end, [&](uint64_t pos) {  // IJ:977-983: the first overlap reported
      if (seen == 0) { left_out[i] = uint32_t(pos); ++seen; }
    });
    if (seen) { if (overlap_out) overlap_out[i] = 1; continue; }
    const auto& r = it->second.sorted;  // IJ:909-956
    size_t left = 0, right = r.size();
    while (left < right) {
      const size_t mid = (left + right) / 2;
      if (r[mid].first < end) left = mid + 1; else right = mid;
    }
    int64_t min_distance = INT32_MAX;
    const size_t cand[2] = {left == 0 ? 0 : left - 1, left};
    for (size_t c : cand) {
      if (c >= r.size()) continue;
      int64_t d;
      if (end < r[c].first) d = int64_t(r[c].first) - end;
      else if (r[c].last < start) d = int64_t(start) - r[c].last;
      else d = 0;
      if (d < min_distance) { min_distance = d; left_out[i] = uint32_t(r[c].meta); }
    }
  }
}

void orc_free(void* p) { std::free(p); }

void orc_gather_i32(const int32_t* col, const uint32_t* idx, uint64_t n, int32_t* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = col[idx[i]];
}

// Independent brute force: every (b, p) with key equal and b.start <= p.end && b.end >= p.start,
// in probe order then build order.
int64_t orc_brute(const uint64_t* bkey, const int32_t* bstart, const int32_t* bend, uint64_t nb,
                  const uint64_t* pkey, const int32_t* pstart, const int32_t* pend, uint64_t np,
                  uint32_t** left_out, uint32_t** right_out) {
  std::vector<uint32_t> l, r;
  for (uint64_t p = 0; p < np; ++p)
    for (uint64_t b = 0; b < nb; ++b)
      if (bkey[b] == pkey[p] && bstart[b] <= pend[p] && bend[b] >= pstart[p]) {
        l.push_back(uint32_t(b));
        r.push_back(uint32_t(p));
      }
  auto* lo = static_cast<uint32_t*>(std::malloc(std::max<size_t>(l.size(), 1) * 4));
  auto* ro = static_cast<uint32_t*>(std::malloc(std::max<size_t>(r.size(), 1) * 4));
  std::memcpy(lo, l.data(), l.size() * 4);
  std::memcpy(ro, r.data(), r.size() * 4);
  *left_out = lo;
  *right_out = ro;
  return int64_t(l.size());
}

// Timed probe for the CPU baseline.  Batches of `batch_rows` probe rows are dealt
// round-robin to `threads` workers; each runs the IJ:1582-1618 loop (hits -> u32
// left indices + RLE -> right indices) and, when the six int32 columns are given
// (`materialise` != 0), the IJ:1620-1632 `take` of 3 build + 3 probe columns.
// Returns wall seconds of the probe phase; *pairs_out = total pairs; *digest_out =
// order-independent multiset digest of (left,right) with right made global.
static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27; x *= 0x94d049bb133111ebULL;
  x ^= x >> 31; return x;
}

static volatile uint64_t g_sink;

double orc_time_probe(const void* h, const uint64_t* key, const int32_t* start, const int32_t* end,
                      uint64_t n, int32_t threads, uint32_t batch_rows, int32_t materialise,
                      const int32_t* bcols[3], const int32_t* pcols[3], uint64_t* pairs_out,
                      uint64_t* digest_out) {
  const auto* idx = static_cast<const IndexBase*>(h);
  if (threads < 1) threads = 1;
  if (batch_rows == 0) batch_rows = 8192;
  const uint64_t nbatches = (n + batch_rows - 1) / batch_rows;
  std::vector<uint64_t> pairs(threads, 0), digest(threads, 0), sink(threads, 0);
  const bool want_digest = digest_out != nullptr;

  auto worker = [&](int t) {
    std::vector<uint32_t> left, right, rle, hits;
    std::vector<int32_t> out[6];
    hits.reserve(256);
    uint64_t my_pairs = 0, my_digest = 0;
    for (uint64_t b = t; b < nbatches; b += threads) {
      const uint64_t lo = b * batch_rows, hi = std::min<uint64_t>(n, lo + batch_rows);
      left.clear(); rle.clear(); right.clear();
      for (uint64_t i = lo; i < hi; ++i) {
        hits.clear();
        idx->probe_row(key[i], start[i], end[i], hits);
        rle.push_back(uint32_t(hits.size()));
        left.insert(left.end(), hits.begin(), hits.end());
      }
      right.reserve(left.size());
      for (uint32_t i = 0; i < rle.size(); ++i)
        for (uint32_t k = 0; k < rle[i]; ++k) right.push_back(i);
      if (materialise) {
        for (int c = 0; c < 3; ++c) {
          out[c].resize(left.size());
          for (size_t k = 0; k < left.size(); ++k) out[c][k] = bcols[c][left[k]];
          out[3 + c].resize(right.size());
          const int32_t* pc = pcols[c] + lo;
          for (size_t k = 0; k < right.size(); ++k) out[3 + c][k] = pc[right[k]];
        }
        if (!left.empty()) sink[t] += uint64_t(uint32_t(out[0][0])) + uint64_t(uint32_t(out[5].back()));
      }
      if (want_digest)  // checker only; off when timing the baseline
        for (size_t k = 0; k < left.size(); ++k)
          my_digest += mix64((uint64_t(left[k]) << 32) | uint64_t(uint32_t(right[k] + lo)));
      my_pairs += left.size();
    }
    pairs[t] = my_pairs;
    digest[t] = my_digest;
  };

  auto t0 = Clock::now();
  if (threads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker, t);
    for (auto& th : pool) th.join();
  }
  const double sec = std::chrono::duration<double>(Clock::now() - t0).count();
  uint64_t tp = 0, td = 0;
  for (int t = 0; t < threads; ++t) { tp += pairs[t]; td += digest[t]; }
  if (pairs_out) *pairs_out = tp;
  if (digest_out) *digest_out = td;
  g_sink = sink[0];
  return sec;
}

}  // extern "C"
