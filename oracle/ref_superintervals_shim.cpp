// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Thin C shim over the reference's OWN C++ interval library, superintervals
// (sequila/sequila-core/superintervals/src/superintervals.hpp), which backs the
// reference's `SuperIntervals` algorithm arm (IJ:858-870, IJ:1012-1017).  The
// reference's tests assert that arm returns the same rows as `Coitrees`
// (IJ:1752-1758, IT:74), so it is a second, reference-authored answer for the
// same predicate.  The header is compiled from where it lies under
// /root/reference (see oracle/Makefile: -I$(REF)/.../superintervals/src); no
// reference source is copied into this repository.  Output: oracle/_ref/libsi_ref.so.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "superintervals.hpp"

extern "C" int64_t siref_join(const uint64_t* bkey, const int32_t* bstart, const int32_t* bend,
                              uint64_t nb, const uint64_t* pkey, const int32_t* pstart,
                              const int32_t* pend, uint64_t np, uint32_t** left_out,
                              uint32_t** right_out) {
  std::unordered_map<uint64_t, si::IntervalMap<int32_t, uint32_t>> maps;
  for (uint64_t i = 0; i < nb; ++i) maps[bkey[i]].add(bstart[i], bend[i], uint32_t(i));
  for (auto& kv : maps) kv.second.build();
  std::vector<uint32_t> l, r, found;
  for (uint64_t p = 0; p < np; ++p) {
    auto it = maps.find(pkey[p]);
    if (it == maps.end()) continue;
    found.clear();
    it->second.search_values(pstart[p], pend[p], found);
    for (uint32_t v : found) {
      l.push_back(v);
      r.push_back(uint32_t(p));
    }
  }
  auto* lo = static_cast<uint32_t*>(std::malloc((l.size() + 1) * 4));
  auto* ro = static_cast<uint32_t*>(std::malloc((r.size() + 1) * 4));
  std::memcpy(lo, l.data(), l.size() * 4);
  std::memcpy(ro, r.data(), r.size() * 4);
  *left_out = lo;
  *right_out = ro;
  return int64_t(l.size());
}

extern "C" void siref_free(void* p) { std::free(p); }

// ---- timing of the probe loop over the reference's library (bench.py cpu_baseline, beside the coitrees
// restatement): one shared read-only set of maps, probe batches of `batch_rows` rows dealt round-robin to
// `threads` threads (= DataFusion CollectLeft with target_partitions = threads), every batch collects its
// (left, right) pairs the way process_probe_batch does (IJ:1586-1618).
#include <atomic>
#include <chrono>
#include <thread>

struct SiRefIndex {
  std::unordered_map<uint64_t, si::IntervalMap<int32_t, uint32_t>> maps;
  double build_seconds = 0;
};

extern "C" void* siref_build(const uint64_t* bkey, const int32_t* bstart, const int32_t* bend, uint64_t nb) {
  auto* ix = new SiRefIndex();
  const auto t0 = std::chrono::steady_clock::now();
  for (uint64_t i = 0; i < nb; ++i) ix->maps[bkey[i]].add(bstart[i], bend[i], uint32_t(i));
  for (auto& kv : ix->maps) kv.second.build();
  ix->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return ix;
}

extern "C" double siref_build_seconds(const void* h) { return static_cast<const SiRefIndex*>(h)->build_seconds; }

extern "C" double siref_time_probe(const void* h, const uint64_t* pkey, const int32_t* pstart, const int32_t* pend,
                                   uint64_t np, int32_t threads, uint32_t batch_rows, uint64_t* pairs_out) {
  const auto* ix = static_cast<const SiRefIndex*>(h);
  if (threads < 1) threads = 1;
  if (batch_rows < 1) batch_rows = 8192;
  const uint64_t n_batches = (np + batch_rows - 1) / batch_rows;
  std::atomic<uint64_t> total{0};
  auto worker = [&](int t) {
    std::vector<uint32_t> l, r, found;
    uint64_t mine = 0;
    for (uint64_t b = uint64_t(t); b < n_batches; b += uint64_t(threads)) {
      l.clear();
      r.clear();
      const uint64_t lo = b * batch_rows, hi = lo + batch_rows < np ? lo + batch_rows : np;
      for (uint64_t p = lo; p < hi; ++p) {
        auto it = ix->maps.find(pkey[p]);
        if (it == ix->maps.end()) continue;
        found.clear();
        it->second.search_values(pstart[p], pend[p], found);
        for (uint32_t v : found) {
          l.push_back(v);
          r.push_back(uint32_t(p - lo));
        }
      }
      mine += l.size();
    }
    total += mine;
  };
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 1; t < threads; ++t) th.emplace_back(worker, t);
  worker(0);
  for (auto& x : th) x.join();
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (pairs_out) *pairs_out = total.load();
  return sec;
}

extern "C" void siref_destroy(void* h) { delete static_cast<SiRefIndex*>(h); }
