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
