// 64-bit hash of the `on` key columns of one row — the stand-in for
// create_hashes(on, RandomState::with_seeds(0,0,0,0)) (reference interval_join.rs:136, 1037, 1211).
// The reference groups by this u64 only (IJ:1042-1048, 965), so any function that is injective on the
// keys present reproduces its result set; what matters is that EVERY producer of key hashes in this
// library computes the same one: the exec node (host, sq_exec.cpp) and the device-side text scanner
// (sq_scan.cu) both include this header.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SQ_HD __host__ __device__ __forceinline__
#else
#define SQ_HD inline
#endif

namespace sqkey {

constexpr uint64_t kGolden = 0x9e3779b97f4a7c15ull;
constexpr uint64_t kFnvOffset = 0xcbf29ce484222325ull;
constexpr uint64_t kFnvPrime = 0x100000001b3ull;

SQ_HD uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31; return x;
}

// hash of a row before any key column is folded in; also the key of every row of a range-only join,
// whose `on` is the constant pair (1, 1) (sequila_physical_planner.rs:136)
SQ_HD uint64_t seed() { return mix64(1); }

// fold the per-column hash `h` of one more key column into the row hash `acc`
SQ_HD uint64_t fold(uint64_t acc, uint64_t h) { return mix64(acc ^ (mix64(h) + kGolden)); }

// fixed-width key value (1/2/4/8 bytes, zero-extended into `raw`)
SQ_HD uint64_t of_fixed(uint64_t raw) { return raw; }

// Utf8 key value.  Strings of up to 8 bytes (contig names) hash from their bytes packed little-endian
// into one word (`raw8`) and their length; longer ones through FNV-1a over all bytes (`fnv`).  A
// streaming producer keeps both while it reads the bytes and picks by length at the end.
SQ_HD uint64_t of_string(uint64_t raw8, uint64_t fnv, uint64_t len) {
  return len <= 8 ? mix64(raw8 ^ (kGolden * (len + 1))) : fnv;
}

struct StringHasher {
  uint64_t raw8 = 0, fnv = kFnvOffset, len = 0;
  SQ_HD void push(uint8_t c) {
    if (len < 8) raw8 |= uint64_t(c) << (8 * len);
    fnv = (fnv ^ c) * kFnvPrime;
    ++len;
  }
  SQ_HD uint64_t finish() const { return of_string(raw8, fnv, len); }
};

}  // namespace sqkey
