// Build side of the `Cuda` interval join: replaces update_hashmap + IntervalJoinAlgorithm::new
// (reference interval_join.rs:1023-1051 and :767-793) with a flat, tree-free index in HBM:
//
//   key hash --(device hash table)--> dense key id
//   radix sort of (key id, start) carrying the build row
//   per key segment: start[], end[], runmax[], row[]  (runmax = running max of end inside the
//   segment) and a direct-address bin directory over the starts
//
// A probe (qs, qe) against its key segment [sb, se) then has its hits inside the contiguous
// candidate range [lo, hi):  hi = first j with start[j] >  qe  (all j < hi have start <= qe)
//                            lo = first j with runmax[j] >= qs (the first row whose end >= qs)
// and the hits are exactly the rows of [lo, hi) with end[j] >= qs — the same set coitrees'
// pruned descent visits (nosimd.rs:343-384), for any input including inverted intervals.
// hi is found through the bin directory (one L2-resident load + a search inside one cache
// line instead of a ~23-step binary search over HBM); lo by galloping back from hi over runmax.
//
// Kernels here are HBM-bound streaming passes; the sort is CUB's radix sort (library code, like
// cuBLAS for a GEMM) restricted to the significant bits 32 + ceil(log2(#keys)).
#include <cub/device/device_radix_sort.cuh>

#include <cstdlib>

#include "sq_internal.cuh"
#include "sq_packed_common.cuh"  // chain_lookback: the chained scan with decoupled look-back

namespace sq {

// ---------------------------------------------------------------------------------------------
// Key hash -> dense id.  Open addressing, linear probing, capacity a power of two.  Genomic joins
// have a few dozen keys (contigs), so after the first few warps every lookup is a read hit; only
// the first occurrence of a key does an atomicCAS.  Lanes holding the same key elect one leader.
// ---------------------------------------------------------------------------------------------
struct HtStatus {
  unsigned int distinct;
  unsigned int overflow;
  unsigned int has_sentinel;
  unsigned int pad;
};

__global__ void __launch_bounds__(256) k_ht_insert(const uint64_t* __restrict__ keys, uint64_t n,
                                                   uint64_t* __restrict__ ht_keys, uint32_t mask,
                                                   HtStatus* st) {
  // per-CTA filter of keys this CTA has already seen in the table: a genomic join has a few dozen keys, so
  // after its first iterations a CTA answers every row from shared memory (hash, one load, one compare)
  __shared__ uint64_t s_seen[256];
  s_seen[threadIdx.x] = kEmptyKey;
  __syncthreads();
  // kU rows per thread and trip, their loads issued together (one 8-byte load per thread in flight is latency-bound:
  // 0.54 ms per 100M keys); warp-uniform trip count so that the whole warp reaches the ballots together
  constexpr int kU = 4;
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x * kU;
  const uint64_t first = (uint64_t(blockIdx.x) * blockDim.x + (threadIdx.x & ~31u)) * kU + (threadIdx.x & 31);
  for (uint64_t i0 = first; i0 - (threadIdx.x & 31) < n; i0 += stride) {
    uint64_t ks[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint64_t i = i0 + uint64_t(u) * 32;
      ks[u] = i < n ? keys[i] : kEmptyKey;
      if (i < n && ks[u] == kEmptyKey) st->has_sentinel = 1;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint64_t key = ks[u];
      const uint32_t h = uint32_t(mix64(key));
      const bool miss = key != kEmptyKey && *reinterpret_cast<volatile uint64_t*>(&s_seen[h & 255u]) != key;
      const unsigned mm = __ballot_sync(0xffffffffu, miss);
      if (!miss) continue;
      // one lane per distinct missing key in the warp does the table work
      const unsigned peers = __match_any_sync(mm, key);
      if ((__ffs(peers) - 1) != int(threadIdx.x & 31)) continue;
      uint32_t slot = h & mask;
      for (uint32_t step = 0; step <= mask; ++step) {
        uint64_t cur = *reinterpret_cast<volatile uint64_t*>(ht_keys + slot);
        if (cur == key) break;
        if (cur == kEmptyKey) {
          const uint64_t old = atomicCAS(reinterpret_cast<unsigned long long*>(ht_keys + slot),
                                         (unsigned long long)kEmptyKey, (unsigned long long)key);
          if (old == kEmptyKey) {
            const unsigned d = atomicAdd(&st->distinct, 1u) + 1u;
            if (d > (mask + 1u) / 2u) st->overflow = 1;  // keep load factor <= 1/2
            break;
          }
          if (old == key) break;
        }
        slot = (slot + 1) & mask;
        if (step == mask) st->overflow = 1;
      }
      *reinterpret_cast<volatile uint64_t*>(&s_seen[h & 255u]) = key;  // in the table now (or the table overflowed)
    }
  }
}

__global__ void __launch_bounds__(256) k_ht_assign(const uint64_t* __restrict__ ht_keys,
                                                   uint32_t* __restrict__ ht_ids, uint32_t cap,
                                                   unsigned int* counter) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= cap) return;
  ht_ids[slot] = (ht_keys[slot] != kEmptyKey) ? atomicAdd(counter, 1u) : kNoKey;
}

// The first attempt at the key table (kFirstCap slots: up to 2048 keys) also collects what the narrow sort keys below need:
// the table slot of every build row and the range of start per slot.  The slot a key lands in indexes the CTA's range table in
// shared memory directly — the global table has resolved the collisions — which is read first as a filter: ranges only
// widen, so a stale read can only ask for an update that is not needed, and after a CTA's first rows almost no row passes it
// on unsorted input; sorted input (one key per warp, every row a new maximum) takes the whole-warp reduction.  The key
// whose hash equals the empty marker gets the extra slot kFirstCap.
constexpr uint32_t kFirstCap = 4096;

__global__ void __launch_bounds__(256) k_ht_insert_ranges(const uint64_t* __restrict__ keys, const int32_t* __restrict__ start,
                                                          uint64_t n, uint64_t* __restrict__ ht_keys, HtStatus* st,
                                                          uint32_t* __restrict__ row_slot, int32_t* __restrict__ slot_min,
                                                          int32_t* __restrict__ slot_max) {
  __shared__ int32_t s_min[kFirstCap + 1], s_max[kFirstCap + 1];
  for (uint32_t k = threadIdx.x; k <= kFirstCap; k += blockDim.x) { s_min[k] = INT32_MAX; s_max[k] = INT32_MIN; }
  __syncthreads();
  constexpr uint32_t mask = kFirstCap - 1;
  constexpr int kU = 4;  // rows per thread and trip, their loads issued together; warp-uniform trip count
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x * kU;
  const uint64_t first = (uint64_t(blockIdx.x) * blockDim.x + (threadIdx.x & ~31u)) * kU + (threadIdx.x & 31);
  for (uint64_t i0 = first; i0 - (threadIdx.x & 31) < n; i0 += stride) {
    uint64_t ks[kU];
    int32_t ss[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint64_t i = i0 + uint64_t(u) * 32;
      ks[u] = i < n ? keys[i] : 0;
      ss[u] = i < n ? start[i] : 0;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint64_t i = i0 + uint64_t(u) * 32;
      const bool valid = i < n;
      const unsigned act = __ballot_sync(0xffffffffu, valid);
      if (!valid) continue;
      const uint64_t key = ks[u];
      uint32_t slot = kFirstCap;
      if (key == kEmptyKey) {
        st->has_sentinel = 1;
      } else {  // find the key or put it there
        slot = uint32_t(mix64(key)) & mask;
        for (uint32_t step = 0;; ++step) {
          // a cached read: a slot never changes once it holds a key, so only "empty" can be stale — and the CAS that
          // follows an empty read returns the truth
          uint64_t cur = ht_keys[slot];
          if (cur == kEmptyKey) {
            cur = atomicCAS(reinterpret_cast<unsigned long long*>(ht_keys + slot), (unsigned long long)kEmptyKey,
                            (unsigned long long)key);
            if (cur == kEmptyKey) {
              const unsigned d = atomicAdd(&st->distinct, 1u) + 1u;
              if (d > kFirstCap / 2u) st->overflow = 1;  // keep load factor <= 1/2
              break;
            }
          }
          if (cur == key) break;
          if (step == mask) { st->overflow = 1; break; }
          slot = (slot + 1) & mask;
        }
      }
      row_slot[i] = slot;
      const int32_t sv = ss[u];
      const bool need = sv < s_min[slot] || sv > s_max[slot];
      if (!__any_sync(act, need)) continue;
      const uint32_t slot0 = __shfl_sync(act, slot, __ffs(act) - 1);
      unsigned peers = act;
      if (!__all_sync(act, slot == slot0)) peers = __match_any_sync(act, slot);  // one lane per distinct key of the warp
      const int32_t mn = __reduce_min_sync(peers, sv);
      const int32_t mx = __reduce_max_sync(peers, sv);
      if ((__ffs(peers) - 1) == int(threadIdx.x & 31)) {
        if (mn < s_min[slot]) atomicMin(&s_min[slot], mn);
        if (mx > s_max[slot]) atomicMax(&s_max[slot], mx);
      }
    }
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k <= kFirstCap; k += blockDim.x)
    if (s_min[k] <= s_max[k]) {
      atomicMin(slot_min + k, s_min[k]);
      atomicMax(slot_max + k, s_max[k]);
    }
}

// sort key = (key id << 32) | (start with the sign bit flipped), value = (end << 32) | build row: the end
// travels with the row through the sort (a gather of end[] through the permutation afterwards would be
// 100M random 4-byte reads, 2 ms; 4 more bytes per row and pass in the sort cost a third of that)
__global__ void __launch_bounds__(256) k_make_sort_keys(const uint64_t* __restrict__ keys,
                                                        const int32_t* __restrict__ start,
                                                        const int32_t* __restrict__ end, uint64_t n,
                                                        const uint64_t* __restrict__ ht_keys,
                                                        const uint32_t* __restrict__ ht_ids,
                                                        uint32_t mask, uint32_t sentinel_id,
                                                        uint64_t* __restrict__ sort_key,
                                                        uint64_t* __restrict__ sort_val,
                                                        unsigned int* __restrict__ inverted) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  bool inv = false;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t id = ht_lookup(ht_keys, ht_ids, mask, sentinel_id, keys[i]);
    sort_key[i] = (uint64_t(id) << 32) | uint64_t(uint32_t(start[i]) ^ 0x80000000u);
    sort_val[i] = (uint64_t(uint32_t(end[i])) << 32) | uint64_t(uint32_t(i));
    inv |= end[i] < start[i];
  }
  if (inv) atomicOr(inverted, 1u);  // some build row has end < start: the rank-difference count does not apply
}

// ---------------------------------------------------------------------------------------------
// Narrow sort keys.  With few keys (<= kNarrowMaxKeys) whose start ranges, laid end to end, fit 32 bits (a genome's
// contigs: hg38 = 3.1e9 positions), the sort key is ONE 32-bit word, key32 = base[id] + (start - min_start[id]) with
// base = exclusive sum of the per-key spans: 4 radix passes over 12-byte pairs instead of 5 over 16-byte ones.
//   k_key_ranges      dense id of every build row + per-key min / max of start (shared-memory aggregation per CTA,
//                     one pair of global atomics per key and CTA)
//   (host)            spans -> bases; anything that does not fit falls back to the 64-bit keys above
//   k_make_sort_keys32
//   k_finalize32      id of a sorted row = the segment its key32 falls into (binary search over <= 4096 bases in
//                     shared memory), start = key32 - base + min_start
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kNarrowMaxKeys = 4096;
struct KeyBase {
  uint32_t base;      // first key32 of this key id
  int32_t min_start;  // smallest start of this key id
};

__global__ void __launch_bounds__(256) k_key_ranges(const uint64_t* __restrict__ keys, const int32_t* __restrict__ start,
                                                    uint64_t n, const uint64_t* __restrict__ ht_keys,
                                                    const uint32_t* __restrict__ ht_ids, uint32_t mask, uint32_t sentinel_id,
                                                    uint32_t n_keys, uint32_t* __restrict__ row_id,
                                                    int32_t* __restrict__ key_min, int32_t* __restrict__ key_max) {
  extern __shared__ int32_t s_rng[];  // [n_keys] min, [n_keys] max
  int32_t* s_min = s_rng;
  int32_t* s_max = s_rng + n_keys;
  for (uint32_t k = threadIdx.x; k < n_keys; k += blockDim.x) { s_min[k] = INT32_MAX; s_max[k] = INT32_MIN; }
  __syncthreads();
  // kU rows per thread and trip (loads issued together: one 12-byte row per thread in flight is latency-bound); the trip
  // count is warp-uniform.  The table is read first as a filter — its values only move outwards, so a stale read can
  // only ask for an update that is not needed — and after a CTA's first few thousand rows almost no row passes it on
  // unsorted input; sorted input (one key per warp, every row a new maximum) takes the whole-warp reduction.
  constexpr int kU = 4;
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x * kU;
  const uint64_t first = (uint64_t(blockIdx.x) * blockDim.x + (threadIdx.x & ~31u)) * kU + (threadIdx.x & 31);
  for (uint64_t i0 = first; i0 - (threadIdx.x & 31) < n; i0 += stride) {
    uint64_t key[kU];
    int32_t st[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint64_t i = i0 + uint64_t(u) * 32;
      key[u] = i < n ? keys[i] : 0;
      st[u] = i < n ? start[i] : 0;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint64_t i = i0 + uint64_t(u) * 32;
      const bool valid = i < n;
      const unsigned act = __ballot_sync(0xffffffffu, valid);
      if (!valid) continue;
      const uint32_t id = ht_lookup(ht_keys, ht_ids, mask, sentinel_id, key[u]);
      row_id[i] = id;
      const bool need = st[u] < s_min[id] || st[u] > s_max[id];
      if (!__any_sync(act, need)) continue;
      const uint32_t id0 = __shfl_sync(act, id, __ffs(act) - 1);
      unsigned peers = act;
      if (!__all_sync(act, id == id0)) peers = __match_any_sync(act, id);  // one lane per distinct key of the warp
      const int32_t mn = __reduce_min_sync(peers, st[u]);
      const int32_t mx = __reduce_max_sync(peers, st[u]);
      if ((__ffs(peers) - 1) == int(threadIdx.x & 31)) {
        if (mn < s_min[id]) atomicMin(&s_min[id], mn);
        if (mx > s_max[id]) atomicMax(&s_max[id], mx);
      }
    }
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < n_keys; k += blockDim.x)
    if (s_min[k] <= s_max[k]) {
      atomicMin(key_min + k, s_min[k]);
      atomicMax(key_max + k, s_max[k]);
    }
}

__global__ void __launch_bounds__(256) k_make_sort_keys32(const uint32_t* __restrict__ row_id, const int32_t* __restrict__ start,
                                                          const int32_t* __restrict__ end, uint64_t n,
                                                          const KeyBase* __restrict__ kb, uint32_t* __restrict__ sort_key,
                                                          uint64_t* __restrict__ sort_val, unsigned int* __restrict__ inverted) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  bool inv = false;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const KeyBase b = kb[row_id[i]];
    const int32_t st = start[i], en = end[i];
    sort_key[i] = b.base + (uint32_t(st) - uint32_t(b.min_start));
    sort_val[i] = (uint64_t(uint32_t(en)) << 32) | uint64_t(uint32_t(i));
    inv |= en < st;
  }
  if (inv) atomicOr(inverted, 1u);
}

// block maxima of end: level 1 over 32 sorted rows (one warp per block), levels 2 and 3 over 32 entries of the level below
__global__ void __launch_bounds__(256) k_block_max_up(const int32_t* __restrict__ in, uint64_t n_in, uint64_t n_out,
                                                      int32_t* __restrict__ out) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t b = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; b < n_out; b += stride) {
    int32_t v = INT32_MIN;
    for (uint64_t k = b * 32; k < b * 32 + 32 && k < n_in; ++k) v = max(v, in[k]);
    out[b] = v;
  }
}

// overlap depth of up to 4096 evenly spaced rows: depth(j) = j - first i of j's segment with runmax[i] >= start[j]
__global__ void __launch_bounds__(256) k_depth_sample(const uint32_t* __restrict__ s_id, const int32_t* __restrict__ s_start,
                                                      const int32_t* __restrict__ s_runmax, const SegMeta* __restrict__ meta,
                                                      uint64_t n, uint32_t samples, unsigned long long* __restrict__ sum) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= samples) return;
  const uint64_t j = uint64_t(t) * n / samples;
  const uint32_t sb = meta[s_id[j]].sb;
  const int32_t st = s_start[j];
  uint64_t lo = sb, hi = j;  // first i in [sb, j] with runmax[i] >= st (i = j qualifies: end >= start for well-formed rows)
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (s_runmax[mid] < st) lo = mid + 1; else hi = mid;
  }
  atomicAdd(sum, (unsigned long long)(j - lo));
}

// (id, end) sort keys of the rows in (id, start) order, and back: the ends sorted inside each key segment
__global__ void __launch_bounds__(256) k_end_keys(const uint32_t* __restrict__ s_id, const int32_t* __restrict__ s_end,
                                                  uint64_t n, uint64_t* __restrict__ out) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t j = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < n; j += stride)
    out[j] = (uint64_t(s_id[j]) << 32) | uint64_t(uint32_t(s_end[j]) ^ 0x80000000u);
}
__global__ void __launch_bounds__(256) k_end_values(const uint64_t* __restrict__ sorted, uint64_t n, int32_t* __restrict__ s_send) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t j = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < n; j += stride)
    s_send[j] = int32_t(uint32_t(sorted[j]) ^ 0x80000000u);
}

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

// word of the segmented running max (see k_runmax_*): id in the high half, end (order-preserving) in the low half
__device__ __forceinline__ uint64_t scan_word(uint32_t id, int32_t end) {
  return (uint64_t(id) << 32) | uint64_t(uint32_t(end) ^ 0x80000000u);
}

__device__ __forceinline__ void tile_max_store(uint64_t m, uint64_t* __restrict__ tile_max) {
  __shared__ uint64_t wmax[kScanThreads / 32];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t r = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) r = max(r, wmax[w]);
    tile_max[blockIdx.x] = r;
  }
}

// After the sort, one CTA per tile of kScanTile sorted rows: write start[], row[], end[], id[] (unpacked from the sorted
// key / value words), the segment boundaries seg_off[id] = first sorted position of key id, and the tile's maximum
// scan word (the reduce half of the segmented running max below).  With s_perm the ids the index hands out are the
// sorted positions themselves (row[j] = j) and s_perm[j] keeps the build row: payload stored in that order is read
// with locality by the gathers (the hits of a probe row are neighbours).
__global__ void __launch_bounds__(kScanThreads) k_finalize(const uint64_t* __restrict__ sorted_key,
                                                           const uint64_t* __restrict__ sorted_val, uint64_t n,
                                                           int32_t* __restrict__ s_start, int32_t* __restrict__ s_end,
                                                           uint32_t* __restrict__ s_row, uint32_t* __restrict__ s_perm,
                                                           uint32_t* __restrict__ s_id, uint32_t* __restrict__ seg_off,
                                                           uint32_t n_keys, uint64_t* __restrict__ tile_max) {
  const uint64_t base = uint64_t(blockIdx.x) * kScanTile;
  uint64_t m = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint64_t j = base + uint64_t(k) * kScanThreads + threadIdx.x;
    if (j >= n) continue;
    const uint64_t key = sorted_key[j];
    const uint64_t v = sorted_val[j];
    const uint32_t id = uint32_t(key >> 32);
    const int32_t en = int32_t(uint32_t(v >> 32));
    s_start[j] = int32_t(uint32_t(key) ^ 0x80000000u);
    // position ids (option cuda_build_ids positions): the probe kernels hand out j itself, the build row goes to perm[]
    s_row[j] = s_perm ? uint32_t(j) : uint32_t(v);
    if (s_perm) s_perm[j] = uint32_t(v);
    s_end[j] = en;
    s_id[j] = id;
    if (j == 0 || uint32_t(sorted_key[j - 1] >> 32) != id) seg_off[id] = uint32_t(j);
    if (j == n - 1) seg_off[n_keys] = uint32_t(n);
    m = max(m, scan_word(id, en));
  }
  tile_max_store(m, tile_max);
}

// the same for narrow sort keys: id = the key whose [base, next base) holds key32
__global__ void __launch_bounds__(kScanThreads) k_finalize32(const uint32_t* __restrict__ sorted_key,
                                                             const uint64_t* __restrict__ sorted_val, uint64_t n,
                                                             const KeyBase* __restrict__ kb, uint32_t n_keys,
                                                             int32_t* __restrict__ s_start, int32_t* __restrict__ s_end,
                                                             uint32_t* __restrict__ s_row, uint32_t* __restrict__ s_perm,
                                                             uint32_t* __restrict__ s_id, uint32_t* __restrict__ seg_off,
                                                             uint64_t* __restrict__ tile_max) {
  extern __shared__ uint32_t s_base[];  // [n_keys]
  for (uint32_t k = threadIdx.x; k < n_keys; k += blockDim.x) s_base[k] = kb[k].base;
  __syncthreads();
  const uint64_t base = uint64_t(blockIdx.x) * kScanTile;
  uint64_t m = 0;
#pragma unroll 2
  for (int k = 0; k < kScanItems; ++k) {
    const uint64_t j = base + uint64_t(k) * kScanThreads + threadIdx.x;
    if (j >= n) continue;
    const uint32_t key = sorted_key[j];
    const uint64_t v = sorted_val[j];
    uint32_t lo = 0, len = n_keys;  // last id with base <= key (base[0] = 0)
    while (len > 1) {
      const uint32_t half = len >> 1;
      if (s_base[lo + half] <= key) { lo += half; len -= half; } else len = half;
    }
    const uint32_t id = lo;
    const int32_t en = int32_t(uint32_t(v >> 32));
    s_start[j] = int32_t(key - s_base[id] + uint32_t(kb[id].min_start));
    s_row[j] = s_perm ? uint32_t(j) : uint32_t(v);
    if (s_perm) s_perm[j] = uint32_t(v);
    s_end[j] = en;
    s_id[j] = id;
    if (j == 0 || sorted_key[j - 1] < s_base[id]) seg_off[id] = uint32_t(j);  // every key id has at least one row
    if (j == n - 1) seg_off[n_keys] = uint32_t(n);
    m = max(m, scan_word(id, en));
  }
  tile_max_store(m, tile_max);
}

// One thread per key segment: bin geometry (power-of-two bin width, 8-16 rows per bin when the
// starts are spread evenly; skewed data just gets longer in-bin searches).
__global__ void __launch_bounds__(256) k_seg_meta(const uint32_t* __restrict__ seg_off,
                                                  const int32_t* __restrict__ s_start, uint32_t n_keys,
                                                  SegMeta* __restrict__ meta, uint32_t rows_per_bin) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_keys) return;
  SegMeta m;
  m.sb = seg_off[id];
  m.se = seg_off[id + 1];
  m.min_start = s_start[m.sb];
  const uint32_t range = uint32_t(s_start[m.se - 1]) - uint32_t(m.min_start);
  const uint32_t rows = m.se - m.sb;
  const uint32_t max_bins = rows / rows_per_bin ? rows / rows_per_bin : 1;
  uint32_t shift = 0;
  while (shift < 32 && uint64_t(range >> shift) + 1ull > uint64_t(max_bins)) ++shift;
  m.shift = shift;
  m.nbins = (shift >= 32 ? 0u : (range >> shift)) + 1u;
  m.dir_base = 0;  // filled by the host-side exclusive scan over nbins + 1
  m.line_base = 0;  // filled by the host-side scan too
  m.pad1 = 0;
  meta[id] = m;
}

__device__ __forceinline__ uint32_t start_window(int32_t x) { return (uint32_t(x) ^ 0x80000000u) >> 16; }

// dir[dir_base + b] = first row of the segment whose bin >= b; dir[dir_base + nbins] = se
__global__ void __launch_bounds__(256) k_fill_dir(const uint32_t* __restrict__ s_id,
                                                  const int32_t* __restrict__ s_start, uint64_t n,
                                                  const SegMeta* __restrict__ meta, uint32_t* __restrict__ dir) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t j = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < n; j += stride) {
    const uint32_t id = s_id[j];
    const SegMeta m = meta[id];
    const uint32_t sh = m.shift;
    const uint32_t bin = sh >= 32 ? 0u : ((uint32_t(s_start[j]) - uint32_t(m.min_start)) >> sh);
    uint32_t from = 0;
    if (j > m.sb) {
      const uint32_t prev = sh >= 32 ? 0u : ((uint32_t(s_start[j - 1]) - uint32_t(m.min_start)) >> sh);
      from = prev + 1;
    }
    for (uint32_t b = from; b <= bin; ++b) dir[m.dir_base + b] = uint32_t(j);
    if (j == uint64_t(m.se) - 1) dir[m.dir_base + m.nbins] = m.se;
  }
}

// ---------------------------------------------------------------------------------------------
// Segmented running max of end[] as ONE plain prefix-max over u64 words (id << 32 | end'):
// key ids are non-decreasing along the sorted array, so a later segment's words dominate every
// earlier one and the low half of the prefix max is the running max inside the current segment.
// Reduce-then-scan: per-tile max, one-CTA exclusive scan of tile maxima, per-tile rescan.
// ---------------------------------------------------------------------------------------------
// Directory bins hold 8-16 rows for evenly spread starts: measured on B200 (100M-row index, random
// probes) 8 beats 16 and 32 (1.87 / 1.98 / 2.15 ms per 12.5M probes) and costs n/2 bytes of directory.

__device__ __forceinline__ uint64_t warp_incl_max(uint64_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t o = __shfl_up_sync(0xffffffffu, v, d);
    if (int(threadIdx.x & 31) >= d) v = max(v, o);
  }
  return v;
}

// exclusive prefix max over tile_max[0..n_tiles), in place, one CTA of 1024 threads
__global__ void __launch_bounds__(1024) k_runmax_mid(uint64_t* __restrict__ tile_max, uint32_t n_tiles) {
  __shared__ uint64_t wsum[32];
  __shared__ uint64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_tiles; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint64_t v = i < n_tiles ? tile_max[i] : 0;
    uint64_t inc = warp_incl_max(v);
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint64_t w = wsum[threadIdx.x];
      w = warp_incl_max(w);
      wsum[threadIdx.x] = w;
    }
    __syncthreads();
    const uint64_t carry = carry_s;
    const uint64_t wprev = (threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0;
    const uint64_t prev_in_warp = __shfl_up_sync(0xffffffffu, inc, 1);
    uint64_t excl = max(carry, wprev);
    if (threadIdx.x & 31) excl = max(excl, prev_in_warp);
    if (i < n_tiles) tile_max[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = max(carry, max(wprev, inc));
    __syncthreads();
  }
}

// per-tile rescan, kScanItems consecutive rows per thread (one scan and one barrier per tile; the row-per-thread
// arrangement needed kScanItems of each: 0.77 ms per 100M rows against 1.6 GB of traffic).  Four threads' rows are exactly
// one block of the first level of block maxima, written in the same pass.
__global__ void __launch_bounds__(kScanThreads) k_runmax_final(const uint32_t* __restrict__ s_id,
                                                               const int32_t* __restrict__ s_end, uint64_t n,
                                                               const uint64_t* __restrict__ tile_excl,
                                                               int32_t* __restrict__ runmax, int32_t* __restrict__ bmax1) {
  static_assert(kScanItems == 8, "two 16-byte vectors per thread and array");
  __shared__ uint64_t wtot[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t base = uint64_t(blockIdx.x) * kScanTile + uint64_t(threadIdx.x) * kScanItems;
  int32_t en[kScanItems];
  uint32_t id[kScanItems];
  if (base + kScanItems <= n) {  // base is a multiple of 8 rows: 32-byte aligned in both arrays
    const int4 e0 = *reinterpret_cast<const int4*>(s_end + base), e1 = *reinterpret_cast<const int4*>(s_end + base + 4);
    const uint4 i0 = *reinterpret_cast<const uint4*>(s_id + base), i1 = *reinterpret_cast<const uint4*>(s_id + base + 4);
    en[0] = e0.x; en[1] = e0.y; en[2] = e0.z; en[3] = e0.w; en[4] = e1.x; en[5] = e1.y; en[6] = e1.z; en[7] = e1.w;
    id[0] = i0.x; id[1] = i0.y; id[2] = i0.z; id[3] = i0.w; id[4] = i1.x; id[5] = i1.y; id[6] = i1.z; id[7] = i1.w;
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      en[k] = base + k < n ? s_end[base + k] : INT32_MIN;
      id[k] = base + k < n ? s_id[base + k] : 0u;
    }
  }
  uint64_t w[kScanItems];
  uint64_t run = 0;
  int32_t bm = INT32_MIN;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    run = max(run, base + k < n ? scan_word(id[k], en[k]) : uint64_t(0));
    w[k] = run;  // inclusive inside the thread
    bm = max(bm, en[k]);
  }
  bm = max(bm, __shfl_xor_sync(0xffffffffu, bm, 1));
  bm = max(bm, __shfl_xor_sync(0xffffffffu, bm, 2));
  if ((lane & 3) == 0 && base < n) bmax1[base >> 5] = bm;
  const uint64_t inc = warp_incl_max(run);
  if (lane == 31) wtot[warp] = inc;
  uint64_t excl = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) excl = 0;
  __syncthreads();
  uint64_t pre = tile_excl[blockIdx.x];
  for (int q = 0; q < warp; ++q) pre = max(pre, wtot[q]);
  excl = max(excl, pre);
  int32_t out[kScanItems];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) out[k] = int32_t(uint32_t(max(w[k], excl)) ^ 0x80000000u);
  if (base + kScanItems <= n) {
    *reinterpret_cast<int4*>(runmax + base) = make_int4(out[0], out[1], out[2], out[3]);
    *reinterpret_cast<int4*>(runmax + base + 4) = make_int4(out[4], out[5], out[6], out[7]);
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
      if (base + k < n) runmax[base + k] = out[k];
  }
}


// ---------------------------------------------------------------------------------------------
// Packed lines (see sq_internal.cuh).  A line starts at the first row of a key segment, wherever the
// start crosses a 65536-wide window (so every in-line start offset fits 16 bits whatever the gaps in the
// data) and every 15 rows after the segment start: lines hold 1..15 rows, the lines of a segment are
// contiguous.  k_line_number marks line starts, numbers the lines and inverts
// that, k_seg_lines / k_fill_dir_line translate segment starts and directory entries from rows to lines,
// k_pack_lines writes the lines (8 threads per line, each its 16 bytes).
// status[0] |= 1 when some width does not fit 16 bits (the index then keeps only the SoA arrays);
// status[1] += lines a probe ending at this line's last start would walk back.
// ---------------------------------------------------------------------------------------------
// Numbers the packed lines and fills the bin directory in ONE pass (it was: a directory pass that also wrote line flags, a
// library scan, an inversion pass):
// every thread flags its 8 consecutive rows, the tile's count goes through the chained scan with decoupled look-back
// (tiles in ticket order, as in the probe kernels), and line_incl[j] = lines started up to and including row j,
// line_first[line] = its first row come out of the same registers.  result[0] = number of lines.
__global__ void __launch_bounds__(kScanThreads) k_line_number(const uint32_t* __restrict__ s_id, const int32_t* __restrict__ s_start,
                                                              uint64_t n, const SegMeta* __restrict__ meta,
                                                              unsigned long long* chain_state, unsigned int* ticket,
                                                              uint32_t* __restrict__ line_incl, uint32_t* __restrict__ line_first,
                                                              unsigned long long* result, uint32_t* __restrict__ dir) {
  static_assert(kScanItems == 8, "two 16-byte vectors per thread and array");
  __shared__ uint32_t s_tile;
  __shared__ uint32_t wtot[kScanThreads / 32];
  __shared__ unsigned long long s_excl;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t base = uint64_t(tile) * kScanTile + uint64_t(threadIdx.x) * kScanItems;
  uint32_t id[kScanItems];
  int32_t st[kScanItems];
  if (base + kScanItems <= n) {
    const uint4 i0 = *reinterpret_cast<const uint4*>(s_id + base), i1 = *reinterpret_cast<const uint4*>(s_id + base + 4);
    const int4 a0 = *reinterpret_cast<const int4*>(s_start + base), a1 = *reinterpret_cast<const int4*>(s_start + base + 4);
    id[0] = i0.x; id[1] = i0.y; id[2] = i0.z; id[3] = i0.w; id[4] = i1.x; id[5] = i1.y; id[6] = i1.z; id[7] = i1.w;
    st[0] = a0.x; st[1] = a0.y; st[2] = a0.z; st[3] = a0.w; st[4] = a1.x; st[5] = a1.y; st[6] = a1.z; st[7] = a1.w;
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      id[k] = base + k < n ? s_id[base + k] : 0u;
      st[k] = base + k < n ? s_start[base + k] : 0;
    }
  }
  int32_t prev = base > 0 && base < n ? s_start[base - 1] : 0;
  SegMeta m = meta[base < n ? id[0] : 0u];
  uint32_t m_id = id[0];
  uint32_t flags = 0, cnt = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint64_t j = base + k;
    if (j < n) {
      if (id[k] != m_id) { m_id = id[k]; m = meta[m_id]; }
      const bool f = j == m.sb || (j - m.sb) % kLineRows == 0 || start_window(st[k]) != start_window(prev);
      flags |= uint32_t(f) << k;
      cnt += f;
      // the bin directory of the row searches (the former k_fill_dir pass): dir[dir_base + b] = first row of the segment
      // whose bin >= b; dir[dir_base + nbins] = se
      const uint32_t sh = m.shift;
      const uint32_t bin = sh >= 32 ? 0u : ((uint32_t(st[k]) - uint32_t(m.min_start)) >> sh);
      const uint32_t from = j > m.sb ? (sh >= 32 ? 0u : ((uint32_t(prev) - uint32_t(m.min_start)) >> sh)) + 1u : 0u;
      for (uint32_t b = from; b <= bin; ++b) dir[m.dir_base + b] = uint32_t(j);
      if (j == uint64_t(m.se) - 1) dir[m.dir_base + m.nbins] = m.se;
    }
    prev = st[k];
  }
  // exclusive prefix of cnt over the tile's threads
  uint32_t inc = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) wtot[warp] = inc;
  __syncthreads();
  uint32_t pre = 0, tot = 0;
#pragma unroll
  for (int q = 0; q < kScanThreads / 32; ++q) {
    if (q < warp) pre += wtot[q];
    tot += wtot[q];
  }
  if (warp == 0) {
    const unsigned long long excl = chain_lookback(chain_state, tile, (unsigned long long)tot);
    if (lane == 0) {
      s_excl = excl;
      if (uint64_t(tile + 1) * kScanTile >= n) result[0] = excl + tot;  // the last tile
    }
  }
  __syncthreads();
  uint32_t run = uint32_t(s_excl) + pre + inc - cnt;  // lines started before this thread's first row
  uint32_t out[kScanItems];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint64_t j = base + k;
    if ((flags >> k) & 1u) {
      line_first[run] = uint32_t(j);
      run += 1;
    }
    out[k] = run;
    if (j == n - 1) line_first[run] = uint32_t(n);
  }
  if (base + kScanItems <= n) {
    *reinterpret_cast<uint4*>(line_incl + base) = make_uint4(out[0], out[1], out[2], out[3]);
    *reinterpret_cast<uint4*>(line_incl + base + 4) = make_uint4(out[4], out[5], out[6], out[7]);
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
      if (base + k < n) line_incl[base + k] = out[k];
  }
}

__global__ void __launch_bounds__(256) k_seg_lines(const uint32_t* __restrict__ line_incl, uint32_t n_keys, uint64_t n,
                                                   SegMeta* __restrict__ meta) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id > n_keys) return;
  meta[id].line_base = id < n_keys ? line_incl[meta[id].sb] - 1u : line_incl[n - 1];  // sentinel entry: the line total
}

// dir_line[e] = {L, T}: L = line of the last row below directory entry e (= the last line that can hold a start <=
// any qe of bin e - 1), T = the start of L's first row.  A probe with qe < T starts one line earlier: L begins past
// qe, so it holds no candidate (a bin's 8-16 rows straddle a line boundary about half of the time), and the same
// 8-byte load that names the line says so — no read of the line, no second lookup.
__global__ void __launch_bounds__(256) k_fill_dir_line(const uint32_t* __restrict__ dir, uint64_t n_entries,
                                                       const uint32_t* __restrict__ line_incl,
                                                       const uint32_t* __restrict__ line_first,
                                                       const int32_t* __restrict__ s_start, uint2* __restrict__ dir_line) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t e = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n_entries; e += stride) {
    const uint32_t r = dir[e];
    uint2 v = make_uint2(0u, uint32_t(INT32_MIN));
    if (r) {
      v.x = line_incl[r - 1u] - 1u;
      v.y = uint32_t(s_start[line_first[v.x]]);
    }
    dir_line[e] = v;
  }
}

__global__ void __launch_bounds__(256) k_pack_lines(const int32_t* __restrict__ s_start, const int32_t* __restrict__ s_end,
                                                    const int32_t* __restrict__ s_runmax, const uint32_t* __restrict__ s_row,
                                                    const uint32_t* __restrict__ s_id, const SegMeta* __restrict__ meta, uint64_t n_lines,
                                                    const uint32_t* __restrict__ line_first, const uint32_t* __restrict__ line_incl,
                                                    uint4* __restrict__ lines, unsigned long long* status, uint32_t stat_every) {
  const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint64_t line = t >> 3;
  const uint32_t sub = uint32_t(t & 7);
  if (line >= n_lines) return;
  const uint32_t j0 = line_first[line];
  const uint32_t rows = line_first[line + 1] - j0;  // 1..15
  const int32_t base = s_start[j0];
  bool bad = false;
  auto enc = [&](uint32_t r, uint32_t* lo_word, uint32_t* id_word) {
    if (r >= rows) { *lo_word = 0; *id_word = kEmptyRow; return; }
    const uint32_t j = j0 + r;
    const int64_t ds = int64_t(s_start[j]) - int64_t(base);  // < 65536: one window per line
    const int64_t w = int64_t(s_end[j]) - int64_t(s_start[j]);
    if (ds < 0 || ds > 65535 || w < 0 || w > 65535) bad = true;
    *lo_word = uint32_t(ds & 0xFFFF) | (uint32_t(w & 0xFFFF) << 16);
    *id_word = s_row[j];
  };
  uint4 v;
  if (sub == 0) {
    v.x = uint32_t(base);
    const uint32_t id = s_id[j0];
    const bool seg_first = j0 == 0 || s_id[j0 - 1] != id;  // no earlier row of this key
    v.y = uint32_t(seg_first ? INT32_MIN : s_runmax[j0 - 1]);
    enc(0, &v.z, &v.w);
    // statistic, on every stat_every-th line (a binary search of ~20 dependent loads: on every line it cost half of this
    // kernel): how many earlier lines does a probe starting at this line's last start visit?
    if (line % stat_every == 0) {
      const int32_t qs = s_start[j0 + rows - 1];
      const uint32_t sb = meta[id].sb;
      uint32_t a = sb, len = j0 - sb;  // first row in [sb, j0) with runmax >= qs
      while (len) {
        const uint32_t half = len >> 1;
        if (s_runmax[a + half] < qs) { a += half + 1; len -= half + 1; } else len = half;
      }
      if (a < j0) atomicAdd(status + 1, (unsigned long long)(uint32_t(line) - (line_incl[a] - 1u)));
    }
  } else {
    enc(2 * sub - 1, &v.x, &v.y);
    enc(2 * sub, &v.z, &v.w);
  }
  lines[t] = v;
  if (bad) atomicOr(status, 1ull);
}

// ---------------------------------------------------------------------------------------------
static inline int grid_for(uint64_t n, int threads, int sm_count, int per_sm = 8) {
  const uint64_t want = (n + threads - 1) / threads;
  const uint64_t cap = uint64_t(sm_count) * per_sm;  // grid-stride kernels: a few CTAs per SM
  return int(want < 1 ? 1 : (want > cap ? cap : want));
}

void free_index(sq_index* idx) {
  if (!idx) return;
  cudaSetDevice(idx->ctx->device);
  if (idx->perm_stream) sq_stream_free(idx->perm_stream);
  cudaFree(idx->d_start); cudaFree(idx->d_runmax); cudaFree(idx->d_end); cudaFree(idx->d_row); cudaFree(idx->d_perm);
  cudaFree(idx->d_meta); cudaFree(idx->d_dir); cudaFree(idx->d_ht_keys); cudaFree(idx->d_ht_ids);
  cudaFree(idx->d_lines);
  cudaFree(idx->d_dir_line);
  cudaFree(idx->d_send); cudaFree(idx->d_emeta); cudaFree(idx->d_edir);
  cudaFree(idx->d_bmax);
  for (auto& c : idx->columns) {
    if (c.owned) { cudaFree(c.d_values); cudaFree(c.d_offsets); }
    cudaFree(c.d_validity);
  }
  for (auto& pk : idx->packs) cudaFree(pk.d_rows);
  delete idx;
}

// The index arrays themselves and the temporaries of a build (sort double buffers, scan tiles: several GB for
// 100M rows) come from the stream-ordered allocator (cudaFree in free_index hands pool memory back to the
// pool): a build in a long-running process reuses what earlier queries released — nine cudaMalloc calls of
// hundreds of MB inside the build cost 12-24 ms per 100M rows, more than all its kernels together (10 ms).
// Temporaries: the device's default pool keeps them between builds instead of returning them
// to the driver (cudaMalloc / cudaFree of GB-sized blocks cost tens of milliseconds and synchronise).
struct TmpFree {
  cudaStream_t st;
  cudaMemPool_t pool;
  std::vector<void*> ptrs;
  // ONE block for the temporaries whose sizes follow from the row count (sort double buffers, per-row ids, line numbering:
  // 48 bytes per row), handed out by a bump pointer: a build is then a dozen pool operations instead of forty, and the
  // next build finds the same block again (with forty blocks of mixed sizes coming and going the stream-ordered pool
  // re-maps memory every few builds: 10-60 ms hiccups on a 7 ms build).  What does not fit is allocated on its own.
  char* arena = nullptr;
  size_t arena_cap = 0, arena_used = 0;
  TmpFree(cudaStream_t s, cudaMemPool_t pl) : st(s), pool(pl) {}
  ~TmpFree() { for (void* p : ptrs) cudaFreeAsync(p, st); }
  cudaError_t reserve(size_t bytes) {
    cudaError_t e = cudaMallocFromPoolAsync(reinterpret_cast<void**>(&arena), bytes, pool, st);
    if (e != cudaSuccess) { arena = nullptr; return e; }
    ptrs.push_back(arena);
    arena_cap = bytes;
    return cudaSuccess;
  }
  template <class T> cudaError_t alloc(T** p, size_t bytes) {
    bytes = bytes ? bytes : 16;
    const size_t need = (bytes + 255) & ~size_t(255);
    if (arena && arena_used + need <= arena_cap) {
      *p = reinterpret_cast<T*>(arena + arena_used);
      arena_used += need;
      return cudaSuccess;
    }
    cudaError_t e = cudaMallocFromPoolAsync(reinterpret_cast<void**>(p), bytes, pool, st);
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
};

// The library's OWN stream-ordered pool on the context's device: index arrays and build temporaries are kept between
// builds (release threshold = everything) without touching the device's default pool, which belongs to the embedding
// application (DataFusion, other CUDA libraries in the process).
static cudaMemPool_t ctx_pool(sq_ctx* ctx) {
  std::lock_guard<std::mutex> g(ctx->pool_mu);
  if (ctx->pool) return static_cast<cudaMemPool_t>(ctx->pool);
  cudaMemPoolProps props{};
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = ctx->device;
  cudaMemPool_t pool = nullptr;
  if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  uint64_t keep = ~0ull;
  cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  cudaGetLastError();
  ctx->pool = pool;
  return pool;
}

int build_index_device(sq_ctx* ctx, const uint64_t* d_key, const int32_t* d_start, const int32_t* d_end,
                       uint64_t n, cudaStream_t st, sq_index** out) {
  ErrorSlot& E = ctx->err;
  if (n >= 0xFFFFFFFFull)
    return fail(E, SQ_EINVAL, "build side has %llu rows; the interval join addresses rows as u32 "
                "(reference interval_join.rs:1590)", (unsigned long long)n);
  SQ_CUDA(E, cudaSetDevice(ctx->device));
  sq_index* idx = new sq_index();
  idx->ctx = ctx;
  idx->n_rows = n;
  struct Guard { sq_index* p; ~Guard() { if (p) free_index(p); } } guard{idx};
  cudaMemPool_t pool = ctx_pool(ctx);
  if (!pool) return fail(E, SQ_ECUDA, "cudaMemPoolCreate failed on device %d", ctx->device);
  TmpFree tmp(st, pool);
  if (n) SQ_CUDA(E, tmp.reserve(size_t(n) * 48 + std::min<size_t>(size_t(16) << 20, size_t(n) * 16 + (size_t(256) << 10))));

  cudaEvent_t e0, e1;
  SQ_CUDA(E, cudaEventCreate(&e0));
  SQ_CUDA(E, cudaEventCreate(&e1));
  struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{e0, e1};
  SQ_CUDA(E, cudaEventRecord(e0, st));

  // 1. key hash table (grow until the load factor fits).  The first attempt, kFirstCap slots, also records the table slot
  // of every row and the range of start per slot (narrow sort keys, step 2); bigger tables are filled by k_ht_insert alone.
  HtStatus* d_status = nullptr;
  SQ_CUDA(E, tmp.alloc(&d_status, sizeof(HtStatus) + sizeof(unsigned int) * 4));
  unsigned int* d_counter = reinterpret_cast<unsigned int*>(d_status + 1);
  int32_t* d_srng = nullptr;  // [kFirstCap + 1] min of start per slot, [kFirstCap + 1] max
  const bool want_narrow = n && ctx->opt.build_sort.load(std::memory_order_relaxed) == 0;
  bool slot_ranges = false;
  uint32_t* d_rid = nullptr;  // table slot / key id of every build row (narrow sort keys)
  if (want_narrow) SQ_CUDA(E, tmp.alloc(&d_rid, n * 4));
  uint32_t cap = kFirstCap;
  HtStatus hs{};
  for (;;) {
    SQ_CUDA(E, cudaMallocFromPoolAsync(&idx->d_ht_keys, size_t(cap) * 8, pool, st));
    SQ_CUDA(E, cudaMemsetAsync(idx->d_ht_keys, 0xFF, size_t(cap) * 8, st));
    SQ_CUDA(E, cudaMemsetAsync(d_status, 0, sizeof(HtStatus) + 16, st));
    if (n && cap == kFirstCap && want_narrow) {
      std::vector<int32_t> init(2 * (kFirstCap + 1));
      for (uint32_t k = 0; k <= kFirstCap; ++k) { init[k] = INT32_MAX; init[kFirstCap + 1 + k] = INT32_MIN; }
      SQ_CUDA(E, tmp.alloc(&d_srng, init.size() * 4));
      SQ_CUDA(E, cudaMemcpyAsync(d_srng, init.data(), init.size() * 4, cudaMemcpyHostToDevice, st));
      k_ht_insert_ranges<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>(d_key, d_start, n, idx->d_ht_keys, d_status, d_rid,
                                                                           d_srng, d_srng + kFirstCap + 1);
      SQ_CUDA(E, cudaGetLastError());
      SQ_CUDA(E, cudaMemcpyAsync(&hs, d_status, sizeof(hs), cudaMemcpyDeviceToHost, st));
      SQ_CUDA(E, cudaStreamSynchronize(st));  // also: `init` goes out of scope
      slot_ranges = !hs.overflow;
    } else {
      if (n) {
        k_ht_insert<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>(d_key, n, idx->d_ht_keys, cap - 1, d_status);
        SQ_CUDA(E, cudaGetLastError());
      }
      SQ_CUDA(E, cudaMemcpyAsync(&hs, d_status, sizeof(hs), cudaMemcpyDeviceToHost, st));
      SQ_CUDA(E, cudaStreamSynchronize(st));
    }
    if (!hs.overflow) break;
    cudaFree(idx->d_ht_keys);
    idx->d_ht_keys = nullptr;
    // distinct keys <= n, so a table of >= 2n+2 slots can never overflow
    uint64_t max_cap = 4096;
    while (max_cap < 2 * n + 2) max_cap <<= 1;
    if (cap >= max_cap) return fail(E, SQ_ECUDA, "key hash table overflow at full capacity");
    cap = uint32_t(uint64_t(cap) * 16 > max_cap ? max_cap : uint64_t(cap) * 16);
  }
  idx->ht_cap = cap;
  SQ_CUDA(E, cudaMallocFromPoolAsync(&idx->d_ht_ids, size_t(cap) * 4, pool, st));
  k_ht_assign<<<(cap + 255) / 256, 256, 0, st>>>(idx->d_ht_keys, idx->d_ht_ids, cap, d_counter);
  SQ_CUDA(E, cudaGetLastError());
  uint32_t n_keys = hs.distinct;
  if (hs.has_sentinel) idx->sentinel_id = n_keys++;
  idx->n_keys = n_keys;

  // 2. sorted arrays
  SQ_CUDA(E, cudaMallocFromPoolAsync(&idx->d_start, (n ? n : 1) * 4, pool, st));
  SQ_CUDA(E, cudaMallocFromPoolAsync(&idx->d_runmax, (n ? n : 1) * 4, pool, st));
  SQ_CUDA(E, cudaMallocFromPoolAsync(&idx->d_end, (n ? n : 1) * 4, pool, st));
  SQ_CUDA(E, cudaMallocFromPoolAsync(&idx->d_row, (n ? n : 1) * 4, pool, st));
  SQ_CUDA(E, cudaMallocFromPoolAsync(&idx->d_meta, (size_t(n_keys) + 1) * sizeof(SegMeta), pool, st));
  idx->pos_ids = ctx->opt.build_ids.load(std::memory_order_relaxed) == 1;
  if (idx->pos_ids) SQ_CUDA(E, cudaMallocFromPoolAsync(&idx->d_perm, (n ? n : 1) * 4, pool, st));
  idx->bytes = uint64_t(n ? n : 1) * (idx->pos_ids ? 20 : 16) + (uint64_t(n_keys) + 1) * sizeof(SegMeta) + uint64_t(cap) * 12;

  if (n) {
    uint64_t *d_k0 = nullptr, *d_k1 = nullptr;
    uint64_t *d_v0 = nullptr, *d_v1 = nullptr;
    uint32_t* d_seg_off = nullptr;
    SQ_CUDA(E, tmp.alloc(&d_k0, n * 8));
    SQ_CUDA(E, tmp.alloc(&d_k1, n * 8));
    SQ_CUDA(E, tmp.alloc(&d_v0, n * 8));
    SQ_CUDA(E, tmp.alloc(&d_v1, n * 8));
    SQ_CUDA(E, tmp.alloc(&d_seg_off, (size_t(n_keys) + 1) * 4));
    const int g = grid_for(n, 256, ctx->sm_count);
    unsigned int* d_inverted = d_counter + 1;  // zeroed with the hash-table status above
    const uint32_t n_tiles = uint32_t((n + kScanTile - 1) / kScanTile);
    uint64_t* d_tile = nullptr;
    uint32_t* d_sid = nullptr;  // key id of every sorted row
    SQ_CUDA(E, tmp.alloc(&d_tile, size_t(n_tiles) * 8));
    SQ_CUDA(E, tmp.alloc(&d_sid, n * 4));

    // narrow (32-bit) sort keys when the key ranges laid end to end fit 32 bits (option cuda_build_sort wide: never).
    // d_rid tags every build row with its table slot (first-attempt table: the ranges came with the insert) or its key id
    // (bigger table: k_key_ranges); d_tagkb maps the tag to the key's base, d_kb the key id.
    bool narrow = false;
    KeyBase *d_kb = nullptr, *d_tagkb = nullptr;
    int end_bit = 32;
    if (n_keys <= kNarrowMaxKeys && want_narrow) {
      std::vector<int32_t> h_min(n_keys), h_max(n_keys);  // per key id
      std::vector<uint32_t> h_ids;                        // slot -> key id (slot_ranges)
      if (slot_ranges) {
        std::vector<int32_t> h_srng(2 * (kFirstCap + 1));
        h_ids.resize(kFirstCap);
        SQ_CUDA(E, cudaMemcpyAsync(h_srng.data(), d_srng, h_srng.size() * 4, cudaMemcpyDeviceToHost, st));
        SQ_CUDA(E, cudaMemcpyAsync(h_ids.data(), idx->d_ht_ids, size_t(kFirstCap) * 4, cudaMemcpyDeviceToHost, st));
        SQ_CUDA(E, cudaStreamSynchronize(st));
        for (uint32_t sl = 0; sl < kFirstCap; ++sl)
          if (h_ids[sl] != kNoKey) { h_min[h_ids[sl]] = h_srng[sl]; h_max[h_ids[sl]] = h_srng[kFirstCap + 1 + sl]; }
        if (idx->sentinel_id != kNoKey) { h_min[idx->sentinel_id] = h_srng[kFirstCap]; h_max[idx->sentinel_id] = h_srng[2 * kFirstCap + 1]; }
      } else {
        int32_t* d_rng = nullptr;  // [n_keys] min, [n_keys] max
        SQ_CUDA(E, tmp.alloc(&d_rng, size_t(n_keys) * 8));
        std::vector<int32_t> h_rng(size_t(n_keys) * 2);
        for (uint32_t k = 0; k < n_keys; ++k) { h_rng[k] = INT32_MAX; h_rng[n_keys + k] = INT32_MIN; }
        SQ_CUDA(E, cudaMemcpyAsync(d_rng, h_rng.data(), h_rng.size() * 4, cudaMemcpyHostToDevice, st));
        k_key_ranges<<<grid_for(n, 256, ctx->sm_count), 256, size_t(n_keys) * 8, st>>>(
            d_key, d_start, n, idx->d_ht_keys, idx->d_ht_ids, cap - 1, idx->sentinel_id, n_keys, d_rid, d_rng, d_rng + n_keys);
        SQ_CUDA(E, cudaGetLastError());
        SQ_CUDA(E, cudaMemcpyAsync(h_rng.data(), d_rng, h_rng.size() * 4, cudaMemcpyDeviceToHost, st));
        SQ_CUDA(E, cudaStreamSynchronize(st));
        for (uint32_t k = 0; k < n_keys; ++k) { h_min[k] = h_rng[k]; h_max[k] = h_rng[n_keys + k]; }
      }
      std::vector<KeyBase> h_kb(n_keys);
      uint64_t acc = 0;
      for (uint32_t k = 0; k < n_keys; ++k) {
        h_kb[k].base = uint32_t(acc);
        h_kb[k].min_start = h_min[k];
        acc += uint64_t(int64_t(h_max[k]) - int64_t(h_min[k])) + 1ull;  // every key id has at least one row
        if (acc > (1ull << 32)) break;
      }
      if (acc <= (1ull << 32)) {
        narrow = true;
        end_bit = 1;
        while (end_bit < 32 && (1ull << end_bit) < acc) ++end_bit;
        SQ_CUDA(E, tmp.alloc(&d_kb, size_t(n_keys) * sizeof(KeyBase)));
        SQ_CUDA(E, cudaMemcpyAsync(d_kb, h_kb.data(), size_t(n_keys) * sizeof(KeyBase), cudaMemcpyHostToDevice, st));
        std::vector<KeyBase> h_tag;
        d_tagkb = d_kb;
        if (slot_ranges) {
          h_tag.assign(kFirstCap + 1, KeyBase{0u, 0});
          for (uint32_t sl = 0; sl < kFirstCap; ++sl)
            if (h_ids[sl] != kNoKey) h_tag[sl] = h_kb[h_ids[sl]];
          if (idx->sentinel_id != kNoKey) h_tag[kFirstCap] = h_kb[idx->sentinel_id];
          SQ_CUDA(E, tmp.alloc(&d_tagkb, h_tag.size() * sizeof(KeyBase)));
          SQ_CUDA(E, cudaMemcpyAsync(d_tagkb, h_tag.data(), h_tag.size() * sizeof(KeyBase), cudaMemcpyHostToDevice, st));
        }
        SQ_CUDA(E, cudaStreamSynchronize(st));  // h_kb / h_tag go out of scope below
      }
    }
    if (narrow) {
      // the 32-bit keys and their double buffer share d_k0 (n * 8 bytes)
      uint32_t* d_n0 = reinterpret_cast<uint32_t*>(d_k0);
      uint32_t* d_n1 = d_n0 + n;
      k_make_sort_keys32<<<g, 256, 0, st>>>(d_rid, d_start, d_end, n, d_tagkb, d_n0, d_v0, d_inverted);
      SQ_CUDA(E, cudaGetLastError());
      size_t temp_bytes = 0;
      SQ_CUDA(E, ::sq_cub::cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, d_n0, d_n1, d_v0, d_v1, n, 0, end_bit, st));
      void* d_temp = nullptr;
      SQ_CUDA(E, tmp.alloc(&d_temp, temp_bytes));
      SQ_CUDA(E, ::sq_cub::cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, d_n0, d_n1, d_v0, d_v1, n, 0, end_bit, st));
      k_finalize32<<<n_tiles, kScanThreads, size_t(n_keys) * 4, st>>>(d_n1, d_v1, n, d_kb, n_keys, idx->d_start, idx->d_end,
                                                                       idx->d_row, idx->d_perm, d_sid, d_seg_off, d_tile);
      SQ_CUDA(E, cudaGetLastError());
    } else {
      k_make_sort_keys<<<g, 256, 0, st>>>(d_key, d_start, d_end, n, idx->d_ht_keys, idx->d_ht_ids, cap - 1,
                                          idx->sentinel_id, d_k0, d_v0, d_inverted);
      SQ_CUDA(E, cudaGetLastError());
      int key_bits = 0;
      while ((1ull << key_bits) < uint64_t(n_keys)) ++key_bits;
      end_bit = 32 + key_bits;
      size_t temp_bytes = 0;
      SQ_CUDA(E, ::sq_cub::cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, d_k0, d_k1, d_v0, d_v1, n, 0, end_bit, st));
      void* d_temp = nullptr;
      SQ_CUDA(E, tmp.alloc(&d_temp, temp_bytes));
      SQ_CUDA(E, ::sq_cub::cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, d_k0, d_k1, d_v0, d_v1, n, 0, end_bit, st));
      k_finalize<<<n_tiles, kScanThreads, 0, st>>>(d_k1, d_v1, n, idx->d_start, idx->d_end, idx->d_row, idx->d_perm, d_sid,
                                                   d_seg_off, n_keys, d_tile);
      SQ_CUDA(E, cudaGetLastError());
    }
    idx->narrow_sort = narrow;

    // 3. running max of end inside each key segment (the per-tile maxima came out of the finalize pass), and the block
    // maxima of end (32 / 1024 / 32768 rows): long candidate ranges are walked through them
    const uint64_t n1 = (n + 31) / 32, n2 = (n1 + 31) / 32, n3 = (n2 + 31) / 32;
    SQ_CUDA(E, cudaMallocFromPoolAsync(&idx->d_bmax, (n1 + n2 + n3) * 4, pool, st));
    idx->bmax_n1 = n1;
    idx->bmax_n2 = n2;
    idx->bytes += (n1 + n2 + n3) * 4;
    k_runmax_mid<<<1, 1024, 0, st>>>(d_tile, n_tiles);
    SQ_CUDA(E, cudaGetLastError());
    k_runmax_final<<<n_tiles, kScanThreads, 0, st>>>(d_sid, idx->d_end, n, d_tile, idx->d_runmax, idx->d_bmax);
    SQ_CUDA(E, cudaGetLastError());
    k_block_max_up<<<grid_for(n2, 256, ctx->sm_count), 256, 0, st>>>(idx->d_bmax, n1, n2, idx->d_bmax + n1);
    SQ_CUDA(E, cudaGetLastError());
    k_block_max_up<<<grid_for(n3, 256, ctx->sm_count), 256, 0, st>>>(idx->d_bmax + n1, n2, n3, idx->d_bmax + n1 + n2);
    SQ_CUDA(E, cudaGetLastError());

    // 4. per-segment bin directory (geometry on the device, offsets by a host scan: #keys is small)
    uint32_t rows_per_bin = uint32_t(ctx->opt.rows_per_bin.load(std::memory_order_relaxed));
    if (rows_per_bin == 0) rows_per_bin = n <= (4u << 20) ? 1u : 8u;
    k_seg_meta<<<(n_keys + 255) / 256, 256, 0, st>>>(d_seg_off, idx->d_start, n_keys, idx->d_meta,
                                                     rows_per_bin);
    SQ_CUDA(E, cudaGetLastError());
    std::vector<SegMeta> h_meta(n_keys);
    SQ_CUDA(E, cudaMemcpyAsync(h_meta.data(), idx->d_meta, size_t(n_keys) * sizeof(SegMeta), cudaMemcpyDeviceToHost, st));
    SQ_CUDA(E, cudaStreamSynchronize(st));
    uint64_t dir_total = 0;
    h_meta.resize(size_t(n_keys) + 1);  // + sentinel entry that will carry the line total
    for (uint32_t k = 0; k < n_keys; ++k) {
      SegMeta& m = h_meta[k];
      m.dir_base = uint32_t(dir_total);
      dir_total += uint64_t(m.nbins) + 1;
    }
    if (dir_total >= 0xFFFFFFFFull) return fail(E, SQ_EINVAL, "bin directory too large");
    h_meta[n_keys] = SegMeta{};
    h_meta[n_keys].sb = h_meta[n_keys].se = uint32_t(n);
    SQ_CUDA(E, cudaMemcpyAsync(idx->d_meta, h_meta.data(), (size_t(n_keys) + 1) * sizeof(SegMeta), cudaMemcpyHostToDevice, st));
    SQ_CUDA(E, cudaMallocFromPoolAsync(&idx->d_dir, dir_total * 4, pool, st));
    idx->bytes += dir_total * 4;
    idx->dir_bytes = dir_total * 4;
    // (the directory is filled by the line-numbering pass below: it reads the same rows)

    // 5. packed lines for narrow indexes (every width < 65536)
    uint32_t *d_line_incl = nullptr, *d_line_first = nullptr;
    unsigned long long* d_lchain = nullptr;  // [n_tiles] chained-scan words, ticket, line total
    SQ_CUDA(E, tmp.alloc(&d_line_incl, n * 4));
    SQ_CUDA(E, tmp.alloc(&d_line_first, (n + 1) * 4));  // at most one line per row
    SQ_CUDA(E, tmp.alloc(&d_lchain, (size_t(n_tiles) + 2) * 8));
    SQ_CUDA(E, cudaMemsetAsync(d_lchain, 0, (size_t(n_tiles) + 2) * 8, st));
    k_line_number<<<n_tiles, kScanThreads, 0, st>>>(d_sid, idx->d_start, n, idx->d_meta, d_lchain,
                                                    reinterpret_cast<unsigned int*>(d_lchain + n_tiles), d_line_incl, d_line_first,
                                                    d_lchain + n_tiles + 1, idx->d_dir);
    SQ_CUDA(E, cudaGetLastError());
    unsigned long long h_lines = 0;
    SQ_CUDA(E, cudaMemcpyAsync(&h_lines, d_lchain + n_tiles + 1, 8, cudaMemcpyDeviceToHost, st));
    SQ_CUDA(E, cudaStreamSynchronize(st));  // also: h_meta must outlive its async copy
    const uint64_t line_total = h_lines;
    k_seg_lines<<<(n_keys + 256) / 256, 256, 0, st>>>(d_line_incl, n_keys, n, idx->d_meta);
    SQ_CUDA(E, cudaGetLastError());
    SQ_CUDA(E, cudaMallocFromPoolAsync(&idx->d_dir_line, dir_total * 8, pool, st));
    k_fill_dir_line<<<grid_for(dir_total, 256, ctx->sm_count), 256, 0, st>>>(idx->d_dir, dir_total, d_line_incl, d_line_first,
                                                                             idx->d_start, idx->d_dir_line);
    SQ_CUDA(E, cudaGetLastError());
    unsigned long long* d_pstat = nullptr;
    SQ_CUDA(E, tmp.alloc(&d_pstat, 32));
    SQ_CUDA(E, cudaMemsetAsync(d_pstat, 0, 32, st));
    SQ_CUDA(E, cudaMallocFromPoolAsync(&idx->d_lines, line_total * 128, pool, st));
    const uint32_t stat_every = line_total > (1u << 16) ? 32u : 1u;
    k_pack_lines<<<unsigned((line_total * 8 + 255) / 256), 256, 0, st>>>(idx->d_start, idx->d_end, idx->d_runmax, idx->d_row,
                                                                          d_sid, idx->d_meta, line_total, d_line_first,
                                                                          d_line_incl, idx->d_lines, d_pstat, stat_every);
    SQ_CUDA(E, cudaGetLastError());
    // the sampled overlap depth (step 6 decides on it) and the "some row has end < start" flag come back with the line
    // statistics: one copy, one wait
    const uint32_t samples = uint32_t(n < 4096 ? n : 4096);
    k_depth_sample<<<(samples + 255) / 256, 256, 0, st>>>(d_sid, idx->d_start, idx->d_runmax, idx->d_meta, n, samples, d_pstat + 2);
    SQ_CUDA(E, cudaGetLastError());
    unsigned long long h_pstat[4] = {0, 0, 0, 0};
    unsigned int h_inverted = 0;
    SQ_CUDA(E, cudaMemcpyAsync(h_pstat, d_pstat, 32, cudaMemcpyDeviceToHost, st));
    SQ_CUDA(E, cudaMemcpyAsync(&h_inverted, d_inverted, 4, cudaMemcpyDeviceToHost, st));
    SQ_CUDA(E, cudaStreamSynchronize(st));
    idx->mean_depth = float(double(h_pstat[2]) / double(samples));
    if (h_pstat[0]) {  // wide or inverted intervals: the SoA arrays serve this index
      cudaFree(idx->d_lines);
      cudaFree(idx->d_dir_line);
      idx->d_lines = nullptr;
      idx->d_dir_line = nullptr;
    } else {
      idx->n_lines = line_total;
      idx->mean_back_lines = float(double(h_pstat[1]) / double((line_total + stat_every - 1) / stat_every));
      idx->bytes += line_total * 128 + dir_total * 8;
    }

    // 6. rank structure over the ends, for the indexes the SoA kernels serve (cache-resident, wide or deep ones — the
    // packed-line kernel never reads it): the same rows' ends sorted inside each key segment plus a bin directory, so
    // that a probe row's hit count is |{start <= qe}| - |{end < qs}| without touching a candidate (sq_probe_rank.cu).
    // Needs start <= end on every build row.
    const int rank_opt = ctx->opt.rank_count.load(std::memory_order_relaxed);
    if (!h_inverted && !use_packed(idx) && (rank_opt == 2 || (rank_opt == 1 && idx->mean_depth >= 48.f))) {
      k_end_keys<<<g, 256, 0, st>>>(d_sid, idx->d_end, n, d_k0);
      SQ_CUDA(E, cudaGetLastError());
      int key_bits = 0;
      while ((1ull << key_bits) < uint64_t(n_keys)) ++key_bits;
      end_bit = 32 + key_bits;  // (id, end) words are 64-bit whatever the first sort used
      size_t tb = 0;
      SQ_CUDA(E, ::sq_cub::cub::DeviceRadixSort::SortKeys(nullptr, tb, d_k0, d_v0, n, 0, end_bit, st));
      void* d_t2 = nullptr;
      SQ_CUDA(E, tmp.alloc(&d_t2, tb));
      SQ_CUDA(E, ::sq_cub::cub::DeviceRadixSort::SortKeys(d_t2, tb, d_k0, d_v0, n, 0, end_bit, st));
      SQ_CUDA(E, cudaMallocFromPoolAsync(&idx->d_send, n * 4, pool, st));
      k_end_values<<<g, 256, 0, st>>>(d_v0, n, idx->d_send);
      SQ_CUDA(E, cudaGetLastError());
      SQ_CUDA(E, cudaMallocFromPoolAsync(&idx->d_emeta, (size_t(n_keys) + 1) * sizeof(SegMeta), pool, st));
      k_seg_meta<<<(n_keys + 255) / 256, 256, 0, st>>>(d_seg_off, idx->d_send, n_keys, idx->d_emeta, rows_per_bin);
      SQ_CUDA(E, cudaGetLastError());
      std::vector<SegMeta> h_em(size_t(n_keys) + 1);
      SQ_CUDA(E, cudaMemcpyAsync(h_em.data(), idx->d_emeta, size_t(n_keys) * sizeof(SegMeta), cudaMemcpyDeviceToHost, st));
      SQ_CUDA(E, cudaStreamSynchronize(st));
      uint64_t edir_total = 0;
      for (uint32_t k = 0; k < n_keys; ++k) {
        h_em[k].dir_base = uint32_t(edir_total);
        edir_total += uint64_t(h_em[k].nbins) + 1;
      }
      if (edir_total < 0xFFFFFFFFull) {
        h_em[n_keys] = SegMeta{};
        SQ_CUDA(E, cudaMemcpyAsync(idx->d_emeta, h_em.data(), (size_t(n_keys) + 1) * sizeof(SegMeta), cudaMemcpyHostToDevice, st));
        SQ_CUDA(E, cudaMallocFromPoolAsync(&idx->d_edir, edir_total * 4, pool, st));
        k_fill_dir<<<g, 256, 0, st>>>(d_sid, idx->d_send, n, idx->d_emeta, idx->d_edir);  // same segments, same ids
        SQ_CUDA(E, cudaGetLastError());
        SQ_CUDA(E, cudaStreamSynchronize(st));  // h_em must outlive its async copy
        idx->bytes += n * 4 + edir_total * 4 + (uint64_t(n_keys) + 1) * sizeof(SegMeta);
      } else {
        cudaFree(idx->d_send);
        cudaFree(idx->d_emeta);
        idx->d_send = nullptr;
        idx->d_emeta = nullptr;
      }
    }
  }
  SQ_CUDA(E, cudaEventRecord(e1, st));
  SQ_CUDA(E, cudaStreamSynchronize(st));
  SQ_CUDA(E, cudaEventElapsedTime(&idx->build_ms, e0, e1));
  guard.p = nullptr;
  *out = idx;
  return SQ_OK;
}

}  // namespace sq
