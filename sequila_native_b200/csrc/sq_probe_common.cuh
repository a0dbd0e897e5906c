// Warp-level helpers shared by the probe kernels (sq_probe.cu: SoA walk, sq_probe_packed.cu: packed
// lines): inclusive warp sums, the rank-compacted "flattened" walk over per-lane lists, and the
// status words of the chained scan (decoupled look-back).
#pragma once
#include "sq_internal.cuh"

namespace sq {

// look-back word: [63:62] status, [61:0] value
constexpr uint64_t kFlagAgg = 1ull << 62;  // CTA aggregate available
constexpr uint64_t kFlagInc = 2ull << 62;  // inclusive prefix available
constexpr uint64_t kValMask = (1ull << 62) - 1;

__device__ __forceinline__ uint32_t warp_incl_sum(uint32_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
    if (int(threadIdx.x & 31) >= d) v += o;
  }
  return v;
}

__device__ __forceinline__ uint32_t low_bits(uint32_t nbits) {  // nbits in [0, 32]
  return nbits >= 32 ? 0xffffffffu : ((1u << nbits) - 1u);
}

// ---------------------------------------------------------------------------------------------
// Candidate walks.  A warp owns 32 probe rows, each with a contiguous candidate range [lo, lo+nc).
//   small rows (nc <= 32) are walked FLATTENED: their ranges are concatenated and lane t takes
//     candidate t, so no lane idles on a short list.  The owner of candidate t is found without a
//     search: the non-empty rows are compacted to ranks 0..R-1 once per warp; per 32-candidate
//     chunk one warp-wide OR marks where a new rank starts inside the chunk and one ballot counts
//     the ranks already finished, so rank(t) = finished + popc(starts at or below t).
//   big rows (nc > 32) are walked WIDE: the whole warp takes one row, 32 candidates per step.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kSmallMax = 32;  // rows with <= 32 candidates carry a hit bitmask

struct Flat {
  uint32_t r_incl;   // (rank lane) inclusive candidate prefix of the row with my rank; UINT_MAX past R
  uint32_t r_excl;   // (rank lane) exclusive prefix
  int r_src;         // (rank lane) lane that owns the row with my rank
  uint32_t total;    // candidates of all small rows
};

// snc = this lane's candidate count if its row is small, else 0; inv = 32 bytes of this warp's smem
__device__ __forceinline__ Flat flat_setup(uint32_t snc, int lane, volatile uint8_t* inv) {
  Flat f;
  const uint32_t incl = warp_incl_sum(snc);
  f.total = __shfl_sync(0xffffffffu, incl, 31);
  const unsigned nz = __ballot_sync(0xffffffffu, snc != 0);
  const int R = __popc(nz);
  // invert lane -> rank through shared memory: rank r is owned by the r-th non-empty lane
  __syncwarp();
  if (snc != 0) inv[__popc(nz & ((1u << lane) - 1u))] = uint8_t(lane);
  __syncwarp();
  f.r_src = lane < R ? int(inv[lane]) : 31;
  const uint32_t i_s = __shfl_sync(0xffffffffu, incl, f.r_src);
  const uint32_t n_s = __shfl_sync(0xffffffffu, snc, f.r_src);
  f.r_incl = lane < R ? i_s : 0xffffffffu;
  f.r_excl = lane < R ? i_s - n_s : 0xffffffffu;
  return f;
}

// rank owning flattened candidate t0 + lane (valid when t0 + lane < total)
__device__ __forceinline__ int flat_rank(const Flat& f, uint32_t t0, int lane) {
  const uint32_t d = f.r_excl - t0;  // my rank's first candidate, relative to the chunk
  const unsigned starts = __reduce_or_sync(0xffffffffu, (d - 1u) < 31u ? (1u << d) : 0u);  // 1 <= d <= 31
  const int done = __popc(__ballot_sync(0xffffffffu, f.r_incl <= t0));
  return done + __popc(starts & ((2u << lane) - 2u));  // starts at positions 1..lane
}

}  // namespace sq
