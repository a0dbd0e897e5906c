// What does the B200 memory system allow for the ACCESS PATTERN of the random-order probe (cfg5 shard)?
// Per probe row: 16 streamed bytes in, one random 8-byte directory entry (100 MB table), 1 + p random 128-byte lines
// (1.5 GB table; the second line DEPENDS on the first, as the walk's step back does), 4 bytes of count out and, for the
// join, 6.44 pairs x 8 bytes of coalesced output.  No search, no predicate, no scan: a kernel with the same requests
// and (almost) no instructions.  Its time is the floor any probe over this layout sits on; k_probe_packed's distance to
// it is what the instruction stream and the chained scan cost.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_probe_floor ubench_probe_floor.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31; return x;
}

// 8 lanes per (row, line): one cooperative 128-byte request per line, four rows per warp instruction, eight steps per warp
template <bool DIR, bool SECOND, bool WRITE>
__global__ void __launch_bounds__(128, 8)
k_floor(const uint4* __restrict__ probe, const uint2* __restrict__ dir, uint32_t dir_n, const uint4* __restrict__ lines,
        uint32_t n_lines, uint32_t n, uint32_t* __restrict__ cnt_out, uint32_t* __restrict__ left, uint32_t* __restrict__ right,
        uint32_t back_permille) {
  const int lane = threadIdx.x & 31, sub = lane & 7, g = lane >> 3;
  const uint32_t i = blockIdx.x * 128 + threadIdx.x;
  uint32_t my_line = 0;
  uint32_t acc = 0;
  if (i < n) {
    const uint4 q = probe[i];  // 16 bytes per row, streamed
    const uint64_t h = mix64(uint64_t(i) * 0x9E3779B97F4A7C15ull + q.x);
    if (DIR) {
      const uint2 e = __ldg(dir + uint32_t(h % dir_n));  // random 8-byte entry
      my_line = (e.x ^ uint32_t(h >> 32)) % n_lines;     // depends on the entry
    } else {
      my_line = uint32_t(h >> 32) % n_lines;
    }
  }
  uint4 v[8];
#pragma unroll
  for (int st = 0; st < 8; ++st) {  // row 4 * st + g of the warp
    const uint32_t ln = __shfl_sync(0xffffffffu, my_line, 4 * st + g);
    v[st] = __ldg(lines + size_t(ln) * 8 + sub);
  }
#pragma unroll
  for (int st = 0; st < 8; ++st) {
    uint32_t x = v[st].x ^ v[st].y ^ v[st].z ^ v[st].w;
    x ^= __shfl_xor_sync(0xffffffffu, x, 1);
    x ^= __shfl_xor_sync(0xffffffffu, x, 2);
    x ^= __shfl_xor_sync(0xffffffffu, x, 4);
    if (SECOND) {  // a dependent second line for back_permille of the rows
      const uint32_t ln = __shfl_sync(0xffffffffu, my_line, 4 * st + g);
      if (uint32_t(mix64(ln) % 1000u) < back_permille) {  // which rows: a hash of the line; the ADDRESS depends on the data
        const uint4 w = __ldg(lines + size_t((ln ? ln - 1 : 0) + (x >> 31)) * 8 + sub);
        x ^= w.x ^ w.w;
      }
    }
    const uint32_t got = __shfl_sync(0xffffffffu, x, g * 8);
    if (lane == 4 * st + g) acc = got;  // owner lane keeps its row's value (never true for sub != ...: value only)
    acc ^= (lane >> 3 == g && (lane & 7) == st) ? got : 0u;
  }
  if (i < n) cnt_out[i] = acc;
  if (WRITE) {
    // 6.44 pairs per row on average, written as the real kernel does: a warp's pairs are one contiguous run
    const uint32_t per_warp = 206;  // 32 x 6.44
    const size_t base = size_t(blockIdx.x * 4 + (threadIdx.x >> 5)) * per_warp;
    for (uint32_t t = lane; t < per_warp; t += 32) {
      left[base + t] = acc + t;
      right[base + t] = i;
    }
  }
}

template <bool DIR, bool SECOND, bool WRITE>
float run(const uint4* probe, const uint2* dir, uint32_t dir_n, const uint4* lines, uint32_t n_lines, uint32_t n, uint32_t* cnt,
          uint32_t* left, uint32_t* right, void* flush, size_t flush_bytes) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e9f, sum = 0.f;
  const int reps = 8;
  for (int r = 0; r < reps + 2; ++r) {
    CK(cudaMemsetAsync(flush, r, flush_bytes));
    CK(cudaEventRecord(a));
    k_floor<DIR, SECOND, WRITE><<<(n + 127) / 128, 128>>>(probe, dir, dir_n, lines, n_lines, n, cnt, left, right, 450);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (r >= 2) { sum += ms; best = ms < best ? ms : best; }
  }
  (void)best;
  return sum / reps;
}

int main() {
  const uint32_t n = 12500000, n_lines = 11800000, dir_n = 12500000;
  uint4 *probe, *lines; uint2* dir; uint32_t *cnt, *left, *right; void* flush;
  const size_t flush_bytes = 256u << 20;
  CK(cudaMalloc(&probe, size_t(n) * 16)); CK(cudaMalloc(&lines, size_t(n_lines) * 128)); CK(cudaMalloc(&dir, size_t(dir_n) * 8));
  CK(cudaMalloc(&cnt, size_t(n) * 4)); CK(cudaMalloc(&left, size_t(n / 32 + 4) * 206 * 4)); CK(cudaMalloc(&right, size_t(n / 32 + 4) * 206 * 4));
  CK(cudaMalloc(&flush, flush_bytes));
  CK(cudaMemset(probe, 1, size_t(n) * 16)); CK(cudaMemset(lines, 0, size_t(n_lines - 1) * 128)); CK(cudaMemset(dir, 3, size_t(dir_n) * 8));
  printf("{\"rows\": %u, \"lines_table_GB\": %.2f, \"dir_table_MB\": %.0f,\n", n, n_lines * 128.0 / 1e9, dir_n * 8.0 / 1e6);
  printf(" \"one_line_ms\": %.4f,\n", run<false, false, false>(probe, dir, dir_n, lines, n_lines, n, cnt, left, right, flush, flush_bytes));
  printf(" \"dir_then_line_ms\": %.4f,\n", run<true, false, false>(probe, dir, dir_n, lines, n_lines, n, cnt, left, right, flush, flush_bytes));
  printf(" \"count_floor_ms\": %.4f,\n", run<true, true, false>(probe, dir, dir_n, lines, n_lines, n, cnt, left, right, flush, flush_bytes));
  printf(" \"join_floor_ms\": %.4f,\n", run<true, true, true>(probe, dir, dir_n, lines, n_lines, n, cnt, left, right, flush, flush_bytes));
  printf(" \"what\": \"per row: 16 B in, random 8 B directory entry, random 128 B line (8-lane cooperative request), a dependent second line for 45 %% of the rows, 4 B out; join: + 206 pairs x 8 B per warp, coalesced\"}\n");
  return 0;
}
