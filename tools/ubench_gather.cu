// Micro-benchmark (not part of the product): how fast can a B200 read random granules of G bytes
// out of a buffer much larger than L2?  Guides the layout of the build index (bytes per probe row
// are dominated by the distinct DRAM granules a probe touches).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_gather tools/ubench_gather.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31; return x;
}

// each thread reads G bytes (as 16-byte vectors) at a random G-aligned (or unaligned by `skew` bytes) offset
template <int G>
__global__ void __launch_bounds__(256) k_gather(const uint4* __restrict__ buf, uint64_t n_granules, uint64_t n_items,
                                                uint32_t* __restrict__ out, uint32_t seed, int skew16) {
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  const uint64_t g = mix64(i * 0x9E3779B97F4A7C15ull + seed) % (n_granules - 1);
  const uint4* p = buf + g * (G / 16) + skew16;
  uint32_t acc = 0;
  if (G >= 16) {
    uint4 v[G / 16 ? G / 16 : 1];
#pragma unroll
    for (int k = 0; k < G / 16; ++k) v[k] = __ldg(p + k);
#pragma unroll
    for (int k = 0; k < G / 16; ++k) acc += v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
  } else {
    acc = __ldg(reinterpret_cast<const uint32_t*>(buf) + g);
  }
  out[i] = acc;
}

// dependent pair: a 4-byte directory read, then a G-byte granule whose address depends on it
template <int G>
__global__ void __launch_bounds__(256) k_chain(const uint32_t* __restrict__ dir, uint64_t n_dir, const uint4* __restrict__ buf,
                                               uint64_t n_granules, uint64_t n_items, uint32_t* __restrict__ out,
                                               uint32_t seed) {
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  const uint64_t d = mix64(i * 0x9E3779B97F4A7C15ull + seed) % n_dir;
  const uint64_t g = (uint64_t(__ldg(dir + d)) * 2654435761ull) % (n_granules - 1);
  const uint4* p = buf + g * (G / 16);
  uint4 v[G / 16];
#pragma unroll
  for (int k = 0; k < G / 16; ++k) v[k] = __ldg(p + k);
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < G / 16; ++k) acc += v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
  out[i] = acc;
}


// cooperative: a group of L = G/16 lanes reads one random granule, one LDG.128 per lane (one request
// carries all the sectors of the granule that share a 128-byte line)
template <int G>
__global__ void __launch_bounds__(256) k_gather_coop(const uint4* __restrict__ buf, uint64_t n_granules, uint64_t n_items,
                                                     uint32_t* __restrict__ out, uint32_t seed) {
  constexpr int L = G / 16;
  const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int sub = threadIdx.x % L;
  constexpr int kPer = 4;  // granules per group, independent loads in flight
  const uint64_t grp = t / L;
  uint32_t acc = 0;
  uint4 v[kPer];
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const uint64_t i = grp * kPer + k;
    const uint64_t g = mix64(i * 0x9E3779B97F4A7C15ull + seed) % (n_granules - 1);
    v[k] = (i < n_items) ? __ldg(buf + g * L + sub) : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int k = 0; k < kPer; ++k) acc += v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
  if (grp * kPer < n_items) out[t % n_items] = acc;
}

template <int G>
void run_coop(const uint4* buf, uint64_t bytes, uint64_t n_items, uint32_t* out, const char* tag) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const uint64_t n_gran = bytes / G;
  const uint64_t threads = (n_items + 3) / 4 * (G / 16);
  const int grid = int((threads + 255) / 256);
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    CK(cudaEventRecord(a));
    k_gather_coop<G><<<grid, 256>>>(buf, n_gran, n_items, out, 17u * it + 1u);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (it && ms < best) best = ms;
  }
  printf("%s coop G=%3d items=%llu  %.3f ms  %.2f Gitems/s  useful %.0f GB/s\n", tag, G,
         (unsigned long long)n_items, best, n_items / best / 1e6, double(n_items) * G / best / 1e6);
}

// per-thread, kPer independent granules per thread (more loads in flight per thread, fewer threads)
template <int G, int kPer>
__global__ void __launch_bounds__(256) k_gather_multi(const uint4* __restrict__ buf, uint64_t n_granules, uint64_t n_items,
                                                      uint32_t* __restrict__ out, uint32_t seed) {
  const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  uint4 v[kPer][G / 16];
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const uint64_t i = t * kPer + k;
    const uint64_t g = mix64(i * 0x9E3779B97F4A7C15ull + seed) % (n_granules - 1);
#pragma unroll
    for (int q = 0; q < G / 16; ++q) v[k][q] = __ldg(buf + g * (G / 16) + q);
  }
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < kPer; ++k)
#pragma unroll
    for (int q = 0; q < G / 16; ++q) acc += v[k][q].x ^ v[k][q].y ^ v[k][q].z ^ v[k][q].w;
  if (t * kPer < n_items) out[t] = acc;
}

template <int G, int kPer>
void run_multi(const uint4* buf, uint64_t bytes, uint64_t n_items, uint32_t* out, const char* tag) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const uint64_t n_gran = bytes / G;
  const uint64_t threads = n_items / kPer;
  const int grid = int((threads + 255) / 256);
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    CK(cudaEventRecord(a));
    k_gather_multi<G, kPer><<<grid, 256>>>(buf, n_gran, n_items, out, 17u * it + 1u);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (it && ms < best) best = ms;
  }
  printf("%s multi%d G=%3d items=%llu  %.3f ms  %.2f Gitems/s  useful %.0f GB/s\n", tag, kPer, G,
         (unsigned long long)n_items, best, n_items / best / 1e6, double(n_items) * G / best / 1e6);
}

__global__ void k_fill(uint32_t* p, uint64_t n) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = uint32_t(mix64(i));
}

template <int G>
void run(const uint4* buf, uint64_t bytes, uint64_t n_items, uint32_t* out, int skew16, const char* tag) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const uint64_t n_gran = bytes / (G < 16 ? 4 : G);
  const int grid = int((n_items + 255) / 256);
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    CK(cudaEventRecord(a));
    k_gather<G><<<grid, 256>>>(buf, n_gran, n_items, out, 17u * it + 1u, skew16);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (it && ms < best) best = ms;
  }
  printf("%s G=%3d skew=%d items=%llu  %.3f ms  %.2f Gitems/s  useful %.0f GB/s\n", tag, G, skew16 * 16,
         (unsigned long long)n_items, best, n_items / best / 1e6, double(n_items) * G / best / 1e6);
}

template <int G>
void run_chain(const uint32_t* dir, uint64_t n_dir, const uint4* buf, uint64_t bytes, uint64_t n_items, uint32_t* out) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const uint64_t n_gran = bytes / G;
  const int grid = int((n_items + 255) / 256);
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    CK(cudaEventRecord(a));
    k_chain<G><<<grid, 256>>>(dir, n_dir, buf, n_gran, n_items, out, 17u * it + 1u);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (it && ms < best) best = ms;
  }
  printf("chain dir(%llu MB)->G=%3d items=%llu  %.3f ms  %.2f Gitems/s\n", (unsigned long long)(n_dir * 4 >> 20), G,
         (unsigned long long)n_items, best, n_items / best / 1e6);
}

int main(int argc, char** argv) {
  const uint64_t n_items = 12500000ull * 2;
  uint4* buf; uint32_t* out; uint32_t* dir;
  const uint64_t n_dir = 12500000;
  const uint64_t max_bytes = 6400ull << 20;
  CK(cudaMalloc(&buf, max_bytes + 4096));
  CK(cudaMalloc(&out, n_items * 4));
  CK(cudaMalloc(&dir, n_dir * 4));
  k_fill<<<1184, 256>>>(reinterpret_cast<uint32_t*>(buf), max_bytes / 4);
  k_fill<<<1184, 256>>>(dir, n_dir);
  CK(cudaDeviceSynchronize());
  for (uint64_t mb : {256ull, 1600ull, 6400ull}) {
    const uint64_t bytes = mb << 20;
    printf("--- buffer %llu MB\n", (unsigned long long)mb);
    run<32>(buf, bytes, n_items, out, 0, "gather");
    run<64>(buf, bytes, n_items, out, 0, "gather");
    run<128>(buf, bytes, n_items, out, 0, "gather");
    run_coop<32>(buf, bytes, n_items, out, "gather");
    run_coop<64>(buf, bytes, n_items, out, "gather");
    run_coop<128>(buf, bytes, n_items, out, "gather");
    run_coop<256>(buf, bytes, n_items, out, "gather");
    run_multi<32, 4>(buf, bytes, n_items, out, "gather");
    run_multi<64, 4>(buf, bytes, n_items, out, "gather");
    run_multi<64, 2>(buf, bytes, n_items, out, "gather");
  }
  run_chain<64>(dir, n_dir, buf, 1600ull << 20, n_items, out);
  return 0;
}
