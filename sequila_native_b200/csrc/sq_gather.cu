// Materialise / boundary helper kernels of the `Cuda` interval join.
//
//   k_gather<T>      arrow::compute::take of one fixed-width column (interval_join.rs:1620-1632):
//                    out[k] = values[idx[k]].  Right-side indices are non-decreasing (runs), left-side
//                    indices are random into the build column: HBM/L2-bound gather, 4 elements per
//                    thread in flight.
//   k_cast_i64       evaluate_as_i32 for BIGINT columns (interval_join.rs:1661-1672) with the optional
//                    `- 1` of strict comparisons (intervals.rs:67-69); reports the first row that does
//                    not fit Int32 so the host can reproduce the reference's error text.
//   k_pairs_digest   order-independent multiset digest of emitted pairs (test/bench parity at scale).
#include "sq_internal.cuh"

namespace sq {

template <typename T>
__global__ void __launch_bounds__(256) k_gather(const T* __restrict__ values, const uint32_t* __restrict__ idx,
                                                uint64_t n, T* __restrict__ out) {
  constexpr int kPer = 4;
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x * kPer;
  for (uint64_t base = (uint64_t(blockIdx.x) * blockDim.x) * kPer + threadIdx.x; base < n; base += stride) {
    uint32_t ix[kPer];
    T v[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const uint64_t i = base + uint64_t(k) * blockDim.x;
      ix[k] = i < n ? __ldg(idx + i) : 0u;
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const uint64_t i = base + uint64_t(k) * blockDim.x;
      if (i < n) v[k] = ix[k] == kEmptyRow ? T{} : __ldg(values + ix[k]);  // NULL left side (nearest): zero value
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const uint64_t i = base + uint64_t(k) * blockDim.x;
      if (i < n) out[i] = v[k];
    }
  }
}

static inline int grid_for(uint64_t n, int per_block, int sm_count) {
  const uint64_t want = (n + per_block - 1) / per_block;
  const uint64_t cap = uint64_t(sm_count) * 16;
  return int(want < 1 ? 1 : (want > cap ? cap : want));
}

int launch_gather(sq_stream* s, const void* d_values, const uint32_t* d_idx, uint64_t n, uint32_t width,
                  void* d_out) {
  if (n == 0) return SQ_OK;
  const int g = grid_for(n, 256 * 4, s->ctx->sm_count);
  switch (width) {
    case 4:
      k_gather<uint32_t><<<g, 256, 0, s->stream>>>(static_cast<const uint32_t*>(d_values), d_idx, n,
                                                   static_cast<uint32_t*>(d_out));
      break;
    case 8:
      k_gather<uint64_t><<<g, 256, 0, s->stream>>>(static_cast<const uint64_t*>(d_values), d_idx, n,
                                                   static_cast<uint64_t*>(d_out));
      break;
    case 16:
      k_gather<uint4><<<g, 256, 0, s->stream>>>(static_cast<const uint4*>(d_values), d_idx, n,
                                                static_cast<uint4*>(d_out));
      break;
    default:
      return fail(s->err, SQ_EINVAL, "gather: unsupported value width %u (4, 8 or 16)", width);
  }
  SQ_CUDA(s->err, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

// slot[0] = smallest row index whose value does not fit Int32 (UINT64_MAX when none)
__global__ void __launch_bounds__(256) k_cast_i64(const int64_t* __restrict__ in, uint64_t n, int64_t minus,
                                                  int32_t* __restrict__ out, unsigned long long* slot) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t v = int64_t(uint64_t(in[i]) - uint64_t(minus));
    if (v < int64_t(INT32_MIN) || v > int64_t(INT32_MAX)) atomicMin(slot, (unsigned long long)i);
    out[i] = int32_t(v);
  }
}

// ---------------------------------------------------------------------------------------------
// Several columns per pass (include/sequila_cuda.h, "Several columns per pass").
//   k_pack_columns       rows[i] = {c0[i], c1[i], c2[i], c3[i]} (missing columns = 0): streaming, once per query
//   k_gather_pack        per pair ONE 16-byte read of the build row serves every packed column; 4 pairs per
//                        thread in flight; the column stores are coalesced
//   k_gather_probe_cols  per pair one read of right_idx, then the (sequential, cached) values of every column
// ---------------------------------------------------------------------------------------------
struct Ptr4 {
  const uint32_t* in[4];
  uint32_t* out[4];
};

__global__ void __launch_bounds__(256) k_pack_columns(Ptr4 p, int n_cols, uint64_t n, uint4* __restrict__ rows) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint4 v;
    v.x = __ldg(p.in[0] + i);
    v.y = n_cols > 1 ? __ldg(p.in[1] + i) : 0u;
    v.z = n_cols > 2 ? __ldg(p.in[2] + i) : 0u;
    v.w = n_cols > 3 ? __ldg(p.in[3] + i) : 0u;
    rows[i] = v;
  }
}

template <int N>
__global__ void __launch_bounds__(256) k_gather_pack(const uint4* __restrict__ rows, const uint32_t* __restrict__ idx,
                                                     uint64_t n, Ptr4 p) {
  constexpr int kPer = 4;
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x * kPer;
  for (uint64_t base = (uint64_t(blockIdx.x) * blockDim.x) * kPer + threadIdx.x; base < n; base += stride) {
    uint32_t ix[kPer];
    uint4 v[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const uint64_t i = base + uint64_t(k) * blockDim.x;
      ix[k] = i < n ? __ldg(idx + i) : kEmptyRow;
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k)
      v[k] = ix[k] == kEmptyRow ? make_uint4(0u, 0u, 0u, 0u) : __ldg(rows + ix[k]);  // NULL left side (nearest): zero values
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const uint64_t i = base + uint64_t(k) * blockDim.x;
      if (i < n) {
        p.out[0][i] = v[k].x;
        if (N > 1) p.out[1][i] = v[k].y;
        if (N > 2) p.out[2][i] = v[k].z;
        if (N > 3) p.out[3][i] = v[k].w;
      }
    }
  }
}

template <int N>
__global__ void __launch_bounds__(256) k_gather_probe_cols(const uint32_t* __restrict__ idx, uint64_t n, Ptr4 p) {
  constexpr int kPer = 4;
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x * kPer;
  for (uint64_t base = (uint64_t(blockIdx.x) * blockDim.x) * kPer + threadIdx.x; base < n; base += stride) {
    uint32_t ix[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const uint64_t i = base + uint64_t(k) * blockDim.x;
      ix[k] = i < n ? __ldg(idx + i) : 0u;
    }
#pragma unroll
    for (int c = 0; c < N; ++c) {
      uint32_t v[kPer];
#pragma unroll
      for (int k = 0; k < kPer; ++k) v[k] = __ldg(p.in[c] + ix[k]);
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const uint64_t i = base + uint64_t(k) * blockDim.x;
        if (i < n) p.out[c][i] = v[k];
      }
    }
  }
}

int launch_pack_columns(sq_ctx* ctx, const uint32_t* const* d_cols, int n_cols, uint64_t n_rows, uint4* d_rows) {
  if (n_rows == 0) return SQ_OK;
  Ptr4 p{};
  for (int c = 0; c < n_cols; ++c) p.in[c] = d_cols[c];
  k_pack_columns<<<grid_for(n_rows, 256, ctx->sm_count), 256>>>(p, n_cols, n_rows, d_rows);
  SQ_CUDA(ctx->err, cudaGetLastError());
  SQ_CUDA(ctx->err, cudaDeviceSynchronize());
  return SQ_OK;
}

int launch_gather_pack(sq_stream* s, const uint4* d_rows, const uint32_t* d_idx, uint64_t n, uint32_t* const* d_outs, int n_outs) {
  if (n == 0) return SQ_OK;
  Ptr4 p{};
  for (int c = 0; c < n_outs; ++c) p.out[c] = d_outs[c];
  const int g = grid_for(n, 256 * 4, s->ctx->sm_count);
  switch (n_outs) {
    case 1: k_gather_pack<1><<<g, 256, 0, s->stream>>>(d_rows, d_idx, n, p); break;
    case 2: k_gather_pack<2><<<g, 256, 0, s->stream>>>(d_rows, d_idx, n, p); break;
    case 3: k_gather_pack<3><<<g, 256, 0, s->stream>>>(d_rows, d_idx, n, p); break;
    default: k_gather_pack<4><<<g, 256, 0, s->stream>>>(d_rows, d_idx, n, p); break;
  }
  SQ_CUDA(s->err, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

int launch_gather_probe_columns(sq_stream* s, const uint32_t* const* d_cols, const uint32_t* d_idx, uint64_t n,
                                uint32_t* const* d_outs, int n_cols) {
  if (n == 0) return SQ_OK;
  Ptr4 p{};
  for (int c = 0; c < n_cols; ++c) { p.in[c] = d_cols[c]; p.out[c] = d_outs[c]; }
  const int g = grid_for(n, 256 * 4, s->ctx->sm_count);
  switch (n_cols) {
    case 1: k_gather_probe_cols<1><<<g, 256, 0, s->stream>>>(d_idx, n, p); break;
    case 2: k_gather_probe_cols<2><<<g, 256, 0, s->stream>>>(d_idx, n, p); break;
    case 3: k_gather_probe_cols<3><<<g, 256, 0, s->stream>>>(d_idx, n, p); break;
    default: k_gather_probe_cols<4><<<g, 256, 0, s->stream>>>(d_idx, n, p); break;
  }
  SQ_CUDA(s->err, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

int launch_cast_i64(sq_stream* s, const int64_t* d_in, uint64_t n, int64_t minus, int32_t* d_out,
                    int64_t* bad_value, bool* bad) {
  ErrorSlot& E = s->err;
  *bad = false;
  if (n == 0) return SQ_OK;
  int rc;
  if ((rc = ensure(E, s->d_scalar, 256, false))) return rc;
  if ((rc = ensure(E, s->h_scalar, 256, true))) return rc;
  auto* slot = reinterpret_cast<unsigned long long*>(static_cast<char*>(s->d_scalar.p) + 64);
  SQ_CUDA(E, cudaMemsetAsync(slot, 0xFF, 8, s->stream));
  k_cast_i64<<<grid_for(n, 256, s->ctx->sm_count), 256, 0, s->stream>>>(d_in, n, minus, d_out, slot);
  SQ_CUDA(E, cudaGetLastError());
  s->launches += 1;
  auto* h = static_cast<unsigned long long*>(s->h_scalar.p);
  SQ_CUDA(E, cudaMemcpyAsync(h, slot, 8, cudaMemcpyDeviceToHost, s->stream));
  SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  if (h[0] != ~0ull) {
    int64_t raw = 0;
    SQ_CUDA(E, cudaMemcpy(&raw, d_in + h[0], 8, cudaMemcpyDeviceToHost));
    *bad_value = int64_t(uint64_t(raw) - uint64_t(minus));
    *bad = true;
  }
  return SQ_OK;
}

__global__ void __launch_bounds__(256) k_pairs_digest(const uint32_t* __restrict__ left,
                                                      const uint32_t* __restrict__ right, uint64_t n,
                                                      uint64_t right_offset, unsigned long long* out2) {
  uint64_t sum = 0, x = 0;
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t r = (uint64_t(right[i]) + right_offset) & 0xFFFFFFFFull;
    const uint64_t h = mix64((uint64_t(left[i]) << 32) | r);
    sum += h;
    x ^= h;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, d);
    x ^= __shfl_xor_sync(0xffffffffu, x, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out2, (unsigned long long)sum);
    atomicXor(out2 + 1, (unsigned long long)x);
  }
}

int launch_digest(sq_stream* s, const uint32_t* d_left, const uint32_t* d_right, uint64_t n,
                  uint64_t right_offset, uint64_t out3[3]) {
  ErrorSlot& E = s->err;
  int rc;
  if ((rc = ensure(E, s->d_scalar, 256, false))) return rc;
  if ((rc = ensure(E, s->h_scalar, 256, true))) return rc;
  auto* slot = reinterpret_cast<unsigned long long*>(static_cast<char*>(s->d_scalar.p) + 128);
  SQ_CUDA(E, cudaMemsetAsync(slot, 0, 16, s->stream));
  if (n) {
    k_pairs_digest<<<grid_for(n, 256 * 8, s->ctx->sm_count), 256, 0, s->stream>>>(d_left, d_right, n,
                                                                                  right_offset, slot);
    SQ_CUDA(E, cudaGetLastError());
    s->launches += 1;
  }
  auto* h = static_cast<unsigned long long*>(s->h_scalar.p) + 8;
  SQ_CUDA(E, cudaMemcpyAsync(h, slot, 16, cudaMemcpyDeviceToHost, s->stream));
  SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  out3[0] = n;
  out3[1] = h[0];
  out3[2] = h[1];
  return SQ_OK;
}


// ---------------------------------------------------------------------------------------------
// Utf8 take (arrow::compute::take of a string column, interval_join.rs:1624-1627):
//   k_str_blocksum  per 1024 output rows: sum of the gathered string lengths
//   (launch_scan_u64 over the block sums)
//   k_str_offsets   output offsets = block base + in-block exclusive scan
//   k_str_copy      bytes of every output string
// Contig names are a few bytes long, so one thread copies one string.
// ---------------------------------------------------------------------------------------------
constexpr int kStrBlock = 1024;

__device__ __forceinline__ unsigned long long str_len(const int64_t* __restrict__ off, const uint32_t* __restrict__ idx,
                                                      uint64_t k, uint64_t n) {
  if (k >= n) return 0ull;
  const uint32_t r = __ldg(idx + k);
  if (r == kEmptyRow) return 0ull;  // NULL left side: empty string under a cleared validity bit
  return (unsigned long long)(__ldg(off + r + 1) - __ldg(off + r));
}

__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long* s_w,
                                                              unsigned long long* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = s_w[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += o;
    }
    s_w[lane] = w;
  }
  __syncthreads();
  *total = s_w[31];
  return (warp ? s_w[warp - 1] : 0ull) + inc - v;
}

__global__ void __launch_bounds__(kStrBlock) k_str_blocksum(const int64_t* __restrict__ off,
                                                            const uint32_t* __restrict__ idx, uint64_t n,
                                                            unsigned long long* __restrict__ blk) {
  __shared__ unsigned long long s_w[32];
  unsigned long long tot;
  block_excl_scan(str_len(off, idx, uint64_t(blockIdx.x) * kStrBlock + threadIdx.x, n), s_w, &tot);
  if (threadIdx.x == 0) blk[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kStrBlock) k_str_offsets(const int64_t* __restrict__ off,
                                                           const uint32_t* __restrict__ idx, uint64_t n,
                                                           const unsigned long long* __restrict__ blk_base,
                                                           const unsigned long long* __restrict__ total,
                                                           int64_t* __restrict__ out_off) {
  __shared__ unsigned long long s_w[32];
  const uint64_t k = uint64_t(blockIdx.x) * kStrBlock + threadIdx.x;
  unsigned long long tot;
  const unsigned long long e = block_excl_scan(str_len(off, idx, k, n), s_w, &tot);
  if (k < n) out_off[k] = int64_t(blk_base[blockIdx.x] + e);
  if (k == n) out_off[n] = int64_t(total[0]);
}

__global__ void __launch_bounds__(256) k_str_copy(const int64_t* __restrict__ off, const uint8_t* __restrict__ data,
                                                  const uint32_t* __restrict__ idx, uint64_t n,
                                                  const int64_t* __restrict__ out_off, uint8_t* __restrict__ out) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t k = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += stride) {
    const uint32_t r = __ldg(idx + k);
    if (r == kEmptyRow) continue;
    const int64_t a = __ldg(off + r), b = __ldg(off + r + 1);
    uint8_t* dst = out + out_off[k];
    for (int64_t c = a; c < b; ++c) *dst++ = __ldg(data + c);
  }
}

int launch_str_offsets(sq_stream* s, const int64_t* d_src_off, const uint32_t* d_idx, uint64_t n,
                       int64_t* d_out_off, uint64_t* total) {
  ErrorSlot& E = s->err;
  const uint64_t n_blk64 = (n + 1 + kStrBlock - 1) / kStrBlock;  // +1: the thread k == n writes the last offset
  if (n_blk64 > 0xFFFFFFFFull) return fail(E, SQ_EINVAL, "string gather of %llu rows is too large", (unsigned long long)n);
  const uint32_t n_blk = uint32_t(n_blk64);
  int rc;
  if ((rc = ensure(E, s->d_strblk, size_t(n_blk) * 8 + 16, false))) return rc;
  if ((rc = ensure(E, s->h_scalar, 256, true))) return rc;
  auto* blk = static_cast<unsigned long long*>(s->d_strblk.p);
  auto* tot = blk + n_blk;
  k_str_blocksum<<<n_blk, kStrBlock, 0, s->stream>>>(d_src_off, d_idx, n, blk);
  SQ_CUDA(E, cudaGetLastError());
  s->launches += 1;
  if ((rc = launch_scan_u64(s, blk, n_blk, tot))) return rc;
  k_str_offsets<<<n_blk, kStrBlock, 0, s->stream>>>(d_src_off, d_idx, n, blk, tot, d_out_off);
  SQ_CUDA(E, cudaGetLastError());
  s->launches += 1;
  auto* h = static_cast<unsigned long long*>(s->h_scalar.p) + 16;
  SQ_CUDA(E, cudaMemcpyAsync(h, tot, 8, cudaMemcpyDeviceToHost, s->stream));
  SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  *total = h[0];
  return SQ_OK;
}

int launch_str_copy(sq_stream* s, const int64_t* d_src_off, const uint8_t* d_src_data, const uint32_t* d_idx,
                    uint64_t n, const int64_t* d_out_off, uint8_t* d_out_data) {
  if (n == 0) return SQ_OK;
  k_str_copy<<<grid_for(n, 256, s->ctx->sm_count), 256, 0, s->stream>>>(d_src_off, d_src_data, d_idx, n, d_out_off,
                                                                         d_out_data);
  SQ_CUDA(s->err, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

// Arrow validity take: out bit k = in bit idx[k] (all set when bitmap == nullptr), 0 for a NULL index;
// one thread per output byte; counts the nulls
__global__ void __launch_bounds__(256) k_gather_bits(const uint8_t* __restrict__ bitmap,
                                                     const uint32_t* __restrict__ idx, uint64_t n,
                                                     uint8_t* __restrict__ out, unsigned long long* nulls) {
  const uint64_t n_bytes = (n + 7) / 8;
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  unsigned int my_nulls = 0;
  for (uint64_t b = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; b < n_bytes; b += stride) {
    unsigned int v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint64_t e = b * 8 + k;
      if (e < n) {
        const uint32_t r = __ldg(idx + e);
        const unsigned int bit = r == kEmptyRow ? 0u : (bitmap ? ((__ldg(bitmap + (r >> 3)) >> (r & 7)) & 1u) : 1u);
        v |= bit << k;
        my_nulls += 1u - bit;
      }
    }
    out[b] = uint8_t(v);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) my_nulls += __shfl_xor_sync(0xffffffffu, my_nulls, d);
  if ((threadIdx.x & 31) == 0 && my_nulls) atomicAdd(nulls, (unsigned long long)my_nulls);
}

int launch_gather_bits(sq_stream* s, const uint8_t* d_bitmap, const uint32_t* d_idx, uint64_t n, uint8_t* d_out,
                       uint64_t* null_count) {
  ErrorSlot& E = s->err;
  *null_count = 0;
  if (n == 0) return SQ_OK;
  int rc;
  if ((rc = ensure(E, s->d_scalar, 256, false))) return rc;
  if ((rc = ensure(E, s->h_scalar, 256, true))) return rc;
  auto* slot = reinterpret_cast<unsigned long long*>(static_cast<char*>(s->d_scalar.p) + 192);
  SQ_CUDA(E, cudaMemsetAsync(slot, 0, 8, s->stream));
  k_gather_bits<<<grid_for((n + 7) / 8, 256, s->ctx->sm_count), 256, 0, s->stream>>>(d_bitmap, d_idx, n, d_out, slot);
  SQ_CUDA(E, cudaGetLastError());
  s->launches += 1;
  auto* h = static_cast<unsigned long long*>(s->h_scalar.p) + 24;
  SQ_CUDA(E, cudaMemcpyAsync(h, slot, 8, cudaMemcpyDeviceToHost, s->stream));
  SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  *null_count = h[0];
  return SQ_OK;
}

__global__ void __launch_bounds__(256) k_narrow_offsets(const int64_t* __restrict__ in, uint64_t n, int32_t* __restrict__ out) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = int32_t(in[i]);
}

int launch_narrow_offsets(sq_stream* s, const int64_t* d_in, uint64_t n, int32_t* d_out) {
  if (n == 0) return SQ_OK;
  k_narrow_offsets<<<grid_for(n, 256, s->ctx->sm_count), 256, 0, s->stream>>>(d_in, n, d_out);
  SQ_CUDA(s->err, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

// SQ_TILE_COUNTS_U8: the per-row hit counts of a tile as bytes (a quarter of the bytes over PCIe); *too_big is set when some
// count does not fit, and the caller then moves the 4-byte counts after all
__global__ void __launch_bounds__(256) k_narrow_counts(const uint32_t* __restrict__ cnt, uint32_t n, uint8_t* __restrict__ out,
                                                       unsigned long long* too_big) {
  const uint32_t i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  uint32_t c[4] = {0, 0, 0, 0};
  if (i4 + 4 <= n) {
    const uint4 v = *reinterpret_cast<const uint4*>(cnt + i4);
    c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
    *reinterpret_cast<uint32_t*>(out + i4) = (c[0] & 255u) | ((c[1] & 255u) << 8) | ((c[2] & 255u) << 16) | ((c[3] & 255u) << 24);
  } else {
    for (uint32_t k = 0; i4 + k < n; ++k) { c[k] = cnt[i4 + k]; out[i4 + k] = uint8_t(c[k]); }
  }
  if ((c[0] | c[1] | c[2] | c[3]) > 255u) *too_big = 1ull;
}

int launch_narrow_counts(sq_stream* s, cudaStream_t st, const uint32_t* d_cnt, uint32_t n, uint8_t* d_out, unsigned long long* d_too_big) {
  ErrorSlot& E = s->err;
  SQ_CUDA(E, cudaMemsetAsync(d_too_big, 0, 8, st));
  if (n == 0) return SQ_OK;
  k_narrow_counts<<<(n / 4 + 256) / 256, 256, 0, st>>>(d_cnt, n, d_out, d_too_big);
  SQ_CUDA(E, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

__global__ void __launch_bounds__(256) k_iota(uint32_t* __restrict__ out, uint64_t n) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = uint32_t(i);
}

int launch_iota(sq_stream* s, uint32_t* d_out, uint64_t n) {
  if (n == 0) return SQ_OK;
  k_iota<<<grid_for(n, 256, s->ctx->sm_count), 256, 0, s->stream>>>(d_out, n);
  SQ_CUDA(s->err, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

// ---------------------------------------------------------------------------------------------
// sq_stream_submit_ids: the key column crossed PCIe as 4-byte dictionary ids (12 bytes per probe row instead of 16).
// key = dict[id]; an id outside the dictionary (a NULL key) turns its row into the empty interval
// [INT32_MAX, INT32_MIN], which no build row overlaps (start <= INT32_MIN && end >= INT32_MAX is outside the domain).
__global__ void __launch_bounds__(256) k_expand_ids(const uint32_t* __restrict__ ids, const uint64_t* __restrict__ dict, uint32_t dict_n,
                                                    uint32_t n, uint64_t* __restrict__ key_out, int32_t* __restrict__ start,
                                                    int32_t* __restrict__ end) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t id = ids[i];
  if (id < dict_n) {
    key_out[i] = __ldg(dict + id);
  } else {
    key_out[i] = 0;
    start[i] = INT32_MAX;
    end[i] = INT32_MIN;
  }
}

int launch_expand_ids(sq_stream* s, cudaStream_t st, const uint32_t* d_ids, const uint64_t* d_dict, uint32_t dict_n, uint32_t n,
                      uint64_t* d_key_out, int32_t* d_start, int32_t* d_end) {
  if (n == 0) return SQ_OK;
  k_expand_ids<<<(n + 255) / 256, 256, 0, st>>>(d_ids, d_dict, dict_n, n, d_key_out, d_start, d_end);
  SQ_CUDA(s->err, cudaGetLastError());
  s->launches += 1;
  return SQ_OK;
}

}  // namespace sq
