// Host-side expansion of the run-length encoded right_idx that travels on the wire (sq_probe_emit_pairs /
// sq_probe_join with host buffers): right_idx[k] = the probe row of pair k, i.e. row i repeated counts[i]
// times — the reference's own loop, interval_join.rs:1611-1618.  Plain C++ (no CUDA): compiled by the host
// compiler so that the AVX2 / AVX-512 variants can use target attributes; the widest one the CPU supports is
// picked once at run time.
//
// Rows with few hits dominate, so every row is written as one 64-byte block of sixteen copies (branch-free for
// counts <= 16) and the cursor advances by its count — later rows overwrite the excess; longer runs loop;
// the last sixteen pairs are filled exactly.  The loop is bound by its stores (a dependent cursor update and
// 64 bytes per row): four 16-byte stores with SSE2, two 32-byte ones with AVX2, one with AVX-512.
#include <cstdint>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define SQ_RLE_X86 1
#endif

namespace sq {

static void expand_tail(const uint32_t* counts, uint32_t n_rows, uint32_t* right, uint64_t n_pairs, uint32_t i, uint64_t o) {
  for (; i < n_rows; ++i)  // never writes past n_pairs
    for (uint32_t k = 0; k < counts[i] && o < n_pairs; ++k) right[o++] = i;
}

#ifdef SQ_RLE_X86
static void expand_sse2(const uint32_t* counts, uint32_t n_rows, uint32_t* right, uint64_t n_pairs) {
  uint64_t o = 0;
  uint32_t i = 0;
  if (n_pairs >= 16) {
    const uint64_t safe = n_pairs - 16;
    for (; i < n_rows && o <= safe; ++i) {
      const uint32_t c = counts[i];
      const __m128i v = _mm_set1_epi32(int(i));
      __m128i* p = reinterpret_cast<__m128i*>(right + o);
      _mm_storeu_si128(p, v);
      _mm_storeu_si128(p + 1, v);
      _mm_storeu_si128(p + 2, v);
      _mm_storeu_si128(p + 3, v);
      if (__builtin_expect(c > 16, 0)) {
        uint64_t q = o + 16;
        const uint64_t end = o + c < n_pairs ? o + c : n_pairs;  // counts that sum past n_pairs never write past it
        for (; q + 4 <= end && q + 4 <= n_pairs; q += 4) _mm_storeu_si128(reinterpret_cast<__m128i*>(right + q), v);
        for (; q < end; ++q) right[q] = i;
      }
      o += c;
      if (o > n_pairs) o = n_pairs;
    }
  }
  expand_tail(counts, n_rows, right, n_pairs, i, o);
}

__attribute__((target("avx2"))) static void expand_avx2(const uint32_t* counts, uint32_t n_rows, uint32_t* right,
                                                        uint64_t n_pairs) {
  uint64_t o = 0;
  uint32_t i = 0;
  if (n_pairs >= 16) {
    const uint64_t safe = n_pairs - 16;
    for (; i < n_rows && o <= safe; ++i) {
      const uint32_t c = counts[i];
      const __m256i v = _mm256_set1_epi32(int(i));
      __m256i* p = reinterpret_cast<__m256i*>(right + o);
      _mm256_storeu_si256(p, v);
      _mm256_storeu_si256(p + 1, v);
      if (__builtin_expect(c > 16, 0)) {
        uint64_t q = o + 16;
        const uint64_t end = o + c < n_pairs ? o + c : n_pairs;  // counts that sum past n_pairs never write past it
        for (; q + 8 <= end && q + 8 <= n_pairs; q += 8) _mm256_storeu_si256(reinterpret_cast<__m256i*>(right + q), v);
        for (; q < end; ++q) right[q] = i;
      }
      o += c;
      if (o > n_pairs) o = n_pairs;
    }
  }
  expand_tail(counts, n_rows, right, n_pairs, i, o);
}

__attribute__((target("avx512f"))) static void expand_avx512(const uint32_t* counts, uint32_t n_rows, uint32_t* right,
                                                            uint64_t n_pairs) {
  uint64_t o = 0;
  uint32_t i = 0;
  if (n_pairs >= 16) {
    const uint64_t safe = n_pairs - 16;
    for (; i < n_rows && o <= safe; ++i) {
      const uint32_t c = counts[i];
      const __m512i v = _mm512_set1_epi32(int(i));
      _mm512_storeu_si512(right + o, v);
      if (__builtin_expect(c > 16, 0)) {
        uint64_t q = o + 16;
        const uint64_t end = o + c < n_pairs ? o + c : n_pairs;  // counts that sum past n_pairs never write past it
        for (; q + 16 <= end && q + 16 <= n_pairs; q += 16) _mm512_storeu_si512(right + q, v);
        for (; q < end; ++q) right[q] = i;
      }
      o += c;
      if (o > n_pairs) o = n_pairs;
    }
  }
  expand_tail(counts, n_rows, right, n_pairs, i, o);
}
#endif

typedef void (*ExpandFn)(const uint32_t*, uint32_t, uint32_t*, uint64_t);

static void expand_scalar(const uint32_t* counts, uint32_t n_rows, uint32_t* right, uint64_t n_pairs) {
  expand_tail(counts, n_rows, right, n_pairs, 0, 0);
}

// variant: -1 = the widest the CPU supports (what the library uses), 0 scalar, 1 SSE2, 2 AVX2, 3 AVX-512
static ExpandFn pick(int variant) {
#ifdef SQ_RLE_X86
  __builtin_cpu_init();
  const bool avx512 = __builtin_cpu_supports("avx512f"), avx2 = __builtin_cpu_supports("avx2");
  if (variant == 3 || (variant < 0 && avx512)) return avx512 ? expand_avx512 : nullptr;
  if (variant == 2 || (variant < 0 && avx2)) return avx2 ? expand_avx2 : nullptr;
  if (variant == 1 || variant < 0) return expand_sse2;
#endif
  return (variant <= 0) ? expand_scalar : nullptr;
}

void expand_counts(const uint32_t* counts, uint32_t n_rows, uint32_t* right, uint64_t n_pairs) {
  static const ExpandFn fn = pick(-1);
  fn(counts, n_rows, right, n_pairs);
}

}  // namespace sq

// test / tuning hook (tests/test_rle_host.py): run one named variant; returns 0 when the CPU lacks it
extern "C" __attribute__((visibility("default"))) int32_t sq_rle_expand_variant(int32_t variant, const uint32_t* counts,
                                                                               uint32_t n_rows, uint32_t* right, uint64_t n_pairs) {
  const sq::ExpandFn fn = sq::pick(variant);
  if (!fn) return 0;
  fn(counts, n_rows, right, n_pairs);
  return 1;
}
