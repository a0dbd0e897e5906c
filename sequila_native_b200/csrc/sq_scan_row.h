// Row-level pieces of the device text scanner (sq_scan.cu), host + device: where rows start inside a
// 32-byte chunk of text and the parse of one row.  Kept apart from the kernels so that the host-side
// diagnosis of a failing row and the CPU unit test (tests/test_scan_host.py compiles this header with g++)
// run the very code the kernels run.
#pragma once
#include <cstdint>

#include "sq_keyhash.h"

namespace sq {

struct ScanOpts {
  uint8_t delim, comment;
  bool has_header;
  int32_t col_key, col_start, col_end;
  int64_t start_minus, end_minus;
};

enum RowError : int {
  kRowOk = 0,
  kRowFewFields = 1,
  kRowBadStart = 2,
  kRowBadEnd = 3,
  kRowQuoted = 4,
  kRowCastStart = 5,
  kRowCastEnd = 6,
  kRowKeyTooLong = 7,
};

struct RowResult {
  uint64_t key;       // row hash (sq_keyhash.h)
  uint64_t key_off;   // text offset and length of the key field
  uint32_t key_len;
  int32_t start, end;
  int err;
  int64_t bad_value;  // kRowCast*: the value that does not fit Int32
  int fields;
};

// BIGINT field, [+-]?[0-9]+.  Genomic coordinates have at most 9 digits: those accumulate in 32 bits, longer
// numbers continue in 64 bits, and only numbers of 19 digits pay for the Int64 range check.
struct IntField {
  uint32_t lo = 0;
  uint64_t mag = 0;
  int digits = 0;
  bool neg = false, sign = false, bad = false;
  SQ_HD void push(uint8_t c) {
    const uint32_t d = uint32_t(c) - uint32_t('0');
    if (d < 10u) {
      if (digits < 9) {
        lo = lo * 10u + d;
      } else {
        if (digits == 9) mag = lo;
        if (digits >= 18 && mag > (0x7FFFFFFFFFFFFFFFull - d) / 10ull) bad = true;  // beyond Int64: not a BIGINT
        else mag = mag * 10ull + d;
      }
      ++digits;
    } else if ((c == '-' || c == '+') && digits == 0 && !sign) {
      sign = true;
      neg = c == '-';
    } else {
      bad = true;
    }
  }
  SQ_HD bool ok() const { return digits > 0 && !bad; }
  SQ_HD int64_t value() const {
    const int64_t v = digits <= 9 ? int64_t(lo) : int64_t(mag);
    return neg ? -v : v;
  }
};

// Key field: the first 8 bytes packed little-endian (two 32-bit halves) and the length — all the hash of a
// contig name of up to 8 bytes needs (sq_keyhash.h); longer keys are hashed by a second read of the field.
struct KeyField {
  uint32_t lo = 0, hi = 0, len = 0;
  SQ_HD void push(uint8_t c) {
    if (len < 4u) lo |= uint32_t(c) << (8u * len);
    else if (len < 8u) hi |= uint32_t(c) << (8u * (len - 4u));
    ++len;
  }
};

// Byte sources of parse_row: src(p) = the byte at position p, '\n' for every position at or behind the end of
// the text; positions are absolute offsets (TextSrc) or offsets from the start of a tile (TileSrc, 32 bits:
// the per-byte address arithmetic of the kernel stays in one register).
struct TextSrc {  // plain memory (host diagnosis / CPU harness)
  typedef uint64_t pos_t;
  const uint8_t* text;
  uint64_t n;
  SQ_HD uint8_t operator()(uint64_t p) const { return p < n ? text[p] : uint8_t('\n'); }
  SQ_HD uint64_t absolute(uint64_t p) const { return p; }
};
struct TileSrc {  // the kernel: a tile of the text staged in shared memory, global memory behind it
  typedef uint32_t pos_t;
  const uint8_t* tile;  // bytes [base, base + len) of the text
  uint32_t tile_saddr;  // the same as a 32-bit shared-memory address (device code loads through it: a generic
                        // pointer made the compiler rebuild the shared window address for every byte)
  uint64_t base;
  uint32_t len;
  TextSrc rest;
  SQ_HD uint8_t operator()(uint32_t p) const {
    if (p < len) {
#ifdef __CUDA_ARCH__
      uint32_t v;
      asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(tile_saddr + p));
      return uint8_t(v);
#else
      return tile[p];
#endif
    }
    return rest(base + p);
  }
  SQ_HD uint64_t absolute(uint32_t p) const { return base + p; }
};

#ifdef __CUDACC__
struct SmemSrc {  // the kernel, for a row known to end inside the staged tile: no bounds check per byte
  typedef uint32_t pos_t;
  uint32_t tile_saddr;
  uint64_t base;
  __device__ __forceinline__ uint8_t operator()(uint32_t p) const {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(tile_saddr + p));
    return uint8_t(v);
  }
  __device__ __forceinline__ uint64_t absolute(uint32_t p) const { return base + p; }
};
#endif

// One row, from its first byte `q` to its newline (or the end of the text), field by field: the key field is
// hashed while it is read, the two interval fields are parsed as BIGINT, every other field is skipped, and
// nothing behind the last field the table names is read.  Shared by the kernel, by the host-side diagnosis of
// the first failing row and by the CPU harness of the tests.
template <class Src>
SQ_HD RowResult parse_row(const Src& src, typename Src::pos_t q, const ScanOpts& o) {
  typedef typename Src::pos_t pos_t;
  RowResult r;
  r.key = sqkey::seed();
  r.key_off = src.absolute(q);
  r.key_len = 0;
  r.start = r.end = 0;
  r.err = kRowOk;
  r.bad_value = 0;
  int cast_err = kRowOk;  // reported only when the row parses: the reader fails before the cast is evaluated
  bool have_key = o.col_key < 0, have_start = false, have_end = false;
  const int last_needed = o.col_key > o.col_start ? (o.col_key > o.col_end ? o.col_key : o.col_end)
                                                  : (o.col_start > o.col_end ? o.col_start : o.col_end);
  const uint8_t delim = o.delim;
  // a field ends at the delimiter, at '\n' (src gives '\n' behind the text) or at the CR of a CRLF row end;
  // '\t', '\n' and '\r' are all below 14, so ordinary bytes take one or two compares
#define SQ_FIELD_ENDS(c, p) ((c) == delim || ((c) < 14 && ((c) == '\n' || ((c) == '\r' && src(pos_t((p) + 1)) == '\n'))))
  // what stopped a fast loop at p: 1 = the delimiter, 2 = the row's end, 0 = some other byte
#define SQ_END_KIND(c, p) ((c) == delim ? 1 : ((c) == '\n' || ((c) == '\r' && src(pos_t((p) + 1)) == '\n')) ? 2 : 0)
  pos_t p = q;
  int field = 0;
  for (;;) {
    uint8_t c;
    const pos_t f0 = p;
    // Fast paths for what BED / CSV rows are made of — a key of up to 8 bytes, an unsigned number of up to 9
    // digits, a field to skip — as short loops of straight-line iterations; anything else (a sign, a longer
    // number or name, a stray CR, a malformed field) starts the field again in the general code below.
    int kind = 0;
    if (field == o.col_key) {
      uint64_t raw = 0;
      uint32_t len = 0;
      c = src(p);
      while (len < 8u && !((c == delim) | (c == '\n') | (c == '\r'))) {
        raw |= uint64_t(c) << (8u * len);
        ++len;
        c = src(++p);
      }
      kind = SQ_END_KIND(c, p);
      if (kind != 0 && (raw & 0xFFull) != '"') {
        r.key = sqkey::fold(sqkey::seed(), sqkey::of_string(raw, 0, len));
        r.key_off = src.absolute(f0);
        r.key_len = len;
        have_key = true;
      } else {
        kind = 0;
      }
    } else if (field == o.col_start || field == o.col_end) {
      uint32_t lo = 0;
      int digits = 0;
      c = src(p);
      while (digits < 9) {
        const uint32_t d = uint32_t(c) - uint32_t('0');
        if (d > 9u) break;
        lo = lo * 10u + d;
        ++digits;
        c = src(++p);
      }
      kind = digits > 0 ? SQ_END_KIND(c, p) : 0;
      if (kind != 0) {  // 0 .. 999999999: inside Int64; the cast below still checks Int32 after `minus`
        if (field == o.col_start) {
          have_start = true;
          const int64_t v = int64_t(lo) - o.start_minus;
          if (v < int64_t(INT32_MIN) || v > int64_t(INT32_MAX)) { cast_err = kRowCastStart; r.bad_value = v; }
          else r.start = int32_t(v);
        }
        if (field == o.col_end) {
          have_end = true;
          const int64_t v = int64_t(lo) - o.end_minus;
          if (v < int64_t(INT32_MIN) || v > int64_t(INT32_MAX)) { if (cast_err == kRowOk) { cast_err = kRowCastEnd; r.bad_value = v; } }
          else r.end = int32_t(v);
        }
      }
    } else {
      c = src(p);
      // a skipped field that opens with a quote may hide the delimiter inside it (`id,"a,b",chr1,10,20`): DataFusion's
      // CSV reader honours quotes, this scanner does not — reject the row instead of shifting its columns
      if (c == '"' && r.err == kRowOk) r.err = kRowQuoted;
      while (!((c == delim) | (c == '\n') | (c == '\r'))) c = src(++p);
      kind = SQ_END_KIND(c, p);
    }
    if (kind != 0) {
      ++field;
      if (kind == 2) break;            // the row is over
      if (field > last_needed) break;  // everything the table names has been read
      ++p;
      continue;
    }
    p = f0;  // the general code, from the field's first byte
    if (field == o.col_key) {
      KeyField kf;
      for (;; ++p) {
        c = src(p);
        if (SQ_FIELD_ENDS(c, p)) break;
        kf.push(c);
      }
      uint64_t fnv = 0;
      if (kf.len > 8u) {  // a long name: FNV-1a over the whole field
        fnv = sqkey::kFnvOffset;
        for (pos_t k = f0; k != p; ++k) fnv = (fnv ^ src(k)) * sqkey::kFnvPrime;
      }
      r.key = sqkey::fold(sqkey::seed(), sqkey::of_string(uint64_t(kf.lo) | (uint64_t(kf.hi) << 32), fnv, kf.len));
      r.key_off = src.absolute(f0);
      r.key_len = kf.len;
      have_key = true;
      if (kf.len != 0u && (kf.lo & 0xFFu) == '"' && r.err == kRowOk) r.err = kRowQuoted;
      if (kf.len >= 0xFFFFu && r.err == kRowOk) r.err = kRowKeyTooLong;
    } else if (field == o.col_start || field == o.col_end) {
      IntField iv;
      const uint8_t c0 = src(p);
      for (;; ++p) {
        c = src(p);
        if (SQ_FIELD_ENDS(c, p)) break;
        iv.push(c);
      }
      if (c0 == '"' && r.err == kRowOk) r.err = kRowQuoted;
      // a table may name one field twice (start == end column: point intervals), hence no `else`
      if (field == o.col_start) {
        have_start = true;
        if (!iv.ok()) { if (r.err == kRowOk) r.err = kRowBadStart; }
        else {
          const int64_t v = iv.value() - o.start_minus;
          if (v < int64_t(INT32_MIN) || v > int64_t(INT32_MAX)) { cast_err = kRowCastStart; r.bad_value = v; }  // wins over the end column's
          else r.start = int32_t(v);
        }
      }
      if (field == o.col_end) {
        have_end = true;
        if (!iv.ok()) { if (r.err == kRowOk) r.err = kRowBadEnd; }
        else {
          const int64_t v = iv.value() - o.end_minus;
          if (v < int64_t(INT32_MIN) || v > int64_t(INT32_MAX)) { if (cast_err == kRowOk) { cast_err = kRowCastEnd; r.bad_value = v; } }
          else r.end = int32_t(v);
        }
      }
    } else {
      if (src(p) == '"' && r.err == kRowOk) r.err = kRowQuoted;  // quoted skipped field: see the fast path above
      for (;; ++p) {
        c = src(p);
        if (SQ_FIELD_ENDS(c, p)) break;
      }
    }
    ++field;
    if (c != delim) break;           // '\n', or the CR in front of it: the row is over
    if (field > last_needed) break;  // everything the table names has been read
    ++p;
  }
#undef SQ_FIELD_ENDS
#undef SQ_END_KIND
  r.fields = field;
  if (!(have_key && have_start && have_end)) r.err = kRowFewFields;  // reported before any field error: the row is short
  // a cast failure of `start` is reported before one of `end` (evaluation order, interval_join.rs:1039-1040):
  // when the start column follows the end column in the text, the start's failure still wins
  if (r.err == kRowOk && cast_err != kRowOk) r.err = cast_err;
  return r;
}

// bit j = text[q0 + j] == '\n' for the 32 bytes at q0 (a multiple of 32); positions >= n give 0
SQ_HD uint32_t newline_mask32(const uint8_t* text, uint64_t n, uint64_t q0) {
  uint32_t m = 0;
#ifdef __CUDA_ARCH__
  if (q0 + 32 <= n) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(text + q0));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(text + q0 + 16));
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t eq = __vcmpeq4(w[k], 0x0a0a0a0au) & 0x01010101u;  // one flag bit per matching byte
      m |= ((eq * 0x01020408u) >> 24) << (4 * k);                      // the four flags side by side
    }
    return m;
  }
#endif
  for (uint64_t p = q0; p < n && p < q0 + 32; ++p) m |= (text[p] == '\n' ? 1u : 0u) << (p - q0);
  return m;
}

SQ_HD int lowest_bit(uint32_t m) {
#ifdef __CUDA_ARCH__
  return __ffs(m) - 1;
#else
  return __builtin_ctz(m);
#endif
}

// bit j = a row starts at q0 + j, given the newline mask `nl` of the 32 bytes at q0 (`valid` of them lie inside
// the text) and whether the byte in front of q0 is a newline (or q0 == 0): first byte of the text or the byte
// behind a newline, unless the row is empty ("\n", "\r\n", a lone "\r" at the end) or opens with the
// comment byte
template <class Src>
SQ_HD uint32_t row_starts_from_newlines(const Src& src, uint64_t q0, uint32_t nl, bool prev_nl, uint32_t valid,
                                        uint8_t comment) {
  uint32_t cand = ((nl << 1) | (prev_nl ? 1u : 0u)) & ~nl;
  if (valid < 32u) cand &= (1u << valid) - 1u;
  uint32_t m = cand;
  while (m) {
    const int j = lowest_bit(m);
    m &= m - 1;
    const uint64_t p = q0 + j;
    const uint8_t c = src(p);
    const bool skip = (comment != 0 && c == comment) || (c == '\r' && src(p + 1) == '\n');
    if (skip) cand &= ~(1u << j);
  }
  return cand;
}

SQ_HD uint32_t row_start_mask32(const uint8_t* text, uint64_t n, uint64_t q0, uint8_t comment) {
  if (q0 >= n) return 0u;
  const uint32_t nl = newline_mask32(text, n, q0);
  const bool prev_nl = q0 == 0 || text[q0 - 1] == '\n';
  return row_starts_from_newlines(TextSrc{text, n}, q0, nl, prev_nl, n - q0 >= 32 ? 32u : uint32_t(n - q0), comment);
}

}  // namespace sq
