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

struct IntField {
  uint64_t mag = 0;
  int digits = 0;
  bool neg = false, sign = false, bad = false;
  SQ_HD void push(uint8_t c) {
    if ((c == '-' || c == '+') && digits == 0 && !sign) {
      sign = true;
      neg = c == '-';
    } else if (c >= '0' && c <= '9') {
      const uint64_t d = uint64_t(c - '0');
      if (mag > (0x7FFFFFFFFFFFFFFFull - d) / 10ull) bad = true;  // beyond Int64: not a BIGINT
      else mag = mag * 10ull + d;
      ++digits;
    } else {
      bad = true;
    }
  }
  SQ_HD bool ok() const { return digits > 0 && !bad; }
  SQ_HD int64_t value() const { return neg ? -int64_t(mag) : int64_t(mag); }
};

// One row, from its first byte `q` to its newline (or the end of the text).  Shared by the kernel and by
// the host-side diagnosis of the first failing row.
SQ_HD RowResult parse_row(const uint8_t* text, uint64_t n, uint64_t q, const ScanOpts& o) {
  RowResult r;
  r.key = sqkey::seed();
  r.key_off = q;
  r.key_len = 0;
  r.start = r.end = 0;
  r.err = kRowOk;
  r.bad_value = 0;
  int field = 0;
  int cast_err = kRowOk;  // reported only when the row parses: the reader fails before the cast is evaluated
  sqkey::StringHasher kh;
  IntField iv;
  bool have_key = o.col_key < 0, have_start = false, have_end = false;
  uint64_t field_off = q;
  for (uint64_t p = q;; ++p) {
    const uint8_t c = p < n ? text[p] : uint8_t('\n');
    if (c == '\r' && (p + 1 >= n || text[p + 1] == '\n')) continue;  // the CR of a CRLF row end
    const bool eol = c == '\n';
    if (eol || c == o.delim) {
      if (field == o.col_key) {
        r.key = sqkey::fold(sqkey::seed(), kh.finish());
        r.key_off = field_off;
        r.key_len = uint32_t(kh.len);
        have_key = true;
        if (kh.len >= 0xFFFFull && r.err == kRowOk) r.err = kRowKeyTooLong;
      }
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
      ++field;
      kh = sqkey::StringHasher();
      iv = IntField();
      field_off = p + 1;
      if (eol) break;
      continue;
    }
    if (field == o.col_key) {
      if (kh.len == 0 && c == '"' && r.err == kRowOk) r.err = kRowQuoted;
      kh.push(c);
    }
    if (field == o.col_start || field == o.col_end) {
      if (iv.digits == 0 && !iv.sign && c == '"' && r.err == kRowOk) r.err = kRowQuoted;
      iv.push(c);
    }
  }
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

// bit j = a row starts at q0 + j: first byte of the text or the byte behind a newline, unless the row is
// empty ("\n", "\r\n", a lone "\r" at the end) or opens with the comment byte
SQ_HD uint32_t row_start_mask32(const uint8_t* text, uint64_t n, uint64_t q0, uint8_t comment) {
  if (q0 >= n) return 0u;
  const uint32_t nl = newline_mask32(text, n, q0);
  const bool prev_nl = q0 == 0 || text[q0 - 1] == '\n';
  uint32_t cand = ((nl << 1) | (prev_nl ? 1u : 0u)) & ~nl;
  if (q0 + 32 > n) cand &= (1u << uint32_t(n - q0)) - 1u;
  uint32_t m = cand;
  while (m) {
    const int j = lowest_bit(m);
    m &= m - 1;
    const uint64_t p = q0 + j;
    const uint8_t c = text[p];
    const bool skip = (comment != 0 && c == comment) || (c == '\r' && (p + 1 >= n || text[p + 1] == '\n'));
    if (skip) cand &= ~(1u << j);
  }
  return cand;
}

}  // namespace sq
