// CPU harness around sq_scan_row.h (the row-location and row-parse code the kernels of sq_scan.cu run):
// walks the text in the kernels' 32-byte chunks.  Test infrastructure only (tests/test_scan_host.py).
#include <cstdint>

#include "sq_scan_row.h"

extern "C" int64_t scan_host(const uint8_t* text, uint64_t n, uint8_t delim, int has_header, uint8_t comment,
                             int col_key, int col_start, int col_end, int64_t start_minus, int64_t end_minus,
                             uint64_t* key, int32_t* start, int32_t* end, uint64_t* key_off, uint32_t* key_len,
                             uint64_t cap, int* err_kind, int64_t* bad_value, uint64_t* err_off) {
  sq::ScanOpts o;
  o.delim = delim;
  o.comment = comment;
  o.has_header = has_header != 0;
  o.col_key = col_key;
  o.col_start = col_start;
  o.col_end = col_end;
  o.start_minus = start_minus;
  o.end_minus = end_minus;
  uint64_t row = 0, out = 0;
  *err_kind = 0;
  for (uint64_t q0 = 0; q0 < n; q0 += 32) {
    uint32_t m = sq::row_start_mask32(text, n, q0, comment);
    while (m) {
      const int j = sq::lowest_bit(m);
      m &= m - 1;
      const uint64_t r = row++;
      if (o.has_header && r == 0) continue;
      const sq::RowResult res = sq::parse_row(sq::TextSrc{text, n}, q0 + j, o);
      if (res.err != sq::kRowOk) {
        *err_kind = res.err;
        *bad_value = res.bad_value;
        *err_off = q0 + j;
        return -1;
      }
      if (out < cap) {
        key[out] = res.key;
        start[out] = res.start;
        end[out] = res.end;
        key_off[out] = res.key_off;
        key_len[out] = res.key_len;
      }
      ++out;
    }
  }
  return int64_t(out);
}
