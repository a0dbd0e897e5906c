"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the scan side (include/sequila_scan.h):
delimited text -> (key hash, start, end, dictionary ids) the way the reference's tables are read.

Reference behaviour restated (file:line under /root/reference):
  queries/q1-coitrees.sql:6-14      `CREATE EXTERNAL TABLE .. (contig VARCHAR NOT NULL, start BIGINT NOT NULL,
                                    end BIGINT NOT NULL) STORED AS CSV .. OPTIONS ('delimiter' '\\t', 'has_header'
                                    'false')`: DataFusion's CSV reader (arrow-csv 56.2.0, Cargo.lock) splits rows
                                    at '\\n' / '\\r\\n', fields at the delimiter, parses BIGINT fields as
                                    [+-]?[0-9]+ and skips the header row when told so
  sequila/sequila-core/src/physical_planner/joins/interval_join.rs:1661-1672, 1959-1965
                                    `evaluate_as_i32`: checked cast to Int32, the error text on overflow
  .../interval_join.rs:1037, 1211   one u64 hash per row from the `on` column (any injective function gives the
                                    same join; here the library's own sq_keyhash.h, restated below)
Pinned by tests/test_oracle_golden.py: the reference's own fixtures (testing/data/interval/reads.csv,
targets.csv, bytes committed in tests/golden/reference_tables.json) parse to the rows the reference's tests
list, and agree with Python's csv module.

Only tests/ may import this module; the product package never does.
"""
from __future__ import annotations

import numpy as np

_MASK = (1 << 64) - 1
_GOLDEN = 0x9E3779B97F4A7C15
_FNV_OFFSET = 0xCBF29CE484222325
_FNV_PRIME = 0x100000001B3
CAST_ERROR = "Arrow error: Cast error: Can't cast value {} to type Int32"


def mix64(x: int) -> int:
    x &= _MASK
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & _MASK
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & _MASK
    x ^= x >> 31
    return x


SEED = mix64(1)


def fold(acc: int, h: int) -> int:
    return mix64(acc ^ ((mix64(h) + _GOLDEN) & _MASK))


def string_hash(b: bytes) -> int:
    """sq_keyhash.h `of_string`: <= 8 bytes from the packed word and the length, longer ones FNV-1a."""
    if len(b) <= 8:
        raw = int.from_bytes(b, "little")
        return mix64(raw ^ ((_GOLDEN * (len(b) + 1)) & _MASK))
    h = _FNV_OFFSET
    for c in b:
        h = ((h ^ c) * _FNV_PRIME) & _MASK
    return h


def key_hash_of(b: bytes) -> int:
    """row hash of a single Utf8 `on` column"""
    return fold(SEED, string_hash(b))


class ScanError(Exception):
    def __init__(self, kind: str, message: str):
        super().__init__(message)
        self.kind = kind  # "cast" | "parse"


def _rows(text: bytes, comment):
    for raw in text.split(b"\n"):
        if raw.endswith(b"\r"):
            raw = raw[:-1]
        if not raw:
            continue
        if comment and raw[:1] == comment:
            continue
        yield raw


def _bigint(field: bytes):
    body = field[1:] if field[:1] in (b"+", b"-") else field
    if not body or not body.isdigit() or not body.isascii():
        return None
    v = int(field)
    if not (-(1 << 63) < v < (1 << 63)):
        return None
    return v


def scan_delimited(text: bytes, delimiter: bytes = b"\t", has_header: bool = False, comment: bytes | None = None,
                   col_key: int | None = 0, col_start: int = 1, col_end: int = 2, start_minus: int = 0,
                   end_minus: int = 0):
    """-> dict(key_hash u64[n], start i32[n], end i32[n], ids u32[n] | None, dictionary [bytes])"""
    text = bytes(text)
    keys, starts, ends, ids, dictionary, seen = [], [], [], [], [], {}
    need = max(col_start, col_end, -1 if col_key is None else col_key) + 1
    first = True
    for raw in _rows(text, comment):
        if first and has_header:
            first = False
            continue
        first = False
        f = raw.split(delimiter)
        if len(f) < need:
            raise ScanError("parse", f"{len(f)} field(s), need {need}: {raw!r}")
        # a quoted field may hide the delimiter (DataFusion's CSV reader honours quotes): every field up to the last one the
        # table names — the ones skipped on the way included — must be unquoted, else the columns would shift silently
        if any(f[c][:1] == b'"' for c in range(need)):
            raise ScanError("parse", f"quoted field: {raw!r}")
        vals = []
        for c, minus in ((col_start, start_minus), (col_end, end_minus)):
            v = _bigint(f[c])
            if v is None:
                raise ScanError("parse", f"field {c} is not a BIGINT: {raw!r}")
            vals.append(v - minus)
        for v in vals:  # start before end (interval_join.rs:1039-1040)
            if not (-(1 << 31) <= v < (1 << 31)):
                raise ScanError("cast", CAST_ERROR.format(v))
        starts.append(vals[0])
        ends.append(vals[1])
        if col_key is None:
            keys.append(SEED)
        else:
            k = f[col_key]
            if k not in seen:
                seen[k] = (len(dictionary), key_hash_of(k))
                dictionary.append(k)
            i, h = seen[k]
            ids.append(i)
            keys.append(h)
    return {
        "key_hash": np.array(keys, dtype=np.uint64),
        "start": np.array(starts, dtype=np.int32),
        "end": np.array(ends, dtype=np.int32),
        "ids": None if col_key is None else np.array(ids, dtype=np.uint32),
        "dictionary": dictionary,
    }


def render(contigs, start, end, delimiter: bytes = b"\t", header: bytes | None = None, eol: bytes = b"\n",
           extra=None, final_eol: bool = True) -> bytes:
    """Text of a table whose columns are known: the parse of `render(x)` must give x back."""
    rows = []
    for i in range(len(start)):
        f = [contigs[i], str(int(start[i])).encode(), str(int(end[i])).encode()]
        if extra is not None:
            f += extra[i]
        rows.append(delimiter.join(f))
    body = eol.join(([header] if header is not None else []) + rows)
    return body + (eol if final_eol and rows else b"")
