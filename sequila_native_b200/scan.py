"""Scan side of the join over the C ABI of include/sequila_scan.h: delimited text (BED / CSV) is parsed on
the device straight into the columns the join consumes (key hash, start, end as Int32, dictionary ids of
the key column).  This is the `CREATE EXTERNAL TABLE .. STORED AS CSV .. OPTIONS ('delimiter' '\\t',
'has_header' 'false')` of the reference's queries (queries/q1-coitrees.sql:6-14) without the host-side
Arrow batches, `create_hashes` on strings (interval_join.rs:1037, 1211) or `evaluate_as_i32`
(interval_join.rs:1661-1672).  Nothing here computes on the CPU."""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _native as N
from .cuda_join import CudaContext, CudaIndex, CudaStream, _check


class CudaScan:
    """sq_scan: the device-resident columns of one scanned table."""

    def __init__(self, stream: CudaStream, handle):
        self.stream = stream
        self._lib = stream._lib
        self._h = handle
        self._fin = weakref.finalize(self, self._lib.sq_scan_free, handle)

    @staticmethod
    def _options(delimiter, has_header, comment, col_key, col_start, col_end, start_minus, end_minus):
        o = N.SqScanOptions()
        o.delimiter = ord(delimiter) if isinstance(delimiter, str) else int(delimiter)
        o.has_header = 1 if has_header else 0
        o.comment = (ord(comment) if isinstance(comment, str) else int(comment)) if comment else 0
        o.col_key = -1 if col_key is None else int(col_key)
        o.col_start, o.col_end = int(col_start), int(col_end)
        o.start_minus, o.end_minus = int(start_minus), int(end_minus)
        return o

    @classmethod
    def from_text(cls, stream: CudaStream, text, delimiter="\t", has_header=False, comment=None, col_key=0,
                  col_start=1, col_end=2, start_minus=0, end_minus=0) -> "CudaScan":
        """`text`: bytes-like host buffer, or a torch CUDA uint8 tensor (kernel-level timing without PCIe)."""
        o = cls._options(delimiter, has_header, comment, col_key, col_start, col_end, start_minus, end_minus)
        h = C.c_void_p()
        if hasattr(text, "is_cuda"):
            assert text.is_cuda and text.element_size() == 1
            rc = stream._lib.sq_scan_text_device(stream._h, C.c_void_p(text.data_ptr() if text.numel() else 0),
                                                 text.numel(), C.byref(o), C.byref(h))
        else:
            buf = np.frombuffer(text, dtype=np.uint8)
            rc = stream._lib.sq_scan_text(stream._h, C.c_void_p(buf.ctypes.data if buf.size else 0), buf.size,
                                          C.byref(o), C.byref(h))
        _check(rc, stream._err)
        return cls(stream, h)

    @classmethod
    def from_file(cls, stream: CudaStream, path: str, **kw) -> "CudaScan":
        return cls.from_text(stream, np.fromfile(path, dtype=np.uint8), **kw)

    rows = property(lambda self: int(self._lib.sq_scan_rows(self._h)))
    bytes = property(lambda self: int(self._lib.sq_scan_bytes(self._h)))
    key_hash_ptr = property(lambda self: self._lib.sq_scan_key_hash_device(self._h) or 0)
    start_ptr = property(lambda self: self._lib.sq_scan_start_device(self._h) or 0)
    end_ptr = property(lambda self: self._lib.sq_scan_end_device(self._h) or 0)
    key_ids_ptr = property(lambda self: self._lib.sq_scan_key_ids_device(self._h) or 0)

    @property
    def dictionary(self) -> list:
        """distinct key strings (bytes), index = dictionary id = order of first occurrence"""
        out = []
        for i in range(int(self._lib.sq_scan_dict_size(self._h))):
            p, n = C.c_void_p(), C.c_uint32()
            self._lib.sq_scan_dict_entry(self._h, i, C.byref(p), C.byref(n), None)
            out.append(C.string_at(p, n.value) if n.value else b"")
        return out

    @property
    def dictionary_hashes(self) -> np.ndarray:
        out = np.zeros(int(self._lib.sq_scan_dict_size(self._h)), dtype=np.uint64)
        for i in range(out.size):
            h = C.c_uint64()
            self._lib.sq_scan_dict_entry(self._h, i, None, None, C.byref(h))
            out[i] = h.value
        return out

    @property
    def timing_ms(self):
        """(h2d of the text, locate + parse kernels, id assignment) of the scan, CUDA events"""
        out = (C.c_float * 3)()
        self._lib.sq_scan_timing(self._h, out)
        return tuple(float(x) for x in out)

    def fetch(self, ids: bool = True):
        """(key_hash, start, end[, key_ids]) as numpy arrays (device -> host copy)"""
        n = self.rows
        k, s, e = np.empty(n, np.uint64), np.empty(n, np.int32), np.empty(n, np.int32)
        want_ids = ids and (self.key_ids_ptr != 0 or n == 0) and self._lib.sq_scan_dict_size(self._h) > 0
        i = np.empty(n, np.uint32) if want_ids else None
        vp = lambda a: C.c_void_p(a.ctypes.data if a is not None and a.size else 0)
        _check(self._lib.sq_scan_fetch(self.stream._h, self._h, vp(k), vp(s), vp(e), vp(i)), self.stream._err)
        return (k, s, e, i) if ids else (k, s, e)

    # ---- straight into the join: the columns never leave the device -------------------------------------
    def build_index(self, ctx: CudaContext, with_ids_column: bool = False):
        """sq_index_build_device over the scanned columns (= collect_left_input, interval_join.rs:597-689)."""
        h = C.c_void_p()
        _check(ctx._lib.sq_index_build_device(ctx._h, C.c_void_p(self.key_hash_ptr), C.c_void_p(self.start_ptr),
                                              C.c_void_p(self.end_ptr), self.rows, C.c_void_p(0),  # the scan is complete
                                              C.byref(h)), ctx._err)
        idx = CudaIndex(ctx, h, keep=[self])
        if with_ids_column:
            cid = C.c_int32(-1)
            _check(ctx._lib.sq_index_add_column_device(h, C.c_void_p(self.key_ids_ptr), 4, C.byref(cid)), ctx._err)
            return idx, int(cid.value)
        return idx

    def probe_count(self, stream: CudaStream, index: CudaIndex) -> int:
        """sq_probe_count_device with this table as the probe side (`select count(1)` of q1-coitrees.sql:16-19)."""
        if self.rows > 0xFFFFFFFF:
            raise ValueError("probe tile too large")
        n = C.c_uint64()
        _check(stream._lib.sq_probe_count_device(stream._h, index._h, C.c_void_p(self.key_hash_ptr),
                                                 C.c_void_p(self.start_ptr), C.c_void_p(self.end_ptr), self.rows,
                                                 C.byref(n)), stream._err)
        return int(n.value)
