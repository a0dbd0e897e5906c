"""Thin object layer over the C ABI (include/sequila_cuda.h).

Mirrors the two halves of the reference's private seam `IntervalJoinAlgorithm`
(interval_join.rs:767 `new`, :957 `get`) at batch granularity:

* :class:`CudaIndex`   = `IntervalJoinAlgorithm::new(&Algorithm::Cuda, hashmap)` — built once per
  query from the concatenated build side (interval_join.rs:662-685);
* :class:`CudaStream`  = what one `IntervalJoinStream` (one per partition, :528-556) calls per probe
  batch instead of the `get` loop (:1586-1618).

numpy arrays go through the host entry points (H2D/D2H inside the call); torch CUDA tensors go
through the ``*_device`` entry points (kernel-level benchmark).  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _native as N


def _check(rc: int, msg_fn):
    if rc != N.SQ_OK:
        msg = msg_fn()
        raise N.SequilaCudaError(rc, msg.decode() if isinstance(msg, bytes) else str(msg))


def _np(a, dtype):
    a = np.asarray(a)
    if a.dtype != dtype or not a.flags.c_contiguous:
        a = np.ascontiguousarray(a, dtype=dtype)
    return a


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a.size else C.c_void_p(0)


def _tptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else C.c_void_p(0)


class CudaContext:
    """sq_ctx: one per (process, device).  Raises if no CUDA device is usable (no CPU fallback)."""

    def __init__(self, device: int = 0):
        self._lib = N.lib()
        h = C.c_void_p()
        rc = self._lib.sq_ctx_create(int(device), C.byref(h))
        if rc != N.SQ_OK:
            raise N.SequilaCudaError(rc, self._lib.sq_last_error(None).decode())
        self._h = h
        self.device = int(device)
        self._fin = weakref.finalize(self, self._lib.sq_ctx_destroy, h)

    def _err(self):
        return self._lib.sq_last_error(self._h)

    def set_option(self, key: str, value) -> None:
        """sq_ctx_set_option: one `sequila.cuda_*` key (SET sequila.cuda_probe_layout TO soa, ...)."""
        _check(self._lib.sq_ctx_set_option(self._h, str(key).encode(), str(value).encode()), self._err)

    def get_option(self, key: str) -> str:
        buf = C.create_string_buffer(64)
        _check(self._lib.sq_ctx_get_option(self._h, str(key).encode(), buf, 64), self._err)
        return buf.value.decode()

    def pinned_empty(self, n: int, dtype) -> np.ndarray:
        """numpy array in pinned host memory (sq_host_alloc); full-speed async copies."""
        dtype = np.dtype(dtype)
        nbytes = max(int(n) * dtype.itemsize, 1)
        p = C.c_void_p()
        _check(self._lib.sq_host_alloc(self._h, nbytes, C.byref(p)), self._err)
        buf = (C.c_char * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(n))
        weakref.finalize(buf, self._lib.sq_host_free, self._h, p)
        return arr

    def pinned_copy(self, a) -> np.ndarray:
        a = np.asarray(a)
        out = self.pinned_empty(a.size, a.dtype)
        out[...] = a.reshape(-1)
        return out


class CudaIndex:
    """sq_index: the flat build-side index (sorted starts + running max end per key segment)."""

    def __init__(self, ctx: CudaContext, handle, keep=None):
        self.ctx = ctx
        self._lib = ctx._lib
        self._h = handle
        self._keep = keep  # device tensors borrowed by add_column_device
        self._fin = weakref.finalize(self, self._lib.sq_index_free, handle)

    @classmethod
    def build(cls, ctx: CudaContext, key_hash, start, end) -> "CudaIndex":
        k, s, e = _np(key_hash, np.uint64), _np(start, np.int32), _np(end, np.int32)
        if not (k.shape == s.shape == e.shape and k.ndim == 1):
            raise ValueError("key_hash/start/end must be 1-d arrays of equal length")
        h = C.c_void_p()
        _check(ctx._lib.sq_index_build(ctx._h, _ptr(k), _ptr(s), _ptr(e), k.shape[0], C.byref(h)), ctx._err)
        return cls(ctx, h)

    @classmethod
    def build_device(cls, ctx: CudaContext, key_hash, start, end, cuda_stream: int = 0) -> "CudaIndex":
        import torch
        assert key_hash.dtype == torch.int64 or key_hash.dtype == torch.uint64
        assert start.dtype == torch.int32 and end.dtype == torch.int32
        assert key_hash.is_cuda and start.is_cuda and end.is_cuda
        h = C.c_void_p()
        _check(ctx._lib.sq_index_build_device(ctx._h, _tptr(key_hash), _tptr(start), _tptr(end),
                                              key_hash.numel(), C.c_void_p(cuda_stream), C.byref(h)), ctx._err)
        return cls(ctx, h)

    def add_column(self, values) -> int:
        v = np.ascontiguousarray(values)
        if v.shape[0] != self.rows:
            raise ValueError("column length differs from the build side")
        width = v.dtype.itemsize if v.ndim == 1 else v.dtype.itemsize * int(np.prod(v.shape[1:]))
        cid = C.c_int32(-1)
        _check(self._lib.sq_index_add_column(self._h, _ptr(v), width, C.byref(cid)), self.ctx._err)
        return int(cid.value)

    def add_column_device(self, tensor) -> int:
        cid = C.c_int32(-1)
        _check(self._lib.sq_index_add_column_device(self._h, _tptr(tensor), tensor.element_size(), C.byref(cid)),
               self.ctx._err)
        self._keep = (self._keep or []) + [tensor]
        return int(cid.value)

    def pack_columns(self, col_ids) -> int:
        """sq_index_pack_columns: up to four 4-byte build columns interleaved row-wise for one-read-per-pair gathers"""
        ids = (C.c_int32 * len(col_ids))(*[int(c) for c in col_ids])
        pid = C.c_int32(-1)
        _check(self._lib.sq_index_pack_columns(self._h, ids, len(col_ids), C.byref(pid)), self.ctx._err)
        return int(pid.value)

    bytes = property(lambda self: int(self._lib.sq_index_bytes(self._h)))
    rows = property(lambda self: int(self._lib.sq_index_rows(self._h)))
    keys = property(lambda self: int(self._lib.sq_index_keys(self._h)))
    build_ms = property(lambda self: float(self._lib.sq_index_build_ms(self._h)))
    uses_packed = property(lambda self: bool(self._lib.sq_index_uses_packed(self._h)))
    uses_rank = property(lambda self: bool(self._lib.sq_index_uses_rank(self._h)))
    sort_key_bits = property(lambda self: int(self._lib.sq_index_sort_key_bits(self._h)))
    uses_positions = property(lambda self: bool(self._lib.sq_index_uses_positions(self._h)))

    def position_rows(self) -> np.ndarray:
        """sq_index_position_rows: build row of every position (indexes built with cuda_build_ids = positions)"""
        out = np.empty(self.rows, dtype=np.uint32)
        _check(self._lib.sq_index_position_rows(self._h, out.ctypes.data_as(C.c_void_p)), self.ctx._err)
        return out

    def position_rows_device_ptr(self) -> int:
        return int(self._lib.sq_index_position_rows_device(self._h) or 0)


class CudaDriver:
    """sq_driver (include/sequila_driver.h): the partition loop of a host in native threads — `n_partitions` streams that
    live as long as the driver; run() makes one pass over a probe side (arrays should be pinned, CudaContext.pinned_*)."""

    def __init__(self, ctx: CudaContext, n_partitions: int):
        self.ctx = ctx
        self._lib = ctx._lib
        h = C.c_void_p()
        _check(self._lib.sq_driver_create(ctx._h, int(n_partitions), C.byref(h)), ctx._err)
        self._h = h
        self.n_partitions = int(n_partitions)
        self._fin = weakref.finalize(self, self._lib.sq_driver_free, h)

    def run(self, index: "CudaIndex", key_hash, start, end, n_tiles: int, flags: int = 0, checksum: bool = False) -> dict:
        k, s, e = _np(key_hash, np.uint64), _np(start, np.int32), _np(end, np.int32)
        st = N.SqDriveStats()
        rc = self._lib.sq_driver_run(self._h, index._h, _ptr(k), _ptr(s), _ptr(e), k.shape[0], int(n_tiles), int(flags),
                                     int(bool(checksum)), None, None, C.byref(st))
        _check(rc, lambda: self._lib.sq_driver_last_error(self._h))
        return {f: getattr(st, f) for f, _ in N.SqDriveStats._fields_}


def _driver_run_ids(self, index, dict_hashes, key_id, start, end, n_tiles: int, flags: int = 0, checksum: bool = False) -> dict:
    """sq_driver_run_ids: one pass with the key column as uint32 ids into `dict_hashes` (12 bytes per row on the wire)"""
    d, k, s, e = _np(dict_hashes, np.uint64), _np(key_id, np.uint32), _np(start, np.int32), _np(end, np.int32)
    st = N.SqDriveStats()
    rc = self._lib.sq_driver_run_ids(self._h, index._h, _ptr(d), d.shape[0], _ptr(k), _ptr(s), _ptr(e), k.shape[0], int(n_tiles),
                                     int(flags), int(bool(checksum)), None, None, C.byref(st))
    _check(rc, lambda: self._lib.sq_driver_last_error(self._h))
    return {f: getattr(st, f) for f, _ in N.SqDriveStats._fields_}


CudaDriver.run_ids = _driver_run_ids


class CudaStream:
    """sq_stream: per-partition probe context (CUDA stream + staging + scratch)."""

    def __init__(self, ctx: CudaContext, cuda_stream: int | None = None):
        self.ctx = ctx
        self._lib = ctx._lib
        h = C.c_void_p()
        if cuda_stream is None:
            _check(self._lib.sq_stream_create(ctx._h, C.byref(h)), ctx._err)
        else:
            _check(self._lib.sq_stream_create_on(ctx._h, C.c_void_p(cuda_stream), C.byref(h)), ctx._err)
        self._h = h
        self._fin = weakref.finalize(self, self._lib.sq_stream_free, h)
        self.n_rows = 0
        self.n_pairs = 0
        self._keep = None

    def _err(self):
        return self._lib.sq_stream_last_error(self._h)

    # ---- host (numpy) entry points: what the DataFusion exec node would call ----------------
    def probe_count(self, index: CudaIndex, key_hash, start, end) -> int:
        k, s, e = _np(key_hash, np.uint64), _np(start, np.int32), _np(end, np.int32)
        if not (k.shape == s.shape == e.shape and k.ndim == 1):
            raise ValueError("key_hash/start/end must be 1-d arrays of equal length")
        n = C.c_uint64(0)
        _check(self._lib.sq_probe_count(self._h, index._h, _ptr(k), _ptr(s), _ptr(e), k.shape[0], C.byref(n)),
               self._err)
        self._keep = index
        self.n_rows, self.n_pairs = int(k.shape[0]), int(n.value)
        return self.n_pairs

    def emit_pairs(self, right: bool = True, counts: bool = True, out=None):
        """-> (left_idx, right_idx | None, counts | None) as numpy uint32 arrays."""
        if out is not None:
            left, r, c = out
        else:
            left = np.empty(self.n_pairs, dtype=np.uint32)
            r = np.empty(self.n_pairs, dtype=np.uint32) if right else None
            c = np.empty(self.n_rows, dtype=np.uint32) if counts else None
        _check(self._lib.sq_probe_emit_pairs(self._h, _ptr(left), _ptr(r) if r is not None else None,
                                             _ptr(c) if c is not None else None, left.shape[0]), self._err)
        return left[:self.n_pairs], (r[:self.n_pairs] if r is not None else None), c

    def probe(self, index: CudaIndex, key_hash, start, end):
        self.probe_count(index, key_hash, start, end)
        return self.emit_pairs()

    def probe_join(self, index: CudaIndex, key_hash, start, end, out):
        """sq_probe_join: one fused call into caller buffers out=(left, right|None, counts|None).
        Raises SequilaCudaError(code 5) with .n_pairs set when `left` is too small."""
        k, s, e = _np(key_hash, np.uint64), _np(start, np.int32), _np(end, np.int32)
        left, r, c = out
        n = C.c_uint64(0)
        rc = self._lib.sq_probe_join(self._h, index._h, _ptr(k), _ptr(s), _ptr(e), k.shape[0], _ptr(left),
                                     _ptr(r) if r is not None else None, _ptr(c) if c is not None else None,
                                     left.shape[0], C.byref(n))
        self._keep = index
        self.n_rows, self.n_pairs = int(k.shape[0]), int(n.value)
        _check(rc, self._err)
        return self.n_pairs

    # ---- asynchronous tile pipeline (sq_stream_submit / sq_stream_collect) ------------------------------
    def submit(self, index: CudaIndex, key_hash, start, end, flags: int = 0) -> int:
        """Enqueue H2D + kernels + D2H of one tile; returns its ticket at once.  The three arrays must stay alive
        (and should be pinned, CudaContext.pinned_*) until the ticket is collected."""
        k, s, e = _np(key_hash, np.uint64), _np(start, np.int32), _np(end, np.int32)
        t = C.c_uint64(0)
        _check(self._lib.sq_stream_submit(self._h, index._h, _ptr(k), _ptr(s), _ptr(e), k.shape[0], int(flags), C.byref(t)),
               self._err)
        self._keep = index
        self._tiles = getattr(self, "_tiles", {})
        self._tiles[int(t.value)] = (k, s, e)
        return int(t.value)

    def set_key_dictionary(self, key_hashes) -> None:
        """sq_stream_set_key_dictionary: hash of every dictionary value, uploaded once; submit_ids() then sends 4-byte ids"""
        d = _np(key_hashes, np.uint64)
        _check(self._lib.sq_stream_set_key_dictionary(self._h, _ptr(d), d.shape[0]), self._err)

    def submit_ids(self, index: CudaIndex, key_id, start, end, flags: int = 0) -> int:
        """sq_stream_submit_ids: as submit(), the key column as uint32 dictionary ids (12 bytes per row on the wire)"""
        k, s, e = _np(key_id, np.uint32), _np(start, np.int32), _np(end, np.int32)
        t = C.c_uint64(0)
        _check(self._lib.sq_stream_submit_ids(self._h, index._h, _ptr(k), _ptr(s), _ptr(e), k.shape[0], int(flags), C.byref(t)),
               self._err)
        self._keep = index
        self._tiles = getattr(self, "_tiles", {})
        self._tiles[int(t.value)] = (k, s, e)
        return int(t.value)

    def collect(self, ticket: int):
        """Wait for the oldest ticket -> (n_pairs, left_idx, right_idx | None, counts | None): numpy views of pinned
        buffers from the library's pool; each goes back to the pool when its array is garbage collected."""
        out = N.SqTileOut()
        rc = self._lib.sq_stream_collect(self._h, int(ticket), C.byref(out))
        getattr(self, "_tiles", {}).pop(int(ticket), None)
        _check(rc, self._err)

        def view(addr, n, width=4):
            if not addr:
                return None
            ct, dt = (C.c_uint8, np.uint8) if width == 1 else (C.c_uint32, np.uint32)
            buf = (ct * max(int(n), 1)).from_address(addr)
            arr = np.frombuffer(buf, dtype=dt, count=int(n))
            weakref.finalize(buf, self._lib.sq_host_free, self.ctx._h, C.c_void_p(addr))
            return arr
        self.n_rows, self.n_pairs = int(out.n_rows), int(out.n_pairs)
        # counts: uint32, or uint8 for a tile submitted with TILE_COUNTS_U8 whose counts all fit a byte (sq_tile_out.counts_width)
        return (int(out.n_pairs), view(out.left_idx, out.n_pairs), view(out.right_idx, out.n_pairs),
                view(out.counts, out.n_rows, int(out.counts_width)))

    @property
    def in_flight(self) -> int:
        return int(self._lib.sq_stream_in_flight(self._h))

    def pipeline_stats(self) -> dict:
        o = (C.c_double * 8)()
        _check(self._lib.sq_stream_pipeline_stats(self._h, o), self._err)
        return {"h2d_ms": o[0], "kernel_ms": o[1], "d2h_ms": o[2], "h2d_bytes": int(o[3]), "d2h_bytes": int(o[4]),
                "tiles": int(o[5]), "regrown": int(o[6]), "pairs_per_row": o[7]}

    # ---- bounded output (low-memory protocol) -------------------------------------------------
    def counts(self) -> np.ndarray:
        out = np.empty(self.n_rows, dtype=np.uint32)
        _check(self._lib.sq_stream_counts(self._h, _ptr(out)), self._err)
        return out

    def set_window(self, pair_offset: int, n_pairs: int):
        _check(self._lib.sq_stream_set_window(self._h, int(pair_offset), int(n_pairs)), self._err)
        self._win = int(n_pairs)

    def fetch_pairs(self, n: int):
        left, right = np.empty(n, dtype=np.uint32), np.empty(n, dtype=np.uint32)
        _check(self._lib.sq_fetch_pairs(self._h, _ptr(left), _ptr(right), n), self._err)
        return left, right

    def probe_nearest(self, index: CudaIndex, key_hash, start, end) -> np.ndarray:
        """sq_probe_nearest: one build row (or N.NULL_INDEX) per probe row."""
        k, s, e = _np(key_hash, np.uint64), _np(start, np.int32), _np(end, np.int32)
        out = np.empty(k.shape[0], dtype=np.uint32)
        _check(self._lib.sq_probe_nearest(self._h, index._h, _ptr(k), _ptr(s), _ptr(e), k.shape[0], _ptr(out)), self._err)
        self._keep = index
        self.n_rows = self.n_pairs = int(k.shape[0])
        return out

    def gather_build(self, col_id: int, dtype, width: int | None = None) -> np.ndarray:
        dtype = np.dtype(dtype)
        out = np.empty(self.n_pairs, dtype=dtype)
        _check(self._lib.sq_gather_column(self._h, 0, col_id, None, dtype.itemsize, _ptr(out), self.n_pairs),
               self._err)
        return out

    def gather_probe(self, values) -> np.ndarray:
        v = np.ascontiguousarray(values)
        if v.shape[0] != self.n_rows:
            raise ValueError("probe column length differs from the probe tile")
        out = np.empty(self.n_pairs, dtype=v.dtype)
        _check(self._lib.sq_gather_column(self._h, 1, -1, _ptr(v), v.dtype.itemsize, _ptr(out), self.n_pairs),
               self._err)
        return out

    def cast_i64_to_i32(self, values, minus: int = 0) -> np.ndarray:
        v = _np(values, np.int64)
        out = np.empty(v.shape[0], dtype=np.int32)
        _check(self._lib.sq_cast_i64_to_i32(self._h, _ptr(v), v.shape[0], int(minus), _ptr(out)), self._err)
        return out

    # ---- device (torch) entry points: kernel-level benchmark ---------------------------------
    def probe_count_device(self, index: CudaIndex, key_hash, start, end) -> int:
        n = C.c_uint64(0)
        _check(self._lib.sq_probe_count_device(self._h, index._h, _tptr(key_hash), _tptr(start), _tptr(end),
                                               key_hash.numel(), C.byref(n)), self._err)
        self._keep = (index, key_hash, start, end)
        self.n_rows, self.n_pairs = int(key_hash.numel()), int(n.value)
        return self.n_pairs

    def probe_join_device(self, index: CudaIndex, key_hash, start, end, left, right=None) -> int:
        """sq_probe_join_device: fused count->scan->write pass into caller tensors."""
        n = C.c_uint64(0)
        rc = self._lib.sq_probe_join_device(self._h, index._h, _tptr(key_hash), _tptr(start), _tptr(end),
                                            key_hash.numel(), _tptr(left), _tptr(right) if right is not None else None,
                                            left.numel(), C.byref(n))
        self._keep = (index, key_hash, start, end)
        self.n_rows, self.n_pairs = int(key_hash.numel()), int(n.value)
        _check(rc, self._err)
        return self.n_pairs

    def emit_pairs_device(self, left, right=None):
        _check(self._lib.sq_probe_emit_pairs_device(self._h, _tptr(left), _tptr(right) if right is not None else None,
                                                    left.numel()), self._err)

    def gather_build_device(self, col_id: int, out):
        _check(self._lib.sq_gather_column_device(self._h, 0, col_id, None, out.element_size(), _tptr(out),
                                                 out.numel()), self._err)

    def gather_probe_device(self, values, out):
        _check(self._lib.sq_gather_column_device(self._h, 1, -1, _tptr(values), values.element_size(), _tptr(out),
                                                 out.numel()), self._err)

    def gather_pack_device(self, pack_id: int, outs):
        """every column of a pack for the pairs of the last emit, one pass (outs: torch CUDA tensors of 4-byte elements)"""
        ptrs = (C.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
        _check(self._lib.sq_gather_pack_device(self._h, int(pack_id), ptrs, len(outs), min(o.numel() for o in outs)), self._err)

    def gather_probe_columns_device(self, values, outs):
        """up to four 4-byte probe columns of this tile for the pairs of the last emit, one pass"""
        vin = (C.c_void_p * len(values))(*[v.data_ptr() for v in values])
        ptrs = (C.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
        _check(self._lib.sq_gather_probe_columns_device(self._h, vin, ptrs, len(outs), min(o.numel() for o in outs)), self._err)

    def counts_device_ptr(self) -> int:
        return int(self._lib.sq_stream_counts_device(self._h) or 0)

    def digest_device(self, left, right, n_pairs: int, right_offset: int = 0):
        out = (C.c_uint64 * 3)()
        _check(self._lib.sq_pairs_digest_device(self._h, _tptr(left), _tptr(right), int(n_pairs), int(right_offset),
                                                out), self._err)
        return int(out[0]), int(out[1]), int(out[2])

    # ---- instrumentation ------------------------------------------------------------------------
    def set_profiling(self, enabled: bool = True):
        _check(self._lib.sq_stream_set_profiling(self._h, int(enabled)), self._err)

    def phase_ms(self):
        out = (C.c_float * 5)()
        _check(self._lib.sq_stream_phase_ms(self._h, out), self._err)
        return {"h2d": out[0], "join": out[1], "emit": out[2], "d2h": out[3], "gather": out[4]}

    launches = property(lambda self: int(self._lib.sq_stream_launches(self._h)))
    bytes = property(lambda self: int(self._lib.sq_stream_bytes(self._h)))
