"""ORACLE — TEST INFRASTRUCTURE ONLY (ctypes front for oracle/liborc.so).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``sequila-native_b200``) never does.

Functions follow the reference's interval_join.rs (IJ) as restated in
``join_oracle.cpp``; see that file's header for the file:line map.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)


def build(force: bool = False) -> None:
    """Compile liborc.so (and _ref/libsi_ref.so when /root/reference exists)."""
    if force:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_index_build.restype = C.c_void_p
        L.orc_index_build.argtypes = [_u64p, _i32p, _i32p, C.c_uint64, C.c_int32]
        L.orc_index_free.argtypes = [C.c_void_p]
        L.orc_index_bytes.restype = C.c_uint64
        L.orc_index_bytes.argtypes = [C.c_void_p]
        L.orc_index_keys.restype = C.c_uint64
        L.orc_index_keys.argtypes = [C.c_void_p]
        L.orc_index_build_seconds.restype = C.c_double
        L.orc_index_build_seconds.argtypes = [C.c_void_p]
        L.orc_probe.restype = C.c_int64
        L.orc_probe.argtypes = [C.c_void_p, _u64p, _i32p, _i32p, C.c_uint64,
                                C.POINTER(_u32p), C.POINTER(_u32p), _u32p]
        L.orc_probe_counts.restype = C.c_int64
        L.orc_probe_counts.argtypes = [C.c_void_p, _u64p, _i32p, _i32p, C.c_uint64, _u32p]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_gather_i32.argtypes = [_i32p, _u32p, C.c_uint64, _i32p]
        L.orc_brute.restype = C.c_int64
        L.orc_brute.argtypes = [_u64p, _i32p, _i32p, C.c_uint64, _u64p, _i32p, _i32p, C.c_uint64,
                                C.POINTER(_u32p), C.POINTER(_u32p)]
        L.orc_nearest.restype = None
        L.orc_nearest.argtypes = [_u64p, _i32p, _i32p, C.c_uint64, _u64p, _i32p, _i32p, C.c_uint64, _u32p,
                                  C.POINTER(C.c_uint8)]
        L.orc_time_probe.restype = C.c_double
        L.orc_time_probe.argtypes = [C.c_void_p, _u64p, _i32p, _i32p, C.c_uint64, C.c_int32,
                                     C.c_uint32, C.c_int32, C.POINTER(_i32p), C.POINTER(_i32p),
                                     _u64p, _u64p]
        _LIB = L
    return _LIB


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libsi_ref.so"))


def _ref():
    global _REF
    if _REF is None:
        L = C.CDLL(os.path.join(_HERE, "_ref", "libsi_ref.so"))
        L.siref_join.restype = C.c_int64
        L.siref_join.argtypes = [_u64p, _i32p, _i32p, C.c_uint64, _u64p, _i32p, _i32p, C.c_uint64,
                                 C.POINTER(_u32p), C.POINTER(_u32p)]
        L.siref_free.argtypes = [C.c_void_p]
        if hasattr(L, "siref_build"):  # timing entry points (bench.py cpu_baseline)
            L.siref_build.restype = C.c_void_p
            L.siref_build.argtypes = [_u64p, _i32p, _i32p, C.c_uint64]
            L.siref_build_seconds.restype = C.c_double
            L.siref_build_seconds.argtypes = [C.c_void_p]
            L.siref_time_probe.restype = C.c_double
            L.siref_time_probe.argtypes = [C.c_void_p, _u64p, _i32p, _i32p, C.c_uint64, C.c_int32, C.c_uint32, _u64p]
            L.siref_destroy.argtypes = [C.c_void_p]
        _REF = L
    return _REF


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _p(a, t):
    return a.ctypes.data_as(t)


def _take_pairs(lib_free, n, lp, rp):
    left = np.ctypeslib.as_array(lp, shape=(max(n, 1),))[:n].copy()
    right = np.ctypeslib.as_array(rp, shape=(max(n, 1),))[:n].copy()
    lib_free(lp)
    lib_free(rp)
    return left, right


class OracleIndex:
    """Build-side index = IJ:662-683 (bucket by key hash, one coitrees tree per key).

    variant=1 restates coitrees' scalar tree (nosimd.rs); variant=8 its AVX2
    eight-interval-chunk tree (avx.rs).
    """

    def __init__(self, key_hash, start, end, variant: int = 8):
        self._k = _c(key_hash, np.uint64)
        self._s = _c(start, np.int32)
        self._e = _c(end, np.int32)
        assert self._k.shape == self._s.shape == self._e.shape
        self.n_rows = int(self._k.shape[0])
        self._h = _lib().orc_index_build(_p(self._k, _u64p), _p(self._s, _i32p), _p(self._e, _i32p),
                                         self.n_rows, variant)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib().orc_index_free(h)

    @property
    def bytes(self):
        return int(_lib().orc_index_bytes(self._h))

    @property
    def n_keys(self):
        return int(_lib().orc_index_keys(self._h))

    @property
    def build_seconds(self):
        return float(_lib().orc_index_build_seconds(self._h))

    def probe(self, key_hash, start, end):
        """IJ:1582-1618 on one probe batch -> (left_idx u32, right_idx u32, counts u32)."""
        k, s, e = _c(key_hash, np.uint64), _c(start, np.int32), _c(end, np.int32)
        n = int(k.shape[0])
        counts = np.zeros(n, dtype=np.uint32)
        lp, rp = _u32p(), _u32p()
        m = _lib().orc_probe(self._h, _p(k, _u64p), _p(s, _i32p), _p(e, _i32p), n,
                             C.byref(lp), C.byref(rp), _p(counts, _u32p))
        left, right = _take_pairs(_lib().orc_free, int(m), lp, rp)
        return left, right, counts

    def counts(self, key_hash, start, end):
        k, s, e = _c(key_hash, np.uint64), _c(start, np.int32), _c(end, np.int32)
        counts = np.zeros(k.shape[0], dtype=np.uint32)
        _lib().orc_probe_counts(self._h, _p(k, _u64p), _p(s, _i32p), _p(e, _i32p), k.shape[0],
                                _p(counts, _u32p))
        return counts

    def time_probe(self, key_hash, start, end, threads=1, batch_rows=8192, build_cols=None,
                   probe_cols=None, digest=False):
        """CPU-baseline timing of the probe (+ optional 6-column take). Returns (seconds, pairs, digest)."""
        k, s, e = _c(key_hash, np.uint64), _c(start, np.int32), _c(end, np.int32)
        mat = int(build_cols is not None and probe_cols is not None)
        keep = []
        B = (_i32p * 3)()
        P = (_i32p * 3)()
        if mat:
            for i in range(3):
                b, p = _c(build_cols[i], np.int32), _c(probe_cols[i], np.int32)
                keep += [b, p]
                B[i], P[i] = _p(b, _i32p), _p(p, _i32p)
        pairs = C.c_uint64(0)
        dig = C.c_uint64(0)
        sec = _lib().orc_time_probe(self._h, _p(k, _u64p), _p(s, _i32p), _p(e, _i32p), k.shape[0],
                                    int(threads), int(batch_rows), mat, B, P, C.byref(pairs),
                                    C.byref(dig) if digest else None)
        return float(sec), int(pairs.value), int(dig.value)


def join(bkey, bstart, bend, pkey, pstart, pend, variant: int = 8):
    """Whole join through the coitrees restatement -> (left_idx, right_idx, counts)."""
    return OracleIndex(bkey, bstart, bend, variant).probe(pkey, pstart, pend)


def brute(bkey, bstart, bend, pkey, pstart, pend):
    bk, bs, be = _c(bkey, np.uint64), _c(bstart, np.int32), _c(bend, np.int32)
    pk, ps, pe = _c(pkey, np.uint64), _c(pstart, np.int32), _c(pend, np.int32)
    lp, rp = _u32p(), _u32p()
    m = _lib().orc_brute(_p(bk, _u64p), _p(bs, _i32p), _p(be, _i32p), bk.shape[0],
                         _p(pk, _u64p), _p(ps, _i32p), _p(pe, _i32p), pk.shape[0],
                         C.byref(lp), C.byref(rp))
    return _take_pairs(_lib().orc_free, int(m), lp, rp)


def ref_superintervals_join(bkey, bstart, bend, pkey, pstart, pend):
    """The reference's own superintervals C++ library (oracle/_ref/libsi_ref.so)."""
    bk, bs, be = _c(bkey, np.uint64), _c(bstart, np.int32), _c(bend, np.int32)
    pk, ps, pe = _c(pkey, np.uint64), _c(pstart, np.int32), _c(pend, np.int32)
    lp, rp = _u32p(), _u32p()
    m = _ref().siref_join(_p(bk, _u64p), _p(bs, _i32p), _p(be, _i32p), bk.shape[0],
                          _p(pk, _u64p), _p(ps, _i32p), _p(pe, _i32p), pk.shape[0],
                          C.byref(lp), C.byref(rp))
    return _take_pairs(_ref().siref_free, int(m), lp, rp)


class RefSuperIntervalsIndex:
    """The reference's own superintervals library (its `SuperIntervals` algorithm arm, IJ:858-870, 1012-1017),
    compiled from /root/reference into oracle/_ref/libsi_ref.so: build once, time the probe loop."""

    def __init__(self, bkey, bstart, bend):
        bk, bs, be = _c(bkey, np.uint64), _c(bstart, np.int32), _c(bend, np.int32)
        self._h = _ref().siref_build(_p(bk, _u64p), _p(bs, _i32p), _p(be, _i32p), bk.shape[0])
        self.build_seconds = float(_ref().siref_build_seconds(self._h))

    def time_probe(self, pkey, pstart, pend, threads=1, batch_rows=8192):
        pk, ps, pe = _c(pkey, np.uint64), _c(pstart, np.int32), _c(pend, np.int32)
        pairs = C.c_uint64(0)
        sec = _ref().siref_time_probe(self._h, _p(pk, _u64p), _p(ps, _i32p), _p(pe, _i32p), pk.shape[0], int(threads),
                                      int(batch_rows), C.byref(pairs))
        return float(sec), int(pairs.value)

    def __del__(self):
        try:
            _ref().siref_destroy(self._h)
        except Exception:
            pass


def ref_timing_available() -> bool:
    return ref_available() and hasattr(_ref(), "siref_build")


NULL_INDEX = 0xFFFFFFFF


def nearest(bkey, bstart, bend, pkey, pstart, pend):
    """Algorithm::CoitreesNearest (IJ:794-812, 909-956, 972-990, 1593-1602) -> (left_idx with
    NULL_INDEX for a NULL left side, is_overlap flags)."""
    bk, bs, be = _c(bkey, np.uint64), _c(bstart, np.int32), _c(bend, np.int32)
    pk, ps, pe = _c(pkey, np.uint64), _c(pstart, np.int32), _c(pend, np.int32)
    left = np.empty(pk.shape[0], dtype=np.uint32)
    ov = np.zeros(pk.shape[0], dtype=np.uint8)
    _lib().orc_nearest(_p(bk, _u64p), _p(bs, _i32p), _p(be, _i32p), bk.shape[0],
                       _p(pk, _u64p), _p(ps, _i32p), _p(pe, _i32p), pk.shape[0], _p(left, _u32p),
                       ov.ctypes.data_as(C.POINTER(C.c_uint8)))
    return left, ov.astype(bool)


def gather_i32(col, idx):
    c, i = _c(col, np.int32), _c(idx, np.uint32)
    out = np.empty(i.shape[0], dtype=np.int32)
    _lib().orc_gather_i32(_p(c, _i32p), _p(i, _u32p), i.shape[0], _p(out, _i32p))
    return out


_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def pair_digest(left, right, right_offset: int = 0):
    """Order-independent multiset digest of (left,right) pairs: (count, sum mod 2^64, xor) of a
    64-bit mix.  The CUDA library's sq_pairs_digest computes the same three numbers."""
    l = np.asarray(left, dtype=np.uint64)
    r = np.asarray(right, dtype=np.uint64) + np.uint64(right_offset)
    with np.errstate(over="ignore"):
        x = (l << np.uint64(32)) | (r & np.uint64(0xFFFFFFFF))
        x ^= x >> np.uint64(30)
        x *= _M1
        x ^= x >> np.uint64(27)
        x *= _M2
        x ^= x >> np.uint64(31)
        s = int(np.add.reduce(x, dtype=np.uint64)) if x.size else 0
        xo = int(np.bitwise_xor.reduce(x)) if x.size else 0
    return int(l.size), s, xo


def sorted_pairs(left, right):
    """Canonical form for multiset comparison: pairs sorted by (right, left)."""
    l = np.asarray(left, dtype=np.uint64)
    r = np.asarray(right, dtype=np.uint64)
    return np.sort((r << np.uint64(32)) | l)
