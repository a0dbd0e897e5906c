"""CPU: the scan side.
 1. the scan oracle (oracle/scan_oracle.py) is pinned to the reference's own fixtures: the bytes of
    testing/data/interval/reads.csv / targets.csv parse to the rows the reference's tests list, and to what
    Python's csv module reads;
 2. the row-location and row-parse code the CUDA kernels run (sequila_native_b200/csrc/sq_scan_row.h, host +
    device) is compiled with g++ into a harness and compared with the oracle on the fixtures, hand-written
    edge cases and random tables — so only the CUDA-specific parts (block scan, dictionary table, SIMD
    newline mask) are left to the GPU tests (tests/test_gpu_scan.py)."""
import csv
import ctypes as C
import io
import os
import subprocess

import numpy as np
import pytest

from oracle import scan_oracle as SO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("scan_shim") / "scan_host_shim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC",
                    "-I" + os.path.join(ROOT, "sequila_native_b200", "csrc"), "-o", out,
                    os.path.join(ROOT, "tests", "scan_host_shim.cpp")], check=True)
    lib = C.CDLL(out)
    lib.scan_host.restype = C.c_int64
    lib.scan_host.argtypes = [C.c_void_p, C.c_uint64, C.c_uint8, C.c_int, C.c_uint8, C.c_int, C.c_int, C.c_int,
                              C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]
    return lib


def host_scan(lib, text: bytes, delimiter=b"\t", has_header=False, comment=None, col_key=0, col_start=1, col_end=2,
              start_minus=0, end_minus=0):
    buf = np.frombuffer(text, dtype=np.uint8)
    cap = text.count(b"\n") + 2
    key, start, end = np.zeros(cap, np.uint64), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    koff, klen = np.zeros(cap, np.uint64), np.zeros(cap, np.uint32)
    kind, bad, off = C.c_int(0), C.c_int64(0), C.c_uint64(0)
    n = lib.scan_host(buf.ctypes.data if buf.size else None, buf.size, delimiter[0], int(has_header),
                      comment[0] if comment else 0, -1 if col_key is None else col_key, col_start, col_end,
                      start_minus, end_minus, key.ctypes.data, start.ctypes.data, end.ctypes.data, koff.ctypes.data,
                      klen.ctypes.data, cap, C.byref(kind), C.byref(bad), C.byref(off))
    if n < 0:
        raise SO.ScanError("cast" if kind.value in (5, 6) else "parse",
                           SO.CAST_ERROR.format(bad.value) if kind.value in (5, 6) else f"kind {kind.value} at {off.value}")
    keys = [text[int(o):int(o) + int(l)] for o, l in zip(koff[:n], klen[:n])]
    return {"key_hash": key[:n], "start": start[:n], "end": end[:n], "keys": keys}


def same(h, o):
    assert np.array_equal(h["key_hash"], o["key_hash"])
    assert np.array_equal(h["start"], o["start"]) and np.array_equal(h["end"], o["end"])
    if o["ids"] is not None:
        assert h["keys"] == [o["dictionary"][i] for i in o["ids"]]


CSV = dict(delimiter=b",", has_header=True)


@pytest.mark.parametrize("name", ["reads", "targets"])
def test_oracle_parses_the_reference_fixtures(golden, name):
    text = golden[name + "_csv_text"].encode()
    got = SO.scan_delimited(text, **CSV)
    want = golden[name]  # rows listed by the reference's own tests (tests/golden/make_golden.py)
    assert [[got["dictionary"][i].decode(), int(s), int(e)] for i, s, e in zip(got["ids"], got["start"], got["end"])] == want
    rows = list(csv.reader(io.StringIO(text.decode())))[1:]
    assert [[r[0], int(r[1]), int(r[2])] for r in rows if r] == want
    assert got["dictionary"] == [b"chr1", b"chr2"]
    # one hash per distinct contig, the library's key hash of the bytes
    assert len(set(got["key_hash"].tolist())) == 2
    assert int(got["key_hash"][0]) == SO.key_hash_of(b"chr1")


@pytest.mark.parametrize("name", ["reads", "targets"])
def test_kernel_row_code_on_the_reference_fixtures(shim, golden, name):
    text = golden[name + "_csv_text"].encode()
    same(host_scan(shim, text, **CSV), SO.scan_delimited(text, **CSV))


EDGE_TEXTS = [
    b"",
    b"\n",
    b"\n\n\r\n",
    b"chr1\t1\t2",
    b"chr1\t1\t2\n",
    b"chr1\t1\t2\r\n",
    b"chr1\t1\t2\r",
    b"\nchr1\t1\t2\n\nchr2\t3\t4\n\n",
    b"chr1\t+5\t-7\nchrUn_KI270442v1\t0\t2147483647\nchr1\t-2147483648\t9\n",
    b"#track\nchr1\t1\t2\n#x\nchr2\t5\t6",
    b"a\t1\t2\tname\t0\t+\nb\t3\t4\tname2\t1\t-\n",
    b"x" * 40 + b"\t12\t34\n" + b"y" * 31 + b"\t1\t2\n" + b"z\t7\t8\n",
    b"\t1\t2\n",  # empty key
    b"12345678\t1\t2\n123456789\t1\t2\n1234567\t1\t2\n",  # around the 8-byte key path
]


@pytest.mark.parametrize("i", range(len(EDGE_TEXTS)))
def test_kernel_row_code_edge_texts(shim, i):
    text = EDGE_TEXTS[i]
    kw = dict(comment=b"#")
    same(host_scan(shim, text, **kw), SO.scan_delimited(text, **kw))


def test_column_choice_header_and_minus(shim):
    text = b"s,e,name,contig\n10,20,a,chr1\n30,40,b,chr2\r\n50,60,c,chr1\n"
    kw = dict(delimiter=b",", has_header=True, col_key=3, col_start=0, col_end=1, end_minus=1)
    o = SO.scan_delimited(text, **kw)
    assert o["end"].tolist() == [19, 39, 59] and o["dictionary"] == [b"chr1", b"chr2"] and o["ids"].tolist() == [0, 1, 0]
    same(host_scan(shim, text, **kw), o)
    kw = dict(delimiter=b",", has_header=True, col_key=None, col_start=0, col_end=1)
    o = SO.scan_delimited(text, **kw)
    assert set(o["key_hash"].tolist()) == {SO.SEED}  # range-only join: the key of on=[(1,1)]
    same(host_scan(shim, text, **kw), o)


BAD = [
    (b"chr1\t1\n", "parse"),
    (b"chr1\t1\t2\nchr1\n", "parse"),
    (b"chr1\tx\t2\n", "parse"),
    (b"chr1\t1\t\n", "parse"),
    (b"chr1\t1 \t2\n", "parse"),
    (b"chr1\t1\t2.5\n", "parse"),
    (b"chr1\t-\t2\n", "parse"),
    (b"chr1\t1-\t2\n", "parse"),
    (b'"chr1"\t1\t2\n', "parse"),
    (b"chr1\t99999999999999999999\t2\n", "parse"),
    (b"chr1\t2147483648\t2\n", "cast"),
    (b"chr1\t1\t-2147483649\n", "cast"),
    (b"chr1\t1\t2\nchr1\t5\t3000000000\nchr1\t4000000000\t1\n", "cast"),
    (b"chr1\t3000000000\tabc\n", "parse"),  # the reader fails before the cast is evaluated
]


@pytest.mark.parametrize("text,kind", BAD)
def test_errors(shim, golden, text, kind):
    with pytest.raises(SO.ScanError) as eo:
        SO.scan_delimited(text)
    with pytest.raises(SO.ScanError) as eh:
        host_scan(shim, text)
    assert eo.value.kind == eh.value.kind == kind
    if kind == "cast":  # the reference's text (interval_join.rs:1959-1965), first offending value
        assert str(eo.value) == str(eh.value)
        assert str(eo.value).startswith(golden["cast_error_format"].split("{}")[0])


def test_cast_error_names_start_before_end(shim):
    text = b"5000000000,c,4000000000\n"  # end column first in the text, both overflow
    kw = dict(delimiter=b",", col_key=1, col_start=2, col_end=0)
    with pytest.raises(SO.ScanError) as eo:
        SO.scan_delimited(text, **kw)
    with pytest.raises(SO.ScanError) as eh:
        host_scan(shim, text, **kw)
    assert str(eo.value) == str(eh.value) == SO.CAST_ERROR.format(4000000000)


def random_table(rng, n, crlf=False, blanks=False, extra=False, long_names=False):
    names = [b"chr%d" % i for i in range(1, 23)] + [b"chrX", b"chrY", b"chrM"]
    if long_names:
        names += [b"chrUn_KI270%03dv1" % i for i in range(40)] + [b"HLA-DRB1*15:01:01:0%d" % i for i in range(5)]
    ids = rng.integers(0, len(names), n)
    start = rng.integers(-1000, 250_000_000, n)
    end = start + rng.integers(0, 100_000, n)
    rows = []
    for i in range(n):
        f = [names[ids[i]], b"%d" % start[i], b"%d" % end[i]]
        if extra:
            f += [b"name%d" % i, b"0", b"+-"[i & 1:(i & 1) + 1]]
        rows.append(b"\t".join(f))
        if blanks and rng.random() < 0.05:
            rows.append(b"" if rng.random() < 0.5 else b"#comment line %d" % i)
    eol = b"\r\n" if crlf else b"\n"
    text = eol.join(rows) + (eol if rng.random() < 0.5 else b"")
    return text, names, ids, start, end


@pytest.mark.parametrize("seed", range(8))
def test_kernel_row_code_random_tables(shim, seed):
    rng = np.random.default_rng(seed)
    text, names, ids, start, end = random_table(rng, 3000, crlf=bool(seed & 1), blanks=bool(seed & 2), extra=bool(seed & 4),
                                                long_names=seed >= 4)
    kw = dict(comment=b"#")
    o = SO.scan_delimited(text, **kw)
    assert np.array_equal(o["start"], start.astype(np.int32)) and np.array_equal(o["end"], end.astype(np.int32))
    assert [o["dictionary"][i] for i in o["ids"]] == [names[i] for i in ids]
    same(host_scan(shim, text, **kw), o)


def test_render_round_trip():
    rng = np.random.default_rng(3)
    _, names, ids, start, end = random_table(rng, 500)
    text = SO.render([names[i] for i in ids], start, end, delimiter=b",", header=b"contig,pos_start,pos_end")
    o = SO.scan_delimited(text, delimiter=b",", has_header=True)
    assert np.array_equal(o["start"], start.astype(np.int32)) and np.array_equal(o["end"], end.astype(np.int32))


@pytest.mark.parametrize("text", [b'7,"a,b",chr1,10,20\n', b'id1,x,chr1,10,20\n8,"q",chr2,1,2\n', b'"r",y,chr1,10,20\r\n'])
def test_quoted_skipped_field_is_rejected_not_shifted(shim, text):
    """a quoted field BEFORE the named columns may hide the delimiter (`id,"a,b",chr1,10,20`): DataFusion's CSV reader
    honours quotes; this scanner rejects the row instead of reading shifted columns.  Quotes after the last named column are
    never looked at."""
    kw = dict(delimiter=b",", col_key=2, col_start=3, col_end=4)
    with pytest.raises(SO.ScanError) as eo:
        SO.scan_delimited(text, **kw)
    with pytest.raises(SO.ScanError) as eh:
        host_scan(shim, text, **kw)
    assert eo.value.kind == eh.value.kind == "parse"
    ok = b'7,a,chr1,10,20,"trailing, quoted"\n'
    same(host_scan(shim, ok, **kw), SO.scan_delimited(ok, **kw))
