"""GPU parity of the scan side (include/sequila_scan.h): delimited text -> device columns
through the C ABI, bit-exact against oracle/scan_oracle.py (pinned to the reference's fixtures in
tests/test_scan_host.py), then straight into the join: the q1-coitrees.sql shape from file bytes to the
reference's golden result tables without a host-side table in between."""
import os

import numpy as np
import pytest

import sequila_native_b200 as sn
from oracle import scan_oracle as SO
from helpers import sort_rows
from test_scan_host import BAD, EDGE_TEXTS, random_table

pytestmark = pytest.mark.gpu

CSV = dict(delimiter=",", has_header=True)


def okw(kw):
    """CudaScan keyword arguments -> the oracle's (bytes instead of str)"""
    out = dict(kw)
    for k in ("delimiter", "comment"):
        if isinstance(out.get(k), str):
            out[k] = out[k].encode()
    return out


def check(stream, text, **kw):
    sc = sn.CudaScan.from_text(stream, text, **kw)
    o = SO.scan_delimited(text, **okw(kw))
    k, s, e, ids = sc.fetch()
    assert sc.rows == len(o["start"])
    assert np.array_equal(k, o["key_hash"])
    assert np.array_equal(s, o["start"]) and np.array_equal(e, o["end"])
    if o["ids"] is not None and sc.rows:
        assert sc.dictionary == o["dictionary"]  # ids in order of first occurrence
        assert np.array_equal(ids, o["ids"])
        assert [int(h) for h in sc.dictionary_hashes] == [SO.key_hash_of(b) for b in o["dictionary"]]
    return sc, o


@pytest.fixture()
def stream(cuda_ctx):
    return sn.CudaStream(cuda_ctx)


@pytest.mark.parametrize("name", ["reads", "targets"])
def test_reference_fixture_bytes(stream, golden, name):
    sc, o = check(stream, golden[name + "_csv_text"].encode(), **CSV)
    rows = [[sc.dictionary[i].decode(), int(s), int(e)] for i, s, e in zip(o["ids"], o["start"], o["end"])]
    assert rows == golden[name]


def test_q1_shape_from_file_bytes_to_the_golden_tables(cuda_ctx, stream, golden):
    """reads.csv JOIN targets.csv ON contig AND closed overlap: 16 rows (integration_test.rs:42-65); without the
    contig key: 32 rows (integration_test.rs:122-161) — both sides scanned on the device, columns never on the host."""
    for key_col, want in ((0, golden["equi_rows"]), (None, golden["range_rows"])):
        a = sn.CudaScan.from_text(stream, golden["reads_csv_text"].encode(), col_key=key_col, **CSV)
        b = sn.CudaScan.from_text(stream, golden["targets_csv_text"].encode(), col_key=key_col, **CSV)
        idx = a.build_index(cuda_ctx)
        assert b.probe_count(stream, idx) == len(want)
        # the pairs themselves, through the host entry points on the fetched columns of the same scans
        k, s, e = b.fetch(ids=False)
        stream.probe_count(idx, k, s, e)
        left, right, _ = stream.emit_pairs()
        rows = [golden["reads"][int(l)] + golden["targets"][int(r)] for l, r in zip(left, right)]
        assert sort_rows(rows) == sort_rows(want)


@pytest.mark.parametrize("i", range(len(EDGE_TEXTS)))
def test_edge_texts(stream, i):
    check(stream, EDGE_TEXTS[i], comment="#")


def test_column_choice_header_minus_and_keyless(stream):
    text = b"s,e,name,contig\n10,20,a,chr1\n30,40,b,chr2\r\n50,60,c,chr1\n"
    check(stream, text, delimiter=",", has_header=True, col_key=3, col_start=0, col_end=1, end_minus=1)
    sc, o = check(stream, text, delimiter=",", has_header=True, col_key=None, col_start=0, col_end=1)
    assert sc.dictionary == [] and sc.key_ids_ptr == 0


@pytest.mark.parametrize("text,kind", BAD)
def test_errors(stream, golden, text, kind):
    with pytest.raises(SO.ScanError) as eo:
        SO.scan_delimited(text)
    with pytest.raises(sn.SequilaCudaError) as eg:
        sn.CudaScan.from_text(stream, text)
    assert eg.value.code == (6 if kind == "cast" else 7)  # SQ_ECAST / SQ_EPARSE
    if kind == "cast":  # the reference's text (interval_join.rs:1959-1965) with the first offending value
        assert str(eg.value) == str(eo.value)


def test_first_bad_row_wins_in_a_large_text(stream):
    rng = np.random.default_rng(5)
    text, *_ = random_table(rng, 50_000)
    rows = text.split(b"\n")
    rows[31_000] = b"chr1\t7000000000\t5"
    rows[12_345] = b"chr1\t12\t6000000000"
    rows[40_000] = b"chr1\tzz\t5"
    with pytest.raises(sn.SequilaCudaError) as eg:
        sn.CudaScan.from_text(stream, b"\n".join(rows))
    assert eg.value.code == 6 and str(eg.value) == SO.CAST_ERROR.format(6000000000)


@pytest.mark.parametrize("seed", range(6))
def test_random_tables(stream, seed):
    rng = np.random.default_rng(100 + seed)
    text, *_ = random_table(rng, 120_000, crlf=bool(seed & 1), blanks=bool(seed & 2), extra=bool(seed & 4),
                            long_names=seed >= 3)
    check(stream, text, comment="#")


def test_dictionary_table_growth(cuda_ctx, stream):
    cuda_ctx.set_option("cuda_scan_dict_capacity", 4)  # the table starts with 4 slots: 70 distinct names force two regrowths
    try:
        rng = np.random.default_rng(9)
        text, *_ = random_table(rng, 20_000, long_names=True)
        sc, o = check(stream, text)
        assert len(sc.dictionary) == 70
    finally:
        cuda_ctx.set_option("cuda_scan_dict_capacity", 1 << 16)


def test_device_text_entry_point(cuda_ctx, stream):
    import torch
    rng = np.random.default_rng(11)
    text, *_ = random_table(rng, 30_000)
    t = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    sc = sn.CudaScan.from_text(stream, t)
    o = SO.scan_delimited(text)
    k, s, e, ids = sc.fetch()
    assert np.array_equal(k, o["key_hash"]) and np.array_equal(s, o["start"]) and np.array_equal(e, o["end"])
    assert sc.dictionary == o["dictionary"] and np.array_equal(ids, o["ids"])


def bed_text(table):
    names = np.array([n.encode() for n in sn.synth.CONTIG_NAMES], dtype=object)
    c = names[table["contig"]]
    return b"\n".join(b"%s\t%d\t%d" % (c[i], table["start"][i], table["end"][i]) for i in range(len(c))) + b"\n"


def test_scanned_join_equals_the_oracle_join_at_cfg2_scale(cuda_ctx, stream, oracle):
    """BED text of the cfg2 tables (scaled) -> scan -> build -> probe: the columns equal the generator's, the pair
    multiset equals the coitrees oracle's, the ids column gathers back the contig of every left row."""
    b, p = sn.synth.cfg2(scale=0.3)
    tb, tp = bed_text(b), bed_text(p)
    sb = sn.CudaScan.from_text(stream, tb)
    sp = sn.CudaScan.from_text(stream, tp)
    for sc, t in ((sb, b), (sp, p)):
        k, s, e, ids = sc.fetch()
        assert np.array_equal(s, t["start"]) and np.array_equal(e, t["end"])
        assert [sc.dictionary[i].decode() for i in ids[:1000]] == [sn.synth.CONTIG_NAMES[c] for c in t["contig"][:1000]]
    idx, col = sb.build_index(cuda_ctx, with_ids_column=True)
    n = sp.probe_count(stream, idx)
    ol, orr, oc = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    assert n == len(ol)
    k, s, e = sp.fetch(ids=False)
    stream.probe_count(idx, k, s, e)
    left, right, counts = stream.emit_pairs()
    assert np.array_equal(counts, oc)
    assert np.array_equal(oracle.sorted_pairs(left, right), oracle.sorted_pairs(ol, orr))
    got_ids = stream.gather_build(col, np.uint32)
    bk, bs, be, bids = sb.fetch()
    assert np.array_equal(got_ids, bids[left])


def test_large_host_text_goes_through_the_pinned_ring(stream):
    """texts of 16 MB and more are copied by four host threads through a ring of pinned chunks (sq_scan.cu,
    copy_text_to_device): every chunk must land at its offset — the columns equal the generator's."""
    import io
    import pyarrow as pa
    import pyarrow.csv as pacsv
    b, _ = sn.synth.cfg5(nb=1_500_000, np_=1000)
    names = pa.array(sn.synth.CONTIG_NAMES)
    contig = pa.DictionaryArray.from_arrays(pa.array(b["contig"].astype(np.int32)), names).cast(pa.string())
    buf = io.BytesIO()
    pacsv.write_csv(pa.table({"c": contig, "s": pa.array(b["start"]), "e": pa.array(b["end"])}), buf,
                    pacsv.WriteOptions(include_header=False, delimiter="\t", quoting_style="none"))
    text = buf.getvalue()
    assert len(text) > (33 << 20)  # several rounds of the ring, last chunk partial
    for _ in range(2):  # the second call reuses the ring
        sc = sn.CudaScan.from_text(stream, text)
        k, s, e, ids = sc.fetch()
        assert np.array_equal(s, b["start"]) and np.array_equal(e, b["end"])
        assert np.array_equal(np.array([sn.synth.CONTIG_NAMES.index(d.decode()) for d in sc.dictionary])[ids], b["contig"])
