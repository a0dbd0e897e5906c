"""GPU: the asynchronous tile pipeline (sq_stream_submit / sq_stream_collect) against the oracle — several tiles in
flight per stream, buffers sized from the previous tile's fan-out (first tile and fan-out jumps go through the
re-emit path), every wire flag, both probe layouts, and the `sequila.cuda_*` option surface it is configured by."""
import numpy as np
import pytest

import sequila_native_b200 as sn
from sequila_native_b200 import _native as N
from helpers import canon

pytestmark = pytest.mark.gpu


def tiles_of(p, n_tiles):
    b = np.linspace(0, len(p["key"]), n_tiles + 1).astype(np.int64)
    return [(int(b[i]), int(b[i + 1])) for i in range(n_tiles)]


def run_pipeline(ctx, idx, p, n_tiles, flags=0, depth=None, pinned=True):
    """-> per tile (n_pairs, left, right, counts) in probe order, submitting ahead as far as the stream allows"""
    if depth:
        ctx.set_option("cuda_pipeline_depth", depth)
    st = sn.CudaStream(ctx)
    cols = {k: (ctx.pinned_copy(p[k]) if pinned else np.ascontiguousarray(p[k])) for k in ("key", "start", "end")}
    out, pending = [], []
    depth = int(ctx.get_option("cuda_pipeline_depth"))
    for lo, hi in tiles_of(p, n_tiles):
        if len(pending) == depth:
            out.append(st.collect(pending.pop(0)))
        pending.append(st.submit(idx, cols["key"][lo:hi], cols["start"][lo:hi], cols["end"][lo:hi], flags))
        assert st.in_flight == len(pending)
    while pending:
        out.append(st.collect(pending.pop(0)))
    assert st.in_flight == 0
    return st, out


def check_tiles(oracle, b, p, n_tiles, out, flags):
    oidx = oracle.OracleIndex(b["key"], b["start"], b["end"])
    for (lo, hi), (n, left, right, counts) in zip(tiles_of(p, n_tiles), out):
        ol, orr, oc = oidx.probe(p["key"][lo:hi], p["start"][lo:hi], p["end"][lo:hi])
        assert n == len(ol)
        if flags & N.TILE_COUNT_ONLY:
            assert left is None and right is None
        else:
            assert len(left) == n
            r = right if right is not None else np.repeat(np.arange(hi - lo, dtype=np.uint32), counts)
            assert np.all(np.diff(r.astype(np.int64)) >= 0)
            assert np.array_equal(canon(left, r), canon(ol, orr))
        if not flags & N.TILE_NO_COUNTS:
            assert np.array_equal(counts, oc)
        else:
            assert counts is None
        if flags & (N.TILE_RIGHT_IDX | N.TILE_EXPAND_RIGHT) and not flags & N.TILE_COUNT_ONLY and n:
            assert right is not None and np.array_equal(right, np.repeat(np.arange(hi - lo, dtype=np.uint32), oc))


@pytest.mark.parametrize("layout", ["packed", "soa"])
@pytest.mark.parametrize("flags", [0, N.TILE_RIGHT_IDX, N.TILE_EXPAND_RIGHT, N.TILE_COUNT_ONLY,
                                   N.TILE_COUNT_ONLY | N.TILE_NO_COUNTS, N.TILE_COUNTS_U8, N.TILE_COUNTS_U8 | N.TILE_RIGHT_IDX,
                                   N.TILE_COUNTS_U8 | N.TILE_COUNT_ONLY])
def test_pipeline_matches_oracle(cuda_ctx, oracle, layout, flags):
    cuda_ctx.set_option("cuda_probe_layout", layout)
    try:
        b, p = sn.synth.cfg5(scale=0.004)
        idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
        st, out = run_pipeline(cuda_ctx, idx, p, 9, flags)
        check_tiles(oracle, b, p, 9, out, flags)
        stats = st.pipeline_stats()
        assert stats["tiles"] == 9 and stats["h2d_bytes"] == 16 * len(p["key"])
        if flags & N.TILE_COUNTS_U8:  # cfg5: a handful of hits per row, every tile's counts travel as bytes
            assert all(o[3].dtype == np.uint8 for o in out)
        if not flags & N.TILE_COUNT_ONLY:
            assert stats["regrown"] <= 1  # only the first tile may have been sized blind
    finally:
        cuda_ctx.set_option("cuda_probe_layout", "auto")


def test_pipeline_survives_fanout_jumps_and_empty_tiles(cuda_ctx, oracle):
    """tiles alternate between ~1 and ~100 hits per row, with empty tiles and absent keys in between: every estimate
    is wrong, so both the device re-emit and the pinned re-allocation paths run"""
    rng = np.random.default_rng(3)
    b4, p4 = sn.synth.cfg4(scale=0.02)
    b2, p2 = sn.synth.cfg2(scale=0.02)
    b2 = dict(b2)
    b2["key"] = b2["key"] ^ np.uint64(0x5555)  # distinct key space: one index holds both shapes
    p2 = dict(p2)
    p2["key"] = p2["key"] ^ np.uint64(0x5555)
    b = {k: np.concatenate([b4[k], b2[k]]) for k in ("key", "start", "end")}
    idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
    oidx = oracle.OracleIndex(b["key"], b["start"], b["end"])
    st = sn.CudaStream(cuda_ctx)
    pieces = []
    for i in range(8):
        src = p4 if i % 2 else p2
        lo = int(rng.integers(0, len(src["key"]) - 3000))
        n = 0 if i == 4 else int(rng.integers(500, 3000))
        t = {k: cuda_ctx.pinned_copy(src[k][lo:lo + n]) for k in ("key", "start", "end")}
        if i == 6:
            t["key"][:] = np.uint64(12345)  # key hash absent from the build side: zero pairs (interval_join.rs:965)
        pieces.append(t)
    tickets = []
    got = []
    for t in pieces:
        if len(tickets) == 3:
            got.append(st.collect(tickets.pop(0)))
        tickets.append(st.submit(idx, t["key"], t["start"], t["end"]))
    with pytest.raises(sn.SequilaCudaError) as e:  # a fourth tile does not fit the default depth of 3
        st.submit(idx, pieces[0]["key"], pieces[0]["start"], pieces[0]["end"])
    assert e.value.code == N.SQ_EBUSY
    with pytest.raises(sn.SequilaCudaError) as e:  # flag bits the library does not know
        sn.CudaStream(cuda_ctx).submit(idx, pieces[0]["key"], pieces[0]["start"], pieces[0]["end"], flags=64)
    assert e.value.code == N.SQ_EINVAL
    with pytest.raises(sn.SequilaCudaError) as e:  # tiles are collected in submission order
        st.collect(tickets[1])
    assert e.value.code == N.SQ_ESTATE
    while tickets:
        got.append(st.collect(tickets.pop(0)))
    for t, (n, left, right, counts) in zip(pieces, got):
        ol, orr, oc = oidx.probe(t["key"], t["start"], t["end"])
        assert n == len(ol) and np.array_equal(counts if counts is not None else np.zeros(0, np.uint32), oc)
        if n:
            assert np.array_equal(canon(left, np.repeat(np.arange(len(oc), dtype=np.uint32), counts)), canon(ol, orr))
    assert got[6][0] == 0 and got[4][0] == 0
    assert st.pipeline_stats()["regrown"] >= 2


def test_byte_counts_fall_back_tile_by_tile(cuda_ctx, oracle):
    """SQ_TILE_COUNTS_U8: tiles whose counts all fit a byte hand over uint8 counts, a tile with a row of more than 255 hits
    hands over the usual 4-byte counts (sq_tile_out.counts_width); the pairs are the oracle's either way"""
    rng = np.random.default_rng(9)
    n = 40_000
    start = np.sort(rng.integers(0, 4_000_000, n)).astype(np.int32)
    b = {"key": np.full(n, 77, np.uint64), "start": start, "end": (start + rng.integers(10, 60, n)).astype(np.int32)}
    b["start"][:400] = 1_000_000 + np.arange(400, dtype=np.int32)  # 400 rows piled over one spot
    b["end"][:400] = 1_000_600
    idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
    q = rng.integers(0, 4_000_000, 30_000)
    p = {"key": np.full(30_000, 77, np.uint64), "start": q.astype(np.int32), "end": (q + 100).astype(np.int32)}
    p["start"][12_345] = 1_000_100  # one probe row under the pile: > 255 hits, in the second of three tiles
    p["end"][12_345] = 1_000_500
    p["start"][10_000:10_050] = 3_999_000  # keep the rest of that tile's rows ordinary
    p["end"][10_000:10_050] = 3_999_050
    far = (p["end"] < 999_000) | (p["start"] > 1_001_000)
    far[12_345] = True
    p = {k: v[far] for k, v in p.items()}
    at = int(np.flatnonzero((p["start"] == 1_000_100) & (p["end"] == 1_000_500))[0])
    for flags in (N.TILE_COUNTS_U8, N.TILE_COUNTS_U8 | N.TILE_COUNT_ONLY):
        st, out = run_pipeline(cuda_ctx, idx, p, 3, flags)
        check_tiles(oracle, b, p, 3, out, flags)
        widths = [o[3].dtype.itemsize for o in out]
        hot = [lo <= at < hi for lo, hi in tiles_of(p, 3)]
        assert widths == [4 if h else 1 for h in hot] and sum(hot) == 1
        assert int(out[hot.index(True)][3].max()) >= 400
    with pytest.raises(sn.SequilaCudaError) as e:  # the host-side expansion reads 4-byte counts
        sn.CudaStream(cuda_ctx).submit(idx, cuda_ctx.pinned_copy(p["key"]), cuda_ctx.pinned_copy(p["start"]), cuda_ctx.pinned_copy(p["end"]),
                                       flags=N.TILE_COUNTS_U8 | N.TILE_EXPAND_RIGHT)
    assert e.value.code == N.SQ_EINVAL


def test_pipeline_many_streams_concurrently(cuda_ctx, oracle):
    """partitions = OS threads, one sq_stream each, each with tiles in flight over ONE shared index"""
    import concurrent.futures as cf
    b, p = sn.synth.cfg5(scale=0.004)
    idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
    T = 6
    parts = tiles_of(p, T)

    def work(lohi):
        lo, hi = lohi
        sub = {k: p[k][lo:hi] for k in ("key", "start", "end")}
        _, out = run_pipeline(cuda_ctx, idx, sub, 5)
        return sub, out
    with cf.ThreadPoolExecutor(T) as pool:
        res = list(pool.map(work, parts))
    for sub, out in res:
        check_tiles(oracle, b, sub, 5, out, 0)


def test_pageable_inputs_and_depth_option(cuda_ctx, oracle):
    b, p = sn.synth.cfg3(scale=0.005)
    idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
    try:
        _, out = run_pipeline(cuda_ctx, idx, p, 7, 0, depth=2, pinned=False)
        check_tiles(oracle, b, p, 7, out, 0)
        _, out = run_pipeline(cuda_ctx, idx, p, 7, 0, depth=8)
        check_tiles(oracle, b, p, 7, out, 0)
    finally:
        cuda_ctx.set_option("cuda_pipeline_depth", 3)


def test_options_surface(cuda_ctx):
    """the `sequila.cuda_*` keys: symbolic and numeric values, the optional prefix, rejections with a message"""
    c = cuda_ctx
    assert c.get_option("sequila.cuda_probe_layout") == "auto"
    c.set_option("sequila.cuda_probe_layout", "SOA")
    assert c.get_option("cuda_probe_layout") == "soa"
    c.set_option("cuda_probe_layout", "auto")
    c.set_option("cuda_probe_block", 256)
    assert c.get_option("cuda_probe_block") == "256"
    c.set_option("cuda_probe_block", 128)
    for key, bad in (("cuda_probe_block", "100"), ("cuda_probe_layout", "tree"), ("cuda_pipeline_depth", "1"),
                     ("cuda_scan_dict_capacity", "6"), ("cuda_nonsense", "1"), ("cuda_rows_per_bin", "x")):
        with pytest.raises(sn.SequilaCudaError) as e:
            c.set_option(key, bad)
        assert e.value.code == N.SQ_EINVAL and key.split(".")[-1] in str(e.value)
    c.set_option("cuda_l2_persist_mb", 0)
    assert c.get_option("cuda_l2_persist_mb") == "0"
    # every key the session layer accepts (session.CUDA_KEYS, what `SET sequila.cuda_* TO ...` stores) is a key the library
    # knows, documented in the header, and can be set back to the value it reports
    from sequila_native_b200.session import CUDA_KEYS
    import os
    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "sequila_cuda.h")).read()
    for key in CUDA_KEYS:
        v = c.get_option("sequila." + key)
        c.set_option(key, v)
        assert c.get_option(key) == v
        assert key in header, key


@pytest.mark.parametrize("parts,tiles", [(1, 1), (3, 10), (8, 64)])
def test_native_partition_loop(cuda_ctx, oracle, parts, tiles):
    """sq_driver_run (include/sequila_driver.h): native partition threads over submit / collect — pair total and
    the xor of every left_idx that reached the host equal the oracle's, run after run on the same driver"""
    from sequila_native_b200.cuda_join import CudaDriver
    b, p = sn.synth.cfg5(scale=0.004)
    idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
    ol, _, _ = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    cols = {k: cuda_ctx.pinned_copy(p[k]) for k in ("key", "start", "end")}
    drv = CudaDriver(cuda_ctx, parts)
    for flags in (0, N.TILE_COUNT_ONLY | N.TILE_NO_COUNTS, 0):
        r = drv.run(idx, cols["key"], cols["start"], cols["end"], tiles, flags, checksum=True)
        assert r["n_pairs"] == len(ol) and r["n_tiles"] == tiles
        assert r["h2d_bytes"] == 16 * len(p["key"])
        if flags == 0:
            assert r["left_xor"] == int(np.bitwise_xor.reduce(ol.astype(np.uint64)))


@pytest.mark.parametrize("layout", ["packed", "soa"])
def test_key_dictionary_ids_on_the_wire(cuda_ctx, oracle, layout):
    """sq_stream_submit_ids: the key column as 4-byte dictionary ids (12 bytes per probe row) gives the rows the u64 hashes
    give; ids outside the dictionary (SQ_NULL_INDEX = NULL key) match nothing; the native driver's ids pass agrees"""
    from sequila_native_b200.cuda_join import CudaDriver
    cuda_ctx.set_option("cuda_probe_layout", layout)
    try:
        b, p = sn.synth.cfg5(scale=0.004)
        idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
        dict_hashes = sn.synth.key_hash(np.arange(24))
        ids = p["contig"].astype(np.uint32).copy()
        rng = np.random.default_rng(2)
        nulls = rng.random(len(ids)) < 0.03
        ids[nulls] = np.where(rng.random(int(nulls.sum())) < 0.5, N.NULL_INDEX, 24 + 5).astype(np.uint32)
        keys = p["key"].copy()
        keys[nulls] = np.uint64(0xDEAD)  # a hash the build side never saw: the oracle's picture of a NULL key
        cols = {"ids": cuda_ctx.pinned_copy(ids), "start": cuda_ctx.pinned_copy(p["start"]), "end": cuda_ctx.pinned_copy(p["end"])}
        st = sn.CudaStream(cuda_ctx)
        with pytest.raises(sn.SequilaCudaError) as e:
            st.submit_ids(idx, cols["ids"][:10], cols["start"][:10], cols["end"][:10])
        assert e.value.code == N.SQ_ESTATE
        st.set_key_dictionary(dict_hashes)
        cuts = tiles_of(p, 5)
        tickets = [st.submit_ids(idx, cols["ids"][a:z], cols["start"][a:z], cols["end"][a:z]) for a, z in cuts[:3]]
        out = [st.collect(t) for t in tickets]
        tickets = [st.submit_ids(idx, cols["ids"][a:z], cols["start"][a:z], cols["end"][a:z]) for a, z in cuts[3:]]
        out += [st.collect(t) for t in tickets]
        assert st.pipeline_stats()["h2d_bytes"] == 12 * len(ids)
        q = dict(p)
        q["key"] = keys
        check_tiles(oracle, b, q, 5, out, 0)
        ol, _, _ = oracle.join(b["key"], b["start"], b["end"], keys, p["start"], p["end"])
        drv = CudaDriver(cuda_ctx, 3)
        r = drv.run_ids(idx, dict_hashes, cols["ids"], cols["start"], cols["end"], 7, 0, checksum=True)
        assert r["n_pairs"] == len(ol) and r["h2d_bytes"] == 12 * len(ids)
        assert r["left_xor"] == int(np.bitwise_xor.reduce(ol.astype(np.uint64)))
    finally:
        cuda_ctx.set_option("cuda_probe_layout", "auto")
