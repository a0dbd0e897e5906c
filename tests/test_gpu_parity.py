"""GPU: parity of the CUDA path (through the C ABI) with the oracle — bit-exact multiset of
(left_row, right_row) pairs, right_idx non-decreasing, per-row counts equal."""
import numpy as np
import pytest

import sequila_native_b200 as sn
from helpers import canon, encode_tables, rows_from_pairs, sort_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["packed", "packed2", "soa", "staged", "soa_walk"])
def probe_layout(request, cuda_ctx):
    """Every parity test runs against both probe implementations: the fused kernel over packed
    lines (narrow indexes) and the count/scan/write kernels over the SoA arrays (option
    sequila.cuda_probe_layout, sq_probe_packed.cu::use_packed)."""
    # "staged" = packed lines served by the shared-memory (TMA) staged kernel whatever the probe order: unsorted tiles
    # then exercise its global-memory path, sorted ones the staged path
    # "packed2" = the packed-line kernel with two tiles per CTA (option cuda_probe_tiles)
    # "soa" = the rank-difference kernel where the build side allows it; "soa_walk" = count / scan / write always
    cuda_ctx.set_option("sequila.cuda_probe_layout", {"staged": "packed", "packed2": "packed", "soa_walk": "soa"}.get(request.param, request.param))
    cuda_ctx.set_option("sequila.cuda_rank_count", "off" if request.param == "soa_walk" else "force")
    cuda_ctx.set_option("sequila.cuda_staged_probe", "on" if request.param == "staged" else "off")
    cuda_ctx.set_option("sequila.cuda_probe_tiles", 2 if request.param == "packed2" else 1)
    yield request.param
    cuda_ctx.set_option("sequila.cuda_probe_tiles", 1)
    cuda_ctx.set_option("sequila.cuda_probe_layout", "auto")
    cuda_ctx.set_option("sequila.cuda_staged_probe", "auto")
    cuda_ctx.set_option("sequila.cuda_rank_count", "on")


def cuda_join(ctx, L, R):
    idx = sn.CudaIndex.build(ctx, L["key"], L["start"], L["end"])
    st = sn.CudaStream(ctx)
    n = st.probe_count(idx, R["key"], R["start"], R["end"])
    l, r, c = st.emit_pairs()
    assert len(l) == n == len(r) and int(c.sum()) == n
    return l, r, c


def assert_same(oracle, ctx, b, p):
    ol, orr, oc = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    l, r, c = cuda_join(ctx, b, p)
    assert len(l) == len(ol)
    assert np.all(np.diff(r.astype(np.int64)) >= 0), "right_idx must be non-decreasing (probe order kept)"
    assert np.array_equal(c, oc)
    assert np.array_equal(canon(l, r), canon(ol, orr))


def test_cfg1_fixture_equi_16_rows(cuda_ctx, golden):
    L, R = encode_tables(golden["reads"], golden["targets"], equi=True)
    l, r, _ = cuda_join(cuda_ctx, L, R)
    assert rows_from_pairs(L, R, l, r) == sort_rows(golden["equi_rows"])


def test_cfg1_fixture_range_only_32_rows(cuda_ctx, golden):
    L, R = encode_tables(golden["reads"], golden["targets"], equi=False)
    l, r, _ = cuda_join(cuda_ctx, L, R)
    assert rows_from_pairs(L, R, l, r) == sort_rows(golden["range_rows"])


def test_closed_and_strict_boundaries(cuda_ctx, golden):
    L, R = encode_tables(golden["boundary_a"], golden["boundary_b"])
    l, r, _ = cuda_join(cuda_ctx, L, R)
    assert rows_from_pairs(L, R, l, r) == sort_rows(golden["closed_rows"])
    Ls, Rs = dict(L, end=L["end"] - 1), dict(R, end=R["end"] - 1)
    l, r, _ = cuda_join(cuda_ctx, Ls, Rs)
    assert rows_from_pairs(L, R, l, r) == sort_rows(golden["strict_rows"])


@pytest.mark.parametrize("name,scale", [("cfg2", 0.1), ("cfg3", 0.02), ("cfg4", 0.05), ("cfg5", 0.002)])
def test_synthetic_configs_match_oracle(cuda_ctx, oracle, name, scale):
    b, p = sn.synth.CONFIGS[name](scale=scale)
    assert_same(oracle, cuda_ctx, b, p)


def test_random_differential_including_inverted_and_absent_keys(cuda_ctx, oracle):
    rng = np.random.default_rng(42)
    for it in range(40):
        nb, npq, nk = int(rng.integers(0, 5000)), int(rng.integers(0, 3000)), int(rng.integers(1, 40))
        span, wmax = int(rng.integers(10, 20000)), int(rng.integers(1, 600))
        b = {"key": rng.integers(0, nk, nb).astype(np.uint64) * 7919 + 3,
             "start": rng.integers(-50, span, nb).astype(np.int32)}
        b["end"] = (b["start"] + rng.integers(0, wmax, nb)).astype(np.int32)
        p = {"key": rng.integers(0, nk + 2, npq).astype(np.uint64) * 7919 + 3,
             "start": rng.integers(-50, span, npq).astype(np.int32)}
        p["end"] = (p["start"] + rng.integers(0, wmax, npq)).astype(np.int32)
        if it % 3 == 0:  # inverted intervals (strict compare on point intervals produces them)
            m = rng.random(nb) < 0.1
            b["end"][m] = b["start"][m] - 5
            m = rng.random(npq) < 0.1
            p["end"][m] = p["start"][m] - 5
        assert_same(oracle, cuda_ctx, b, p)


def test_edge_cases(cuda_ctx, oracle):
    e = np.array([], dtype=np.int32)
    k = np.array([], dtype=np.uint64)
    one = {"key": np.array([7], dtype=np.uint64), "start": np.array([5], dtype=np.int32),
           "end": np.array([10], dtype=np.int32)}
    empty = {"key": k, "start": e, "end": e}
    assert_same(oracle, cuda_ctx, empty, one)       # empty build side
    assert_same(oracle, cuda_ctx, one, empty)       # empty probe batch
    assert_same(oracle, cuda_ctx, empty, empty)
    assert_same(oracle, cuda_ctx, one, one)
    # key hash equal to the hash table's internal sentinel, and key 0
    s = {"key": np.array([2 ** 64 - 1, 0, 2 ** 64 - 1], dtype=np.uint64),
         "start": np.array([1, 1, 3], dtype=np.int32), "end": np.array([2, 2, 9], dtype=np.int32)}
    assert_same(oracle, cuda_ctx, s, s)
    # extreme coordinates inside the parity domain (i32::MIN < v < i32::MAX)
    x = {"key": np.zeros(4, dtype=np.uint64),
         "start": np.array([-2 ** 31 + 1, -5, 0, 2 ** 31 - 3], dtype=np.int32),
         "end": np.array([-2 ** 31 + 2, 2 ** 31 - 2, 0, 2 ** 31 - 2], dtype=np.int32)}
    assert_same(oracle, cuda_ctx, x, x)
    # one giant interval early in the segment (worst case for a flat scan) + many duplicates
    n = 3000
    g = {"key": np.zeros(n, dtype=np.uint64), "start": np.arange(n, dtype=np.int32) * 10}
    g["end"] = g["start"] + 3
    g["end"][0] = 10 ** 8
    g["start"][100:200] = 555
    g["end"][100:200] = 560
    assert_same(oracle, cuda_ctx, g, g)


def test_many_keys_grows_the_key_table(cuda_ctx, oracle):
    rng = np.random.default_rng(3)
    nb = 20000
    b = {"key": rng.integers(0, 2 ** 63, nb).astype(np.uint64), "start": rng.integers(0, 1000, nb).astype(np.int32)}
    b["end"] = b["start"] + 10
    p = {"key": np.concatenate([b["key"][:5000], rng.integers(0, 2 ** 63, 500).astype(np.uint64)]),
         "start": np.zeros(5500, dtype=np.int32), "end": np.full(5500, 2000, dtype=np.int32)}
    assert_same(oracle, cuda_ctx, b, p)


def test_probe_tiles_reuse_one_index_and_stream(cuda_ctx, oracle):
    """one build, many probe batches on one stream (interval_join.rs:1134-1167 state machine)"""
    b, p = sn.synth.cfg2(scale=0.05)
    idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
    oidx = oracle.OracleIndex(b["key"], b["start"], b["end"])
    st = sn.CudaStream(cuda_ctx)
    for lo in range(0, len(p["key"]), 8192):
        sl = slice(lo, lo + 8192)
        l, r, c = st.probe(idx, p["key"][sl], p["start"][sl], p["end"][sl])
        ol, orr, oc = oidx.probe(p["key"][sl], p["start"][sl], p["end"][sl])
        assert np.array_equal(c, oc) and np.array_equal(canon(l, r), canon(ol, orr))


def test_fused_join_call_and_capacity_protocol(cuda_ctx, oracle):
    """sq_probe_join: one fused pass when the pairs fit; SQ_ECAPACITY + exact n_pairs otherwise,
    after which sq_probe_emit_pairs completes the same tile."""
    b, p = sn.synth.cfg4(scale=0.02)
    ol, orr, oc = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
    st = sn.CudaStream(cuda_ctx)
    big = (np.empty(len(ol) + 7, np.uint32), np.empty(len(ol) + 7, np.uint32), np.empty(len(p["key"]), np.uint32))
    for _ in range(2):  # second call reuses the grown speculative buffers: single pass
        n = st.probe_join(idx, p["key"], p["start"], p["end"], big)
        assert n == len(ol) and np.array_equal(big[2], oc)
        assert np.array_equal(canon(big[0][:n], big[1][:n]), canon(ol, orr))
    small = (np.empty(10, np.uint32), np.empty(10, np.uint32), None)
    with pytest.raises(sn.SequilaCudaError) as e:
        st.probe_join(idx, p["key"], p["start"], p["end"], small)
    assert e.value.code == 5 and st.n_pairs == len(ol)
    l, r, c = st.emit_pairs()
    assert np.array_equal(canon(l, r), canon(ol, orr)) and np.array_equal(c, oc)


def test_device_entry_points(cuda_ctx, oracle):
    import torch
    b, p = sn.synth.cfg3(scale=0.01)
    ol, orr, oc = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(dev)
    idx = sn.CudaIndex.build_device(cuda_ctx, t(b["key"]), t(b["start"]), t(b["end"]),
                                    torch.cuda.current_stream().cuda_stream)
    st = sn.CudaStream(cuda_ctx, cuda_stream=torch.cuda.current_stream().cuda_stream)
    pk, ps, pe = t(p["key"]), t(p["start"]), t(p["end"])
    n = st.probe_count_device(idx, pk, ps, pe)
    assert n == len(ol)
    left = torch.empty(n, dtype=torch.int32, device=dev)
    right = torch.empty(n, dtype=torch.int32, device=dev)
    st.emit_pairs_device(left, right)
    torch.cuda.synchronize()
    assert np.array_equal(canon(left.cpu().numpy().view(np.uint32), right.cpu().numpy().view(np.uint32)), canon(ol, orr))
    left.zero_(); right.zero_()
    assert st.probe_join_device(idx, pk, ps, pe, left, right) == n
    torch.cuda.synchronize()
    assert np.array_equal(canon(left.cpu().numpy().view(np.uint32), right.cpu().numpy().view(np.uint32)), canon(ol, orr))
    assert st.digest_device(left, right, n) == oracle.pair_digest(ol, orr)
    with pytest.raises(sn.SequilaCudaError) as e:
        st.probe_join_device(idx, pk, ps, pe, left[: n // 2], right[: n // 2])
    assert e.value.code == 5 and st.n_pairs == n


def test_gather_columns_match_take(cuda_ctx, oracle):
    """materialise = arrow take of every column (interval_join.rs:1620-1632)"""
    b, p = sn.synth.cfg3(scale=0.01)
    idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
    cols = [idx.add_column(b[c]) for c in ("contig", "start", "end")]
    wide = idx.add_column(b["start"].astype(np.int64) * 3)
    st = sn.CudaStream(cuda_ctx)
    l, r, _ = st.probe(idx, p["key"], p["start"], p["end"])
    for cid, name in zip(cols, ("contig", "start", "end")):
        assert np.array_equal(st.gather_build(cid, np.int32), oracle.gather_i32(b[name], l))
    assert np.array_equal(st.gather_build(wide, np.int64), (b["start"].astype(np.int64) * 3)[l])
    for name in ("contig", "start", "end"):
        assert np.array_equal(st.gather_probe(p[name]), oracle.gather_i32(p[name], r))
    views = np.arange(len(p["key"]) * 2, dtype=np.uint64).reshape(-1, 2)  # 16-byte values (Utf8View views)
    got = st.gather_probe(views.view(np.dtype([("a", np.uint64), ("b", np.uint64)])).reshape(-1))
    assert np.array_equal(got.view(np.uint64).reshape(-1, 2), views[r])


def test_cast_i64_to_i32_and_overflow_message(cuda_ctx, golden):
    st = sn.CudaStream(cuda_ctx)
    v = np.array([1, 2, 2 ** 31 - 1, -2 ** 31], dtype=np.int64)
    assert np.array_equal(st.cast_i64_to_i32(v), v.astype(np.int32))
    assert np.array_equal(st.cast_i64_to_i32(v[:3], minus=1), (v[:3] - 1).astype(np.int32))
    bad = np.array([1, 3, golden["cast_error_value"], 2 ** 40], dtype=np.int64)
    with pytest.raises(sn.SequilaCudaError) as e:
        st.cast_i64_to_i32(bad)
    assert str(e.value) == golden["cast_error_format"].format(golden["cast_error_value"])
    assert e.value.code == 6


def test_error_paths(cuda_ctx):
    st = sn.CudaStream(cuda_ctx)
    with pytest.raises(sn.SequilaCudaError) as e:
        st.emit_pairs()
    assert e.value.code == 4  # SQ_ESTATE: emit without count
    b, p = sn.synth.cfg2(scale=0.01)
    idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
    n = st.probe_count(idx, p["key"], p["start"], p["end"])
    assert n > 1
    small = (np.empty(1, np.uint32), np.empty(1, np.uint32), np.empty(len(p["key"]), np.uint32))
    with pytest.raises(sn.SequilaCudaError) as e:
        st.emit_pairs(out=small)
    assert e.value.code == 5  # SQ_ECAPACITY


@pytest.mark.parametrize("wire", ["rle", "copy"])
def test_right_idx_wire_encodings(cuda_ctx, oracle, wire):
    """right_idx crosses PCIe as per-row counts and is expanded on the host (default), or as itself."""
    cuda_ctx.set_option("cuda_right_idx_wire", wire)
    try:
        _wire_cases(cuda_ctx, oracle)
    finally:
        cuda_ctx.set_option("cuda_right_idx_wire", "rle")


def _wire_cases(cuda_ctx, oracle):
    for name, scale in (("cfg2", 0.05), ("cfg3", 0.01), ("cfg4", 0.02)):
        b, p = sn.synth.CONFIGS[name](scale=scale)
        assert_same(oracle, cuda_ctx, b, p)
    one = {"key": np.array([7], dtype=np.uint64), "start": np.array([5], dtype=np.int32), "end": np.array([10], dtype=np.int32)}
    assert_same(oracle, cuda_ctx, one, one)


def test_concurrent_partitions_share_one_index(cuda_ctx, oracle):
    """Many OS threads, one sq_stream each, probe ONE index at the same time (DataFusion partitions over a
    CollectLeft build side, interval_join.rs:473-487, 528-556): every partition's pairs equal the oracle's."""
    import concurrent.futures as cf
    b, p = sn.synth.cfg5(scale=0.004)
    idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
    oidx = oracle.OracleIndex(b["key"], b["start"], b["end"])
    n_parts, n_tiles = 8, 6
    bounds = np.linspace(0, len(p["key"]), n_parts * n_tiles + 1).astype(int)

    def partition(w):
        st = sn.CudaStream(cuda_ctx)
        out = []
        for t in range(w, n_parts * n_tiles, n_parts):
            lo, hi = bounds[t], bounds[t + 1]
            k, s, e = p["key"][lo:hi], p["start"][lo:hi], p["end"][lo:hi]
            n = st.probe_count(idx, k, s, e)
            l, r, c = st.emit_pairs()
            out.append((t, n, l.copy(), r.copy(), c.copy()))
        return out

    with cf.ThreadPoolExecutor(n_parts) as pool:
        results = [x for part in pool.map(partition, range(n_parts)) for x in part]
    assert len(results) == n_parts * n_tiles
    for t, n, l, r, c in results:
        lo, hi = bounds[t], bounds[t + 1]
        ol, orr, oc = oidx.probe(p["key"][lo:hi], p["start"][lo:hi], p["end"][lo:hi])
        assert n == len(ol) and np.array_equal(c, oc) and np.array_equal(canon(l, r), canon(ol, orr))
