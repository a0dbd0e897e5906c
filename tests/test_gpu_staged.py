"""GPU: the shared-memory (TMA) staged probe kernel (sq_probe_staged.cu) against the oracle — position-sorted probe
tiles (staged path), shuffled ones (global path), overlap chains longer than the staged halo (continuation in global
memory), tiles that straddle contigs, probes far sparser than the index (range too wide to stage), row counts around
the CTA size, count-only / no-right variants, and the adaptive kernel choice of a stream."""
import numpy as np
import pytest
import torch

import sequila_native_b200 as sn
from sequila_native_b200 import _native as N
from helpers import canon

pytestmark = pytest.mark.gpu


@pytest.fixture()
def staged(cuda_ctx):
    cuda_ctx.set_option("cuda_probe_layout", "packed")
    cuda_ctx.set_option("cuda_staged_probe", "on")
    yield cuda_ctx
    cuda_ctx.set_option("cuda_probe_layout", "auto")
    cuda_ctx.set_option("cuda_staged_probe", "auto")


def sort_side(p):
    o = np.lexsort((p["start"], p["contig"]))
    return {k: np.ascontiguousarray(v[o]) for k, v in p.items()}


def check(oracle, ctx, b, p):
    idx = sn.CudaIndex.build(ctx, b["key"], b["start"], b["end"])
    assert idx.uses_packed
    st = sn.CudaStream(ctx)
    n = st.probe_count(idx, p["key"], p["start"], p["end"])
    l, r, c = st.emit_pairs()
    ol, orr, oc = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    assert n == len(ol) and np.array_equal(c, oc)
    assert np.all(np.diff(r.astype(np.int64)) >= 0)
    assert np.array_equal(canon(l, r), canon(ol, orr))
    return idx


@pytest.mark.parametrize("scale", [0.0005, 0.004])
@pytest.mark.parametrize("order", ["sorted", "shuffled", "blocks"])
def test_cfg5_shapes(staged, oracle, scale, order):
    b, p = sn.synth.cfg5(scale=scale)
    if order == "sorted":
        p = sort_side(p)
    elif order == "blocks":  # sorted runs of 1000 rows in shuffled order: most CTAs stage, the ones at run borders do not
        q = sort_side(p)
        n = len(q["key"])
        blocks = np.random.default_rng(1).permutation((n + 999) // 1000)
        o = np.concatenate([np.arange(k * 1000, min((k + 1) * 1000, n)) for k in blocks])
        p = {k: np.ascontiguousarray(v[o]) for k, v in q.items()}
    check(oracle, staged, b, p)


@pytest.mark.parametrize("n_probe", [1, 31, 127, 128, 129, 257, 1000])
def test_row_counts_around_the_cta_size(staged, oracle, n_probe):
    b, p = sn.synth.cfg5(scale=0.001)
    p = sort_side(p)
    lo = 5000
    check(oracle, staged, b, {k: v[lo:lo + n_probe] for k, v in p.items()})


def test_overlap_chains_longer_than_the_halo(staged, oracle):
    """every 200th build row is 40,000 bp wide: probes behind it walk back through hundreds of lines — past the two halo
    lines of the staged range, so the scan continues in global memory; still narrow (< 65536) so the index stays packed"""
    rng = np.random.default_rng(5)
    nb = 200_000
    start = np.sort(rng.integers(0, 5_000_000, nb)).astype(np.int32)
    width = rng.integers(50, 150, nb)
    width[::200] = 40_000
    b = {"contig": np.zeros(nb, np.int32), "key": sn.synth.key_hash(np.zeros(nb, np.int32)), "start": start,
         "end": (start + width - 1).astype(np.int32)}
    ps = np.sort(rng.integers(0, 5_000_000, 60_000)).astype(np.int32)
    p = {"contig": np.zeros(len(ps), np.int32), "key": sn.synth.key_hash(np.zeros(len(ps), np.int32)), "start": ps,
         "end": (ps + rng.integers(1, 300, len(ps))).astype(np.int32)}
    check(oracle, staged, b, p)


def test_sparse_probes_over_a_dense_index_and_many_contigs(staged, oracle):
    """1 probe per ~400 build rows: a CTA's 128 rows span far more lines than fit the staging buffer (global path);
    and 600 tiny contigs so that every CTA straddles several key segments"""
    b, p = sn.synth.cfg5(scale=0.004)
    p = sort_side({k: v[::400] for k, v in p.items()})
    check(oracle, staged, b, p)
    rng = np.random.default_rng(9)
    nb = 120_000
    contig = rng.integers(0, 600, nb).astype(np.int32)
    start = rng.integers(0, 20_000, nb).astype(np.int32)
    b = {"contig": contig, "key": sn.synth.key_hash(contig), "start": start, "end": (start + rng.integers(0, 120, nb)).astype(np.int32)}
    pc = rng.integers(0, 640, 50_000).astype(np.int32)  # 40 contigs absent from the build side
    ps = rng.integers(-100, 20_100, 50_000).astype(np.int32)
    p = sort_side({"contig": pc, "key": sn.synth.key_hash(pc), "start": ps, "end": (ps + rng.integers(0, 200, 50_000)).astype(np.int32)})
    check(oracle, staged, b, p)


def test_inverted_probe_rows_and_extreme_coordinates(staged, oracle):
    rng = np.random.default_rng(11)
    nb = 50_000
    start = np.sort(rng.integers(-2_000_000_000, 2_000_000_000, nb)).astype(np.int64)
    b = {"contig": np.zeros(nb, np.int32), "key": sn.synth.key_hash(np.zeros(nb, np.int32)), "start": start.astype(np.int32),
         "end": np.minimum(start + rng.integers(0, 60_000, nb), 2_147_483_646).astype(np.int32)}
    ps = np.sort(rng.integers(-2_100_000_000, 2_100_000_000, 20_000)).astype(np.int64)
    pe = ps + rng.integers(-50_000, 200_000_000, 20_000)  # some rows inverted (end < start), some very wide
    p = {"contig": np.zeros(len(ps), np.int32), "key": sn.synth.key_hash(np.zeros(len(ps), np.int32)),
         "start": ps.astype(np.int32), "end": np.clip(pe, -2_147_483_647, 2_147_483_646).astype(np.int32)}
    check(oracle, staged, b, p)


def test_device_variants_and_overflow(staged, oracle):
    """count-only launch, join without right_idx, and a join into buffers that are too small (reported, not written)"""
    b, p = sn.synth.cfg5(scale=0.004)
    p = sort_side(p)
    dev = torch.device("cuda", 0)
    bd = {k: torch.from_numpy(b[k].view(np.int64) if k == "key" else b[k]).to(dev) for k in ("key", "start", "end")}
    pd = {k: torch.from_numpy(p[k].view(np.int64) if k == "key" else p[k]).to(dev) for k in ("key", "start", "end")}
    ts = torch.cuda.current_stream().cuda_stream
    idx = sn.CudaIndex.build_device(staged, bd["key"], bd["start"], bd["end"], ts)
    st = sn.CudaStream(staged, cuda_stream=ts)
    ol, orr, oc = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    assert st.probe_count_device(idx, pd["key"], pd["start"], pd["end"]) == len(ol)
    assert np.array_equal(st.counts(), oc)
    left = torch.empty(len(ol), dtype=torch.int32, device=dev)
    assert st.probe_join_device(idx, pd["key"], pd["start"], pd["end"], left, None) == len(ol)
    right = np.repeat(np.arange(len(oc), dtype=np.uint32), oc)
    assert np.array_equal(canon(left.cpu().numpy().view(np.uint32), right), canon(ol, orr))
    small = torch.empty(len(ol) // 2, dtype=torch.int32, device=dev)
    with pytest.raises(sn.SequilaCudaError) as e:
        st.probe_join_device(idx, pd["key"], pd["start"], pd["end"], small, None)
    assert e.value.code == N.SQ_ECAPACITY and st.n_pairs == len(ol)


def test_adaptive_kernel_choice(cuda_ctx, oracle):
    """option auto: host tiles are judged by the order of their rows, device tiles by the staged kernel's own report"""
    cuda_ctx.set_option("cuda_probe_layout", "packed")
    try:
        b, p = sn.synth.cfg5(scale=0.004)
        idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
        ol, orr, oc = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
        q = sort_side(p)
        sl, sr, sc = oracle.join(b["key"], b["start"], b["end"], q["key"], q["start"], q["end"])
        for side, (wl, wr, wc) in ((p, (ol, orr, oc)), (q, (sl, sr, sc))):
            st = sn.CudaStream(cuda_ctx)
            for _ in range(4):
                l, r, c = st.probe(idx, side["key"], side["start"], side["end"])
                assert np.array_equal(c, wc) and np.array_equal(canon(l, r), canon(wl, wr))
        dev = torch.device("cuda", 0)
        for side, (wl, wr, wc) in ((p, (ol, orr, oc)), (q, (sl, sr, sc))):
            d = {k: torch.from_numpy(side[k].view(np.int64) if k == "key" else side[k]).to(dev) for k in ("key", "start", "end")}
            st = sn.CudaStream(cuda_ctx, cuda_stream=torch.cuda.current_stream().cuda_stream)
            for _ in range(6):  # shuffled rows: the first launch reports unstageable CTAs, the next ones go to k_probe_packed
                left = torch.empty(len(wl), dtype=torch.int32, device=dev)
                right = torch.empty(len(wl), dtype=torch.int32, device=dev)
                assert st.probe_join_device(idx, d["key"], d["start"], d["end"], left, right) == len(wl)
                assert np.array_equal(canon(left.cpu().numpy().view(np.uint32), right.cpu().numpy().view(np.uint32)), canon(wl, wr))
    finally:
        cuda_ctx.set_option("cuda_probe_layout", "auto")


def test_count_only_tiles_in_auto_mode(cuda_ctx, oracle):
    """option auto: count-only launches of position-sorted dense tiles go to the staged kernel (host tiles through the
    pipeline, device tiles through sq_probe_count_device); the per-row counts equal the oracle's either way"""
    cuda_ctx.set_option("cuda_probe_layout", "packed")
    try:
        b, p = sn.synth.cfg5(scale=0.004)
        q = sort_side(p)
        idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
        want = oracle.OracleIndex(b["key"], b["start"], b["end"]).counts(q["key"], q["start"], q["end"])
        st = sn.CudaStream(cuda_ctx)
        cols = {k: cuda_ctx.pinned_copy(q[k]) for k in ("key", "start", "end")}
        cuts = np.linspace(0, len(want), 6).astype(np.int64)
        got = []
        for a, z in zip(cuts, cuts[1:]):
            t = st.submit(idx, cols["key"][a:z], cols["start"][a:z], cols["end"][a:z], N.TILE_COUNT_ONLY)
            n, left, right, counts = st.collect(t)
            assert left is None and n == int(want[a:z].sum())
            got.append(counts.copy())
        assert np.array_equal(np.concatenate(got), want)
        dev = torch.device("cuda", 0)
        d = {k: torch.from_numpy(q[k].view(np.int64) if k == "key" else q[k]).to(dev) for k in ("key", "start", "end")}
        st2 = sn.CudaStream(cuda_ctx, cuda_stream=torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            assert st2.probe_count_device(idx, d["key"], d["start"], d["end"]) == int(want.sum())
            assert np.array_equal(st2.counts(), want)
    finally:
        cuda_ctx.set_option("cuda_probe_layout", "auto")
