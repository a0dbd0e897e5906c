"""GPU: position ids (option sequila.cuda_build_ids = positions, include/sequila_cuda.h).  The index then hands out
positions in its (key, start) order and keeps every payload column in that order; the joined ROWS must be the ones the
default mode (and the oracle) give, in the same order — only the meaning of left_idx changes.  `take` of the build
side = interval_join.rs:1620-1632."""
import numpy as np
import pytest

import sequila_native_b200 as sn

pytestmark = pytest.mark.gpu


@pytest.fixture()
def positions(cuda_ctx):
    """a context option: read by the next build"""
    def mode(on):
        cuda_ctx.set_option("sequila.cuda_build_ids", "positions" if on else "rows")
    yield mode
    cuda_ctx.set_option("sequila.cuda_build_ids", "rows")


def both_modes(ctx, positions, b):
    positions(False)
    rows = sn.CudaIndex.build(ctx, b["key"], b["start"], b["end"])
    positions(True)
    pos = sn.CudaIndex.build(ctx, b["key"], b["start"], b["end"])
    positions(False)
    assert not rows.uses_positions and pos.uses_positions
    return rows, pos


def workloads():
    b2, p2 = sn.synth.cfg2(scale=0.02)
    b3, p3 = sn.synth.cfg3(scale=0.002)
    b4, p4 = sn.synth.cfg4(scale=0.01)
    return {"cfg2": (b2, p2), "cfg3": (b3, p3), "cfg4": (b4, p4)}


@pytest.mark.parametrize("layout", ["auto", "packed", "soa"])
@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
def test_positions_name_the_same_rows_in_the_same_order(cuda_ctx, oracle, positions, name, layout):
    b, p = workloads()[name]
    cuda_ctx.set_option("sequila.cuda_probe_layout", layout)
    try:
        rows, pos = both_modes(cuda_ctx, positions, b)
        perm = pos.position_rows()
        assert np.array_equal(np.sort(perm), np.arange(len(perm), dtype=np.uint32))  # a permutation of the build rows
        # positions are in (key, start) order, equal starts in build-row order (the sort is stable)
        k, s = b["key"][perm], b["start"][perm]
        same_key = k[1:] == k[:-1]
        assert np.all(s[1:][same_key] >= s[:-1][same_key])
        tie = same_key & (s[1:] == s[:-1])
        assert np.all(perm[1:][tie] > perm[:-1][tie])
        st_r, st_p = sn.CudaStream(cuda_ctx), sn.CudaStream(cuda_ctx)
        n = st_r.probe_count(rows, p["key"], p["start"], p["end"])
        assert st_p.probe_count(pos, p["key"], p["start"], p["end"]) == n
        lr, rr, cr = st_r.emit_pairs()
        lp, rp, cp = st_p.emit_pairs()
        assert np.array_equal(rr, rp) and np.array_equal(cr, cp)
        assert np.array_equal(perm[lp], lr)  # same pairs, same order
        ol, orr, oc = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
        assert np.array_equal(oc, cp)
        assert np.array_equal(oracle.sorted_pairs(perm[lp], rp), oracle.sorted_pairs(ol, orr))
        # payload: 4 / 8 / 16-byte columns come back as the values of the build rows
        v4 = (np.arange(len(perm), dtype=np.int32) * 3 + 1)
        v8 = np.arange(len(perm), dtype=np.int64) * 1_000_003 - 7
        v16 = np.stack([np.arange(len(perm), dtype=np.int64), -np.arange(len(perm), dtype=np.int64)], axis=1)
        for v in (v4, v8, v16):
            got_r = st_r.gather_build(rows.add_column(v), v.dtype if v.ndim == 1 else np.dtype((np.int64, 2)))
            got_p = st_p.gather_build(pos.add_column(v), v.dtype if v.ndim == 1 else np.dtype((np.int64, 2)))
            assert np.array_equal(got_r, v[lr]) and np.array_equal(got_p, got_r)
    finally:
        cuda_ctx.set_option("sequila.cuda_probe_layout", "auto")


def test_packed_columns_and_device_columns(cuda_ctx, positions):
    import torch
    dev = torch.device("cuda", 0)
    b, p = sn.synth.cfg2(scale=0.05)
    to = lambda v: torch.from_numpy(np.ascontiguousarray(v.view(np.int64) if v.dtype == np.uint64 else v)).to(dev)
    bd, pd = {k: to(v) for k, v in b.items()}, {k: to(v) for k, v in p.items()}
    outs = {}
    for on in (False, True):
        positions(on)
        idx = sn.CudaIndex.build_device(cuda_ctx, bd["key"], bd["start"], bd["end"])
        positions(False)
        assert idx.uses_positions == on
        st = sn.CudaStream(cuda_ctx)
        n = st.probe_count_device(idx, pd["key"], pd["start"], pd["end"])
        left = torch.empty(n, dtype=torch.int32, device=dev)
        right = torch.empty(n, dtype=torch.int32, device=dev)
        assert st.probe_join_device(idx, pd["key"], pd["start"], pd["end"], left, right) == n
        cols = [idx.add_column_device(bd[k].to(torch.int32) if k == "contig" else bd[k]) for k in ("contig", "start", "end")]
        pack = idx.pack_columns(cols)
        o = [torch.full((n,), -1, dtype=torch.int32, device=dev) for _ in range(3)]
        st.gather_pack_device(pack, o)
        single = torch.empty(n, dtype=torch.int32, device=dev)
        st.gather_build_device(cols[2], single)
        assert torch.equal(single, o[2])
        rows_of_left = left.cpu().numpy().view(np.uint32)
        if on:
            ptr = idx.position_rows_device_ptr()
            assert ptr != 0
            rows_of_left = idx.position_rows()[rows_of_left]
        else:
            assert idx.position_rows_device_ptr() == 0
        outs[on] = [x.cpu().numpy() for x in o] + [rows_of_left, right.cpu().numpy()]
        assert np.array_equal(outs[on][1], b["start"][rows_of_left])
    for x, y in zip(outs[False], outs[True]):
        assert np.array_equal(x, y)


def test_nearest_names_the_same_rows(cuda_ctx, positions):
    b, p = sn.synth.cfg4(scale=0.01)
    # duplicates of (key, start) with different ends: the tie-breaks of nearest() compare (end, row)
    b = {k: np.concatenate([v, v[:2000]]) for k, v in b.items()}
    b["end"][-2000:] += np.arange(2000, dtype=np.int32) % 3
    rows, pos = both_modes(cuda_ctx, positions, b)
    perm = pos.position_rows()
    got_r = sn.CudaStream(cuda_ctx).probe_nearest(rows, p["key"], p["start"], p["end"])
    got_p = sn.CudaStream(cuda_ctx).probe_nearest(pos, p["key"], p["start"], p["end"])
    null = got_r == sn._native.NULL_INDEX
    assert np.array_equal(got_p == sn._native.NULL_INDEX, null)
    assert np.array_equal(perm[got_p[~null]], got_r[~null])


def test_rows_mode_has_no_position_map_and_empty_build_side(cuda_ctx, positions):
    b, p = sn.synth.cfg2(scale=0.001)
    idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
    with pytest.raises(sn.SequilaCudaError):
        idx.position_rows()
    positions(True)
    empty = sn.CudaIndex.build(cuda_ctx, b["key"][:0], b["start"][:0], b["end"][:0])
    positions(False)
    assert empty.uses_positions and len(empty.position_rows()) == 0
    cid = empty.add_column(np.empty(0, np.int32))
    st = sn.CudaStream(cuda_ctx)
    assert st.probe_count(empty, p["key"], p["start"], p["end"]) == 0
    l, r, c = st.emit_pairs()
    assert len(l) == 0 and int(c.sum()) == 0
    assert len(st.gather_build(cid, np.int32)) == 0
    with pytest.raises(sn.SequilaCudaError):
        cuda_ctx.set_option("sequila.cuda_build_ids", "sorted")
