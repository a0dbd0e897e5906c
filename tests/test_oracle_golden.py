"""CPU: pin the oracle (coitrees restatement + join driver) to the reference's own golden tables,
to brute force and to the reference-shipped superintervals C++ library (oracle/_ref)."""
import numpy as np
import pytest

from helpers import encode_tables, rows_from_pairs, sort_rows

VARIANTS = [1, 8]  # scalar coitrees (nosimd.rs) and the AVX2 chunk tree (avx.rs)


@pytest.mark.parametrize("variant", VARIANTS)
def test_equi_and_range_16_rows(oracle, golden, variant):
    # integration_test.rs:67-118 / interval_join.rs:1748-1812
    L, R = encode_tables(golden["reads"], golden["targets"], equi=True)
    l, r, _ = oracle.join(L["key"], L["start"], L["end"], R["key"], R["start"], R["end"], variant)
    assert rows_from_pairs(L, R, l, r) == sort_rows(golden["equi_rows"])


@pytest.mark.parametrize("variant", VARIANTS)
def test_range_only_32_rows(oracle, golden, variant):
    # integration_test.rs:163-212: on=[(1, 1)] => one key for the whole build side
    L, R = encode_tables(golden["reads"], golden["targets"], equi=False)
    l, r, _ = oracle.join(L["key"], L["start"], L["end"], R["key"], R["start"], R["end"], variant)
    assert rows_from_pairs(L, R, l, r) == sort_rows(golden["range_rows"])


@pytest.mark.parametrize("variant", VARIANTS)
def test_closed_boundaries_10_rows(oracle, golden, variant):
    # integration_test.rs:214-291: a.start <= b.end AND a.end >= b.start
    L, R = encode_tables(golden["boundary_a"], golden["boundary_b"])
    l, r, _ = oracle.join(L["key"], L["start"], L["end"], R["key"], R["start"], R["end"], variant)
    assert rows_from_pairs(L, R, l, r) == sort_rows(golden["closed_rows"])


@pytest.mark.parametrize("variant", VARIANTS)
def test_strict_boundaries_6_rows(oracle, golden, variant):
    # integration_test.rs:293-350: a.start < b.end AND a.end > b.start  => both ends minus one
    # (intervals.rs:95-137: `ls < re` -> re-1 ; `le > rs` -> le-1)
    L, R = encode_tables(golden["boundary_a"], golden["boundary_b"])
    l, r, _ = oracle.join(L["key"], L["start"], L["end"] - 1, R["key"], R["start"], R["end"] - 1, variant)
    assert rows_from_pairs(L, R, l, r) == sort_rows(golden["strict_rows"])


def _random_case(rng, nb, npq, nk, span, wmax, inverted):
    bk = rng.integers(0, nk, nb).astype(np.uint64) * 7919 + 3
    pk = rng.integers(0, nk + 1, npq).astype(np.uint64) * 7919 + 3  # one key absent from the build side
    bs = rng.integers(-50, span, nb).astype(np.int32)
    be = (bs + rng.integers(0, wmax, nb)).astype(np.int32)
    ps = rng.integers(-50, span, npq).astype(np.int32)
    pe = (ps + rng.integers(0, wmax, npq)).astype(np.int32)
    if inverted:
        m = rng.random(nb) < 0.1
        be[m] = bs[m] - rng.integers(1, 20, int(m.sum())).astype(np.int32)
        m = rng.random(npq) < 0.1
        pe[m] = ps[m] - rng.integers(1, 20, int(m.sum())).astype(np.int32)
    return bk, bs, be, pk, ps, pe


@pytest.mark.parametrize("variant", VARIANTS)
def test_differential_vs_brute_force(oracle, variant):
    rng = np.random.default_rng(20261018 + variant)
    for it in range(120):
        case = _random_case(rng, int(rng.integers(0, 900)), int(rng.integers(0, 300)), int(rng.integers(1, 4)),
                            int(rng.integers(10, 5000)), int(rng.integers(1, 400)), inverted=(it % 3 == 0))
        bl, br = oracle.brute(*case)
        l, r, c = oracle.join(*case, variant=variant)
        assert np.all(np.diff(r.astype(np.int64)) >= 0), "right_idx must be non-decreasing"
        assert np.array_equal(oracle.sorted_pairs(l, r), oracle.sorted_pairs(bl, br))
        assert np.array_equal(np.bincount(r, minlength=len(case[3])).astype(np.uint32), c)


def test_dense_and_large_trees(oracle):
    # exercises the density cut-off (simple subtrees) and deep vEB recursion
    rng = np.random.default_rng(5)
    for nb, span, wmax in ((3000, 1000, 900), (20000, 100000, 3000), (70000, 10 ** 6, 50)):
        case = _random_case(rng, nb, 1500, 2, span, wmax, False)
        bl, br = oracle.brute(*case)
        for v in VARIANTS:
            l, r, _ = oracle.join(*case, variant=v)
            assert np.array_equal(oracle.sorted_pairs(l, r), oracle.sorted_pairs(bl, br))


def test_against_reference_superintervals(oracle):
    """oracle/_ref/libsi_ref.so is the reference's own superintervals.hpp compiled from /root/reference:
    its SuperIntervals arm must agree with Coitrees (interval_join.rs:1752-1758)."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libsi_ref.so not built (reference tree absent)")
    rng = np.random.default_rng(11)
    for it in range(60):
        case = _random_case(rng, int(rng.integers(1, 3000)), int(rng.integers(1, 500)), 3,
                            int(rng.integers(100, 100000)), int(rng.integers(1, 2000)), False)
        l, r, _ = oracle.join(*case, variant=8)
        rl, rr = oracle.ref_superintervals_join(*case)
        assert np.array_equal(oracle.sorted_pairs(l, r), oracle.sorted_pairs(rl, rr))


def test_reference_superintervals_timing_entry_points_count_the_same_pairs(oracle):
    """the probe-loop timer over the reference's library (bench.py's `reference_superintervals` anchor) visits the same
    pairs as the coitrees restatement, single- and multi-threaded"""
    if not oracle.ref_timing_available():
        pytest.skip("oracle/_ref/libsi_ref.so without the timing entry points")
    rng = np.random.default_rng(5)
    case = _random_case(rng, 20000, 30000, 5, 200000, 300, False)
    l, _, _ = oracle.join(*case, variant=8)
    si = oracle.RefSuperIntervalsIndex(*case[:3])
    for threads in (1, 3):
        sec, pairs = si.time_probe(*case[3:], threads=threads, batch_rows=1000)
        assert pairs == len(l) and sec > 0


def test_reference_fixture_through_superintervals_ref(oracle, golden):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libsi_ref.so not built")
    L, R = encode_tables(golden["reads"], golden["targets"], equi=True)
    l, r = oracle.ref_superintervals_join(L["key"], L["start"], L["end"], R["key"], R["start"], R["end"])
    assert rows_from_pairs(L, R, l, r) == sort_rows(golden["equi_rows"])


def test_synthetic_configs_fanout(oracle):
    """the generators reproduce the fan-out SURVEY.md §8(d) quotes (scaled sizes, same density)"""
    import sequila_native_b200 as sn
    want = {"cfg2": (0.83, 0.1), "cfg3": (22.0, 8.0), "cfg4": (97.0, 6.0), "cfg5": (6.48, 0.5)}
    scale = {"cfg2": 0.05, "cfg3": 0.02, "cfg4": 0.05, "cfg5": 0.0005}
    for name, (mean, tol) in want.items():
        b, p = sn.synth.CONFIGS[name](scale=scale[name])
        c = oracle.OracleIndex(b["key"], b["start"], b["end"]).counts(p["key"], p["start"], p["end"])
        assert abs(float(c.mean()) - mean) < tol, (name, float(c.mean()))


def test_nearest_matches_the_reference_table(oracle, golden):
    """integration_test.rs:352-399 (CoitreesNearest, strict operators => end - 1 on both sides)"""
    A, B = golden["nearest_a"], golden["nearest_b"]
    ids = {}

    def key(r):
        return ids.setdefault((r[0], r[1]), len(ids) + 1)
    bk = np.array([key(r) for r in A], np.uint64)
    pk = np.array([key(r) for r in B], np.uint64)
    left, ov = oracle.nearest(bk, [r[2] for r in A], [r[3] - 1 for r in A], pk, [r[2] for r in B], [r[3] - 1 for r in B])
    rows = [(A[int(x)] if x != oracle.NULL_INDEX else [None] * 4) + B[i] for i, x in enumerate(left)]
    assert sorted(rows, key=str) == sorted(golden["nearest_rows"], key=str)
    assert not ov.any()


def test_nearest_differential_vs_python_restatement(oracle):
    """the C++ nearest() against a line-by-line Python restatement of interval_join.rs:909-956"""
    rng = np.random.default_rng(5)
    for _ in range(30):
        nb, npq = int(rng.integers(1, 60)), int(rng.integers(1, 80))
        bs = rng.integers(0, 300, nb).astype(np.int32)
        be = (bs + rng.integers(0, 12, nb)).astype(np.int32)
        ps = rng.integers(0, 300, npq).astype(np.int32)
        pe = (ps + rng.integers(0, 12, npq)).astype(np.int32)
        bk = np.zeros(nb, np.uint64)
        pk = (rng.random(npq) < 0.1).astype(np.uint64)  # key 1 is absent from the build side
        left, ov = oracle.nearest(bk, bs, be, pk, ps, pe)
        order = sorted(range(nb), key=lambda j: (bs[j], be[j], j))  # stable sort by (first, last)
        for i in range(npq):
            if pk[i] == 1:
                assert left[i] == oracle.NULL_INDEX
                continue
            hits = [j for j in range(nb) if bs[j] <= pe[i] and be[j] >= ps[i]]
            if hits:
                assert ov[i] and int(left[i]) in hits
                continue
            lo = 0
            while lo < nb and bs[order[lo]] < pe[i]:
                lo += 1
            best, best_d = None, 2 ** 31 - 1
            for c in (max(lo - 1, 0), lo):
                if c < nb:
                    j = order[c]
                    d = bs[j] - pe[i] if pe[i] < bs[j] else (ps[i] - be[j] if be[j] < ps[i] else 0)
                    if d < best_d:
                        best, best_d = j, d
            assert int(left[i]) == best


def test_lapper_conversion_gives_the_same_multiset_for_non_negative_coordinates(oracle):
    """The reference's Lapper arm stores `start as u32 .. end as u32 + 1` and queries
    `find(start as u32, end as u32 + 1)` (interval_join.rs:711-717, 1005-1011; rust-lapper's find is
    half-open: iv.start < stop && iv.stop > start).  For coordinates >= 0 that is the closed predicate."""
    rng = np.random.default_rng(9)
    nb, npq = 700, 500
    bs = rng.integers(0, 5000, nb).astype(np.int64); be = bs + rng.integers(0, 60, nb)
    ps = rng.integers(0, 5000, npq).astype(np.int64); pe = ps + rng.integers(0, 60, npq)
    k = np.zeros(nb, np.uint64); kq = np.zeros(npq, np.uint64)
    lap = [(b, q) for q in range(npq) for b in range(nb) if bs[b] < pe[q] + 1 and be[b] + 1 > ps[q]]
    ol, orr, _ = oracle.join(k, bs, be, kq, ps, pe)
    assert sorted(lap) == sorted(zip(ol.tolist(), orr.tolist()))
