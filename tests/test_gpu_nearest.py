"""GPU: Algorithm::CoitreesNearest on the flat index (sq_probe_nearest, alg=CudaNearest) against the
oracle: identical where the reference's result is determined (no overlap: `nearest()`; exactly one
overlap; key miss: NULL), a member of the overlap set where the reference reports "an arbitrary one"."""
import numpy as np
import pyarrow as pa
import pytest

import sequila_native_b200 as sn
from sequila_native_b200 import _native as N
from sequila_native_b200 import intervals as IV
from sequila_native_b200.interval_join import HashJoinDesc, optimize

pytestmark = pytest.mark.gpu


def check(ctx, oracle, b, p):
    idx = sn.CudaIndex.build(ctx, b["key"], b["start"], b["end"])
    st = sn.CudaStream(ctx)
    got = st.probe_nearest(idx, p["key"], p["start"], p["end"])
    want, ov = oracle.nearest(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    assert np.array_equal(got[~ov], want[~ov])          # nearest() and NULLs: bit-exact
    ol, orr, oc = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    assert np.array_equal(ov, oc > 0)
    single = oc == 1
    assert np.array_equal(got[single], want[single])    # one overlap: nothing arbitrary about it
    hits = set(zip(orr.tolist(), ol.tolist()))
    multi = np.flatnonzero(oc > 1)
    assert all((int(i), int(got[i])) in hits for i in multi)
    return st, idx, got


def test_reference_table_through_the_c_abi(cuda_ctx, oracle, golden):
    A, B = golden["nearest_a"], golden["nearest_b"]
    ids = {}
    key = lambda r: ids.setdefault((r[0], r[1]), len(ids) + 1)
    b = {"key": np.array([key(r) for r in A], np.uint64), "start": np.array([r[2] for r in A], np.int32),
         "end": np.array([r[3] - 1 for r in A], np.int32)}
    p = {"key": np.array([key(r) for r in B], np.uint64), "start": np.array([r[2] for r in B], np.int32),
         "end": np.array([r[3] - 1 for r in B], np.int32)}
    _, _, got = check(cuda_ctx, oracle, b, p)
    rows = [(A[int(x)] if x != N.NULL_INDEX else [None] * 4) + B[i] for i, x in enumerate(got)]
    assert sorted(rows, key=str) == sorted(golden["nearest_rows"], key=str)


def test_random_sparse_and_dense(cuda_ctx, oracle):
    rng = np.random.default_rng(77)
    for it in range(25):
        nb, npq, nk = int(rng.integers(1, 4000)), int(rng.integers(1, 3000)), int(rng.integers(1, 6))
        span, wmax = int(rng.integers(50, 200000)), int(rng.integers(1, 40))
        b = {"key": rng.integers(0, nk, nb).astype(np.uint64) + 1, "start": rng.integers(0, span, nb).astype(np.int32)}
        b["end"] = (b["start"] + rng.integers(0, wmax, nb)).astype(np.int32)
        p = {"key": rng.integers(0, nk + 1, npq).astype(np.uint64) + 1, "start": rng.integers(0, span, npq).astype(np.int32)}
        p["end"] = (p["start"] + rng.integers(0, wmax, npq)).astype(np.int32)
        if it % 4 == 0:  # many equal starts: the (start, end, row) tie rules of the two candidates
            b["start"] = (b["start"] // 64 * 64).astype(np.int32)
            b["end"] = (b["start"] + rng.integers(0, 3, nb)).astype(np.int32)
        check(cuda_ctx, oracle, b, p)


def test_synthetic_cfg2(cuda_ctx, oracle):
    b, p = sn.synth.cfg2(scale=0.05)
    st, idx, got = check(cuda_ctx, oracle, b, p)
    # gathers understand SQ_NULL_INDEX: zero value under a cleared validity bit
    col = idx.add_column(b["start"])
    vals = st.gather_build(col, np.int32)
    ok = got != N.NULL_INDEX
    assert np.array_equal(vals[ok], b["start"][got[ok]]) and not vals[~ok].any()


def test_exec_node_nearest_table(golden):
    cols = ["contig", "strand", "start", "end"]
    mk = lambda rows: pa.record_batch([pa.array([r[0] for r in rows]), pa.array([r[1] for r in rows]),
                                       pa.array([r[2] for r in rows], pa.int32()), pa.array([r[3] for r in rows], pa.int32())],
                                      names=cols)
    a, b = mk(golden["nearest_a"]), mk(golden["nearest_b"])
    cfg = sn.SequilaConfig()
    sn.apply_set(cfg, "SET sequila.interval_join_algorithm TO CudaNearest")
    f = IV.parse_condition_sql("a.start < b.end AND a.end > b.start", "a", cols, "b", cols)
    plan = optimize(HashJoinDesc(a.schema, b.schema, [("contig", "contig"), ("strand", "strand")], f), cfg)
    assert plan.display().endswith("alg=CudaNearest")
    out = list(plan.execute([a], [b]))[0]
    rows = [list(r) for r in zip(*[c.to_pylist() for c in out.columns])]
    assert sorted(rows, key=str) == sorted(golden["nearest_rows"], key=str)   # integration_test.rs:387-396
    assert out.column(0).null_count == 2 and out.column(4).null_count == 0
