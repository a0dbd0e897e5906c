"""GPU: the exec-node layer (include/sequila_exec.h through interval_join.IntervalJoinExec) on Arrow
RecordBatches — the reference's integration tests (integration_test.rs:22-350, interval_join.rs
1740-1968) with `SET sequila.interval_join_algorithm TO cuda`."""
import numpy as np
import pyarrow as pa
import pytest

import sequila_native_b200 as sn
from sequila_native_b200 import intervals as IV
from sequila_native_b200.interval_join import ExecutionError, HashJoinDesc, IntervalJoinExec, optimize
from helpers import sort_rows

pytestmark = pytest.mark.gpu

COLS = ["contig", "pos_start", "pos_end"]


def table(rows, int_type=pa.int32(), str_type=pa.string()):
    return pa.record_batch([pa.array([r[0] for r in rows], str_type), pa.array([r[1] for r in rows], int_type),
                            pa.array([r[2] for r in rows], int_type)], names=COLS)


def cuda_config():
    cfg = sn.SequilaConfig()
    sn.apply_set(cfg, "SET sequila.interval_join_algorithm TO cuda")
    return cfg


def run_join(left, right, cond, equi=True, batch_rows=None, projection=None):
    f = IV.parse_condition_sql(cond, "a", COLS, "b", COLS)
    on = [("contig", "contig")] if equi else []
    plan = optimize(HashJoinDesc(left.schema, right.schema, on, f, projection=projection), cuda_config())
    assert isinstance(plan, IntervalJoinExec)
    lb = [left] if batch_rows is None else [left.slice(i, batch_rows) for i in range(0, max(left.num_rows, 1), batch_rows)]
    rb = [right] if batch_rows is None else [right.slice(i, batch_rows) for i in range(0, max(right.num_rows, 1), batch_rows)]
    out = list(plan.execute(lb, rb))
    assert len(out) == len(rb)  # one output batch per probe batch (interval_join.rs:1580-1640)
    return plan, out


def rows_of(batches):
    rows = []
    for b in batches:
        cols = [c.to_pylist() for c in b.columns]
        rows += [list(r) for r in zip(*cols)]
    return rows


Q1 = "a.pos_start <= b.pos_end AND a.pos_end >= b.pos_start"


@pytest.mark.parametrize("int_type", [pa.int32(), pa.int64()])
@pytest.mark.parametrize("str_type", [pa.string(), pa.large_string()])
def test_fixture_equi_join_16_rows(golden, int_type, str_type):
    """q1-coitrees.sql shape on reads.csv x targets.csv; BIGINT columns as in queries/q1-coitrees.sql:6,11"""
    plan, out = run_join(table(golden["reads"], int_type, str_type), table(golden["targets"], int_type, str_type), Q1)
    assert sort_rows(rows_of(out)) == sort_rows(golden["equi_rows"])
    assert out[0].schema == plan.schema() and out[0].schema.names == COLS + COLS
    m = plan.metrics()
    assert (m.build_input_rows, m.input_rows, m.output_rows, m.keys) == (len(golden["reads"]), len(golden["targets"]), 16, 2)


def test_fixture_range_only_32_rows(golden):
    _, out = run_join(table(golden["reads"]), table(golden["targets"]), Q1, equi=False)
    assert sort_rows(rows_of(out)) == sort_rows(golden["range_rows"])


def test_closed_and_strict_boundaries(golden):
    a, b = table(golden["boundary_a"]), table(golden["boundary_b"])
    _, out = run_join(a, b, "a.pos_start <= b.pos_end AND a.pos_end >= b.pos_start")
    assert sort_rows(rows_of(out)) == sort_rows(golden["closed_rows"])
    _, out = run_join(a, b, "a.pos_start < b.pos_end AND a.pos_end > b.pos_start")
    assert sort_rows(rows_of(out)) == sort_rows(golden["strict_rows"])


def test_many_batches_keep_probe_order_and_match_oracle(oracle):
    rng = np.random.default_rng(11)
    nb, npq = 6000, 5000
    names = np.array(["chr1", "chr2", "chrX", "chrUn_gl000220"])

    def side(n):
        c = rng.integers(0, 4, n)
        s = rng.integers(0, 50000, n).astype(np.int32)
        e = (s + rng.integers(0, 400, n)).astype(np.int32)
        return c, s, e
    bc, bs, be = side(nb)
    pc, ps, pe = side(npq)
    left = pa.record_batch([pa.array(names[bc]), pa.array(bs), pa.array(be)], names=COLS)
    right = pa.record_batch([pa.array(names[pc]), pa.array(ps), pa.array(pe)], names=COLS)
    _, out = run_join(left, right, Q1, batch_rows=2000)
    ol, orr, _ = oracle.join(bc.astype(np.uint64), bs, be, pc.astype(np.uint64), ps, pe)
    want = sorted(zip(names[bc][ol].tolist(), bs[ol].tolist(), be[ol].tolist(), names[pc][orr].tolist(), ps[orr].tolist(), pe[orr].tolist()))
    got = [tuple(r) for r in rows_of(out)]
    assert sorted(got) == want
    # probe order preserved inside every output batch: maintains_input_order = [false, true]
    off = 0
    for b, rb in zip(out, [right.slice(i, 2000) for i in range(0, npq, 2000)]):
        key = list(zip(b.column(3).to_pylist(), b.column(4).to_pylist(), b.column(5).to_pylist()))
        src = list(zip(rb.column(0).to_pylist(), rb.column(1).to_pylist(), rb.column(2).to_pylist()))
        it = iter(src)
        cur = next(it, None)
        for k in key:           # the output's probe columns are a (repeating) subsequence of the probe batch
            while cur is not None and cur != k:
                cur = next(it, None)
            assert cur is not None
        off += rb.num_rows


def test_projection_nulls_and_extra_columns():
    left = pa.record_batch([pa.array(["a", "a", "b"]), pa.array([1, 10, 5], pa.int32()), pa.array([4, 20, 9], pa.int32()),
                            pa.array([1.5, None, 3.5], pa.float64()), pa.array(["x", None, "zz"])],
                           names=COLS + ["score", "name"])
    right = pa.record_batch([pa.array(["a", "b", "c"]), pa.array([3, 9, 1], pa.int32()), pa.array([12, 9, 100], pa.int32()),
                             pa.array([None, 7, 8], pa.int16())], names=COLS + ["flag"])
    f = IV.parse_condition_sql(Q1, "a", COLS, "b", COLS)
    plan = optimize(HashJoinDesc(left.schema, right.schema, [("contig", "contig")], f, projection=[4, 3, 1, 8, 6]), cuda_config())
    out = list(plan.execute([left], [right]))[0]
    assert out.schema.names == ["name", "score", "pos_start", "flag", "pos_start"]
    got = sorted(zip(*[c.to_pylist() for c in out.columns]), key=lambda r: (r[2], r[4]))
    assert got == [("x", 1.5, 1, None, 3), ("zz", 3.5, 5, 7, 9), (None, None, 10, None, 3)]
    assert out.column(0).null_count == 1 and out.column(3).null_count == 2


def test_cast_overflow_is_the_reference_error(golden):
    big = 2 ** 31 + 5
    left = pa.record_batch([pa.array(["a"]), pa.array([1], pa.int64()), pa.array([big], pa.int64())], names=COLS)
    right = pa.record_batch([pa.array(["a"]), pa.array([1], pa.int64()), pa.array([2], pa.int64())], names=COLS)
    with pytest.raises(ExecutionError) as e:
        run_join(left, right, Q1)
    assert golden["cast_error_format"].replace("{}", str(big)) in str(e.value)  # interval_join.rs:1959-1965


def test_empty_sides():
    empty = table([])
    one = table([["a", 1, 2]])
    _, out = run_join(empty, one, Q1)
    assert out[0].num_rows == 0 and out[0].schema.names == COLS + COLS
    _, out = run_join(one, empty, Q1)
    assert out[0].num_rows == 0


def test_partitioned_mode_one_exec_per_partition(oracle):
    """PartitionMode::Partitioned (interval_join.rs:488-503): both sides hash-partitioned on the key, every
    partition builds its own index and probes only its rows; the union is the whole join."""
    from sequila_native_b200 import sharding
    from sequila_native_b200.interval_join import PARTITIONED
    rng = np.random.default_rng(21)
    nb, npq, nk = 4000, 3000, 6
    names = np.array([f"chr{i}" for i in range(nk)])
    bc, pc = rng.integers(0, nk, nb), rng.integers(0, nk, npq)
    bs = rng.integers(0, 30000, nb).astype(np.int32); be = (bs + rng.integers(0, 200, nb)).astype(np.int32)
    ps = rng.integers(0, 30000, npq).astype(np.int32); pe = (ps + rng.integers(0, 200, npq)).astype(np.int32)
    plan_keys = sharding.assign_keys_lpt(np.bincount(bc, minlength=nk), 2)
    got = []
    for part, (brows, prows) in enumerate(zip(sharding.route_rows(bc, plan_keys), sharding.route_rows(pc, plan_keys))):
        left = pa.record_batch([pa.array(names[bc[brows]]), pa.array(bs[brows]), pa.array(be[brows])], names=COLS)
        right = pa.record_batch([pa.array(names[pc[prows]]), pa.array(ps[prows]), pa.array(pe[prows])], names=COLS)
        f = IV.parse_condition_sql(Q1, "a", COLS, "b", COLS)
        plan = optimize(HashJoinDesc(left.schema, right.schema, [("contig", "contig")], f, partition_mode=PARTITIONED), cuda_config())
        assert "mode=Partitioned" in plan.display()
        for b in plan.execute([left], [right], partition=part):
            got += [tuple(r) for r in zip(*[c.to_pylist() for c in b.columns])]
    ol, orr, _ = oracle.join(bc.astype(np.uint64), bs, be, pc.astype(np.uint64), ps, pe)
    want = sorted(zip(names[bc][ol].tolist(), bs[ol].tolist(), be[ol].tolist(), names[pc][orr].tolist(), ps[orr].tolist(), pe[orr].tolist()))
    assert sorted(got) == want


def test_empty_projection_is_a_count_only_probe(oracle):
    """`SELECT count(*) FROM a JOIN b ON ...`: the join node projects no column, the output batches carry
    only their row counts and the pairs are never written (what the reference's benchmarks run)."""
    rng = np.random.default_rng(8)
    nb, npq = 5000, 4000
    bs = rng.integers(0, 40000, nb).astype(np.int32); be = (bs + rng.integers(0, 300, nb)).astype(np.int32)
    ps = rng.integers(0, 40000, npq).astype(np.int32); pe = (ps + rng.integers(0, 300, npq)).astype(np.int32)
    left = pa.record_batch([pa.array(["c"] * nb), pa.array(bs), pa.array(be)], names=COLS)
    right = pa.record_batch([pa.array(["c"] * npq), pa.array(ps), pa.array(pe)], names=COLS)
    plan, out = run_join(left, right, Q1, batch_rows=1500, projection=[])
    ol, _, _ = oracle.join(np.zeros(nb, np.uint64), bs, be, np.zeros(npq, np.uint64), ps, pe)
    assert all(b.num_columns == 0 for b in out) and sum(b.num_rows for b in out) == len(ol)
    assert plan.metrics().output_rows == len(ol)
