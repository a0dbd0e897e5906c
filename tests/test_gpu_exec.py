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


BUILD_IDS = [None]


@pytest.fixture(autouse=True, params=["positions", "rows"])
def build_ids(request):
    """Every test runs with the node's default (position ids: build payload kept in the index's sorted order) and with
    `SET sequila.cuda_build_ids TO rows` (payload in build-row order): the output must not depend on it."""
    BUILD_IDS[0] = request.param
    yield request.param
    BUILD_IDS[0] = None


def cuda_config():
    cfg = sn.SequilaConfig()
    sn.apply_set(cfg, "SET sequila.interval_join_algorithm TO cuda")
    if BUILD_IDS[0] == "rows":  # "positions" is what the node picks by itself
        sn.apply_set(cfg, "SET sequila.cuda_build_ids TO rows")
    return cfg


def run_join(left, right, cond, equi=True, batch_rows=None, projection=None, coalesce=True):
    f = IV.parse_condition_sql(cond, "a", COLS, "b", COLS)
    on = [("contig", "contig")] if equi else []
    plan = optimize(HashJoinDesc(left.schema, right.schema, on, f, projection=projection), cuda_config())
    assert isinstance(plan, IntervalJoinExec)
    lb = [left] if batch_rows is None else [left.slice(i, batch_rows) for i in range(0, max(left.num_rows, 1), batch_rows)]
    rb = [right] if batch_rows is None else [right.slice(i, batch_rows) for i in range(0, max(right.num_rows, 1), batch_rows)]
    out = list(plan.execute(lb, rb, coalesce=coalesce))
    if coalesce:
        assert 1 <= len(out) <= len(rb)  # probe batches leave as tiles of sequila.cuda_coalesce_rows rows
    else:
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


@pytest.mark.parametrize("coalesce", [True, False])
def test_many_batches_keep_probe_order_and_match_oracle(oracle, coalesce):
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
    _, out = run_join(left, right, Q1, batch_rows=2000, coalesce=coalesce)
    if coalesce:
        assert len(out) == 1  # 5000 probe rows < sequila.cuda_coalesce_rows: one tile, one output batch
    ol, orr, _ = oracle.join(bc.astype(np.uint64), bs, be, pc.astype(np.uint64), ps, pe)
    want = sorted(zip(names[bc][ol].tolist(), bs[ol].tolist(), be[ol].tolist(), names[pc][orr].tolist(), ps[orr].tolist(), pe[orr].tolist()))
    got = [tuple(r) for r in rows_of(out)]
    assert sorted(got) == want
    # probe order preserved inside every output batch: maintains_input_order = [false, true]
    off = 0
    for b, rb in zip(out, [right] if coalesce else [right.slice(i, 2000) for i in range(0, npq, 2000)]):
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


# ---- dictionary-encoded contig columns ("dictionary-encode contig once") -------------
def dict_table(rows, index_type=pa.int32(), value_type=pa.string(), dictionary=None):
    """contig as DictionaryArray<index_type, value_type>; `dictionary` fixes the value order (may hold unused values)"""
    values = list(dictionary) if dictionary is not None else sorted({r[0] for r in rows}, reverse=True)
    idx = pa.array([values.index(r[0]) for r in rows], index_type)
    contig = pa.DictionaryArray.from_arrays(idx, pa.array(values, value_type))
    return pa.record_batch([contig, pa.array([r[1] for r in rows], pa.int32()), pa.array([r[2] for r in rows], pa.int32())],
                           names=COLS)


@pytest.mark.parametrize("index_type", [pa.int8(), pa.uint16(), pa.int32(), pa.uint32()])
@pytest.mark.parametrize("value_type", [pa.string(), pa.large_string()])
def test_dictionary_contig_fixture_equi_join_16_rows(golden, index_type, value_type):
    left, right = dict_table(golden["reads"], index_type, value_type), dict_table(golden["targets"], index_type, value_type)
    plan, out = run_join(left, right, Q1)
    assert sort_rows(rows_of(out)) == sort_rows(golden["equi_rows"])
    want = pa.dictionary(index_type, value_type)
    assert out[0].schema == plan.schema()
    assert out[0].schema.field(0).type == want and out[0].schema.field(3).type == want  # types unchanged (Appendix A4)
    out[0].validate(full=True)


def test_dictionary_batches_with_their_own_dictionaries_and_a_utf8_side(oracle):
    """every build batch brings its own dictionary (different order, unused values); the probe side is plain Utf8 in
    one run and dictionary-encoded in the other: same rows as the all-Utf8 join (the key hash of a dictionary
    value is the hash of the string)."""
    rng = np.random.default_rng(21)
    names = ["chr1", "chr2", "chrX", "chrUn_gl000220", "chr10"]
    def rows(n):
        c = rng.integers(0, len(names), n)
        s = rng.integers(0, 5000, n)
        return [[names[c[i]], int(s[i]), int(s[i] + rng.integers(0, 200))] for i in range(n)]
    lrows, rrows = rows(3000), rows(2500)
    f = IV.parse_condition_sql(Q1, "a", COLS, "b", COLS)

    def run(lbatches, rbatches):
        plan = optimize(HashJoinDesc(lbatches[0].schema, rbatches[0].schema, [("contig", "contig")], f), cuda_config())
        return rows_of(list(plan.execute(lbatches, rbatches)))

    want = run([table(lrows)], [table(rrows)])
    assert len(want) > 1000
    orders = [names, names[::-1] + ["unused"], ["zz"] + names[2:] + names[:2]]
    lb = [dict_table(lrows[i * 1000:(i + 1) * 1000], dictionary=orders[i]) for i in range(3)]
    rb_utf8 = [table(rrows[i:i + 700]) for i in range(0, len(rrows), 700)]
    rb_dict = [dict_table(rrows[i:i + 700], pa.int8(), dictionary=orders[(i // 700) % 3]) for i in range(0, len(rrows), 700)]
    assert sort_rows(run(lb, rb_utf8)) == sort_rows(want)
    assert sort_rows(run(lb, rb_dict)) == sort_rows(want)


def test_dictionary_payload_column_with_nulls_and_projection():
    name_vals = pa.array(["geneA", "geneB"])
    name = pa.DictionaryArray.from_arrays(pa.array([0, None, 1, 0], pa.int16()), name_vals)
    left = pa.record_batch([pa.array(["c", "c", "c", "d"]), pa.array([10, 20, 30, 10], pa.int32()),
                            pa.array([15, 25, 35, 15], pa.int32()), name], names=COLS + ["name"])
    right = pa.record_batch([pa.array(["c", "d"]), pa.array([0, 0], pa.int32()), pa.array([100, 100], pa.int32())], names=COLS)
    f = IV.parse_condition_sql(Q1, "a", COLS + ["name"], "b", COLS)
    plan = optimize(HashJoinDesc(left.schema, right.schema, [("contig", "contig")], f, projection=[3, 1, 4]), cuda_config())
    out = list(plan.execute([left], [right]))
    got = sorted(rows_of(out), key=lambda r: (r[2], r[1]))
    assert got == [["geneA", 10, "c"], [None, 20, "c"], ["geneB", 30, "c"], ["geneA", 10, "d"]]
    assert out[0].schema.field(0).type == pa.dictionary(pa.int16(), pa.string())
    out[0].validate(full=True)


def test_dictionary_index_type_too_small_for_the_unified_build_dictionary():
    lb = [dict_table([[f"k{i}_{j}", 1, 2] for j in range(100)], pa.int8()) for i in range(2)]  # 200 distinct values, int8 holds 128
    right = table([["k0_0", 0, 5]])
    f = IV.parse_condition_sql(Q1, "a", COLS, "b", COLS)
    plan = optimize(HashJoinDesc(lb[0].schema, right.schema, [("contig", "contig")], f), cuda_config())
    with pytest.raises(ExecutionError) as e:
        list(plan.execute(lb, [right]))
    assert "do not fit index type" in str(e.value)


def test_coalesced_tiles_equal_per_batch_joins(oracle):
    """sq_exec_probe_push / _pop: 8192-row probe batches (DataFusion's default batch size, interval_join.rs:1192-1233)
    leave as tiles of `sequila.cuda_coalesce_rows` rows; the rows — Utf8, dictionary-encoded (a dictionary per batch) and
    nullable payload columns included — equal those of joining every batch on its own, in the same probe order"""
    rng = np.random.default_rng(21)
    b, p = sn.synth.cfg5(scale=0.002)
    names = np.array(sn.synth.CONTIG_NAMES)
    left = pa.record_batch([pa.array(names[b["contig"]]), pa.array(b["start"]), pa.array(b["end"])], names=COLS)
    n = len(p["key"])
    tag_vals = rng.integers(0, 50, n)
    tags = pa.array([None if v == 7 else f"t{v}" for v in tag_vals])
    score = pa.array(np.where(rng.random(n) < 0.1, np.nan, rng.random(n)), mask=rng.random(n) < 0.05)
    rcols = COLS + ["tag", "score"]
    f = IV.parse_condition_sql(Q1, "a", COLS, "b", rcols)
    rbatches = []
    for i in range(0, n, 8192):
        sl = slice(i, min(i + 8192, n))
        tag_b = tags.slice(i, sl.stop - i).dictionary_encode()  # every batch brings its own dictionary
        rbatches.append(pa.record_batch([pa.array(names[p["contig"][sl]]), pa.array(p["start"][sl]), pa.array(p["end"][sl]),
                                         tag_b, score.slice(i, sl.stop - i)], names=rcols))

    def run(coalesce, target=None):
        plan = optimize(HashJoinDesc(left.schema, rbatches[0].schema, [("contig", "contig")], f), cuda_config())
        if target:
            plan.set_option("sequila.cuda_coalesce_rows", target)
        out = list(plan.execute([left], rbatches, coalesce=coalesce))
        m = plan.metrics()
        plan.close()
        return out, m
    per_batch, m0 = run(False)
    assert len(per_batch) == len(rbatches)
    want = rows_of(per_batch)
    canon_rows = lambda rows: sorted(map(repr, rows))
    for target in (None, 20000, 50000, 1):
        out, m = run(True, target)
        if target == 1:
            assert len(out) == len(rbatches)  # every pushed batch is due at once: the reference's one batch out per batch in
        elif target is None:
            assert len(out) == 1
        else:
            assert 1 < len(out) < len(rbatches)
        got = rows_of(out)
        assert len(got) == len(want) == m.output_rows
        assert canon_rows(got) == canon_rows(want)
        # probe order: the probe-side columns of the output are the same sequence in both modes
        assert [r[3:6] for r in got] == [r[3:6] for r in want]
        assert m.input_rows == n and m.input_batches == len(rbatches)


def test_null_keys_and_null_coordinates():
    """create_hashes skips NULL key slots (interval_join.rs:1037, 1211): a NULL contig never meets the rows whose value slot
    holds the same bytes, only other NULL keys; a NULL start / end is an error here (the reference reads garbage there)"""
    left = pa.record_batch([pa.array(["", "a", None, "a"]), pa.array([1, 10, 100, 1000], pa.int32()), pa.array([5, 20, 200, 2000], pa.int32())],
                           names=COLS)
    right = pa.record_batch([pa.array([None, "", "a"]), pa.array([0, 0, 0], pa.int32()), pa.array([5000, 5000, 5000], pa.int32())], names=COLS)
    _, out = run_join(left, right, Q1)
    rows = sorted(map(repr, rows_of(out)))
    want = sorted(map(repr, [[None, 100, 200, None, 0, 5000], ["", 1, 5, "", 0, 5000], ["a", 10, 20, "a", 0, 5000], ["a", 1000, 2000, "a", 0, 5000]]))
    assert rows == want
    bad = pa.record_batch([pa.array(["a"]), pa.array([None], pa.int32()), pa.array([5], pa.int32())], names=COLS)
    with pytest.raises(ExecutionError) as e:
        run_join(left, bad, Q1)
    assert "NULL" in str(e.value) and "pos_start" in str(e.value)
    with pytest.raises(ExecutionError) as e:
        run_join(bad, right, Q1)
    assert "NULL" in str(e.value)


def test_pipelined_tiles_many_small_tiles_and_an_error_behind_a_tile_in_flight(oracle):
    """tiles are pipelined two deep (sq_exec_probe_pop): the output of a tile leaves one call after the tile went in, in
    probe order; a tile that fails on the host half (NULL coordinate) reports its error while the previous tile's worker is
    still on the GPU, and closing the node afterwards joins that worker"""
    b, p = sn.synth.cfg5(scale=0.0005)
    names = np.array(sn.synth.CONTIG_NAMES)
    tab = lambda s: pa.record_batch([pa.array(names[s["contig"]].tolist()), pa.array(s["start"]), pa.array(s["end"])], names=COLS)
    left, right = tab(b), tab(p)
    f = IV.parse_condition_sql(Q1, "a", COLS, "b", COLS)
    plan = optimize(HashJoinDesc(left.schema, right.schema, [("contig", "contig")], f), cuda_config())
    plan.set_option("cuda_coalesce_rows", 3000)
    plan.collect_build([left])
    rbatches = [right.slice(i, 1000) for i in range(0, right.num_rows, 1000)]
    out = list(plan.probe_batches(rbatches))
    assert len(out) == -(-right.num_rows // 3000)  # one output batch per tile of three batches
    ol, orr, _ = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    got = rows_of(out)
    assert len(got) == len(ol)
    want_right = p["start"][np.sort(orr, kind="stable")]
    assert [r[4] for r in got] == want_right.tolist()  # right side in probe order across all tiles
    # an error in the fourth tile while the third is in flight
    bad = pa.record_batch([pa.array(["chr1"]), pa.array([None], pa.int32()), pa.array([5], pa.int32())], names=COLS)
    seq = rbatches[:9] + [bad] + rbatches[9:12]
    with pytest.raises(ExecutionError) as e:
        list(plan.probe_batches(seq, partition=1))
    assert "NULL" in str(e.value)
    plan.close()


def test_concurrent_partitions_of_one_node_with_coalescing(oracle):
    """four host threads drive four partitions of ONE exec node (one sq_stream each, one shared index), each coalescing its
    own 4096-row batches: the union of their output rows is the oracle's join"""
    import concurrent.futures as cf
    b, p = sn.synth.cfg5(scale=0.002)
    names = np.array(sn.synth.CONTIG_NAMES)
    left = pa.record_batch([pa.array(names[b["contig"]]), pa.array(b["start"]), pa.array(b["end"])], names=COLS)
    right = pa.record_batch([pa.array(names[p["contig"]]), pa.array(p["start"]), pa.array(p["end"])], names=COLS)
    batches = [right.slice(i, 4096) for i in range(0, right.num_rows, 4096)]
    f = IV.parse_condition_sql(Q1, "a", COLS, "b", COLS)
    plan = optimize(HashJoinDesc(left.schema, right.schema, [("contig", "contig")], f), cuda_config())
    plan.set_option("sequila.cuda_coalesce_rows", 30000)
    plan.collect_build([left])
    P = 4

    def drive(part):
        return rows_of(list(plan.probe_batches(batches[part::P], partition=part)))
    with cf.ThreadPoolExecutor(P) as pool:
        got = [r for rows in pool.map(drive, range(P)) for r in rows]
    ol, orr, _ = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    want = sorted(zip(names[b["contig"]][ol].tolist(), b["start"][ol].tolist(), b["end"][ol].tolist(),
                      names[p["contig"]][orr].tolist(), p["start"][orr].tolist(), p["end"][orr].tolist()))
    assert sorted(tuple(r) for r in got) == want
    m = plan.metrics()
    assert m.input_rows == right.num_rows and m.output_rows == len(want)
    plan.close()


def test_parquet_tables_through_the_exec_node(golden, oracle, tmp_path):
    """The reference's benchmarks read Parquet (benches/databio_benchmark.rs:278-283).  Parquet is decoded by the host's
    Arrow reader — as DataFusion's ParquetExec does — and the RecordBatches it yields (dictionary-encoded contig column,
    row groups = batches) go through the exec node: the fixture tables give the reference's 16 rows, a cfg5-shaped pair of
    files the oracle's rows.  (A device-side Parquet decoder is out of scope)"""
    import pyarrow.parquet as pq
    reads, targets = table(golden["reads"], pa.int64()), table(golden["targets"], pa.int64())
    pq.write_table(pa.Table.from_batches([reads]), tmp_path / "reads.parquet", use_dictionary=["contig"])
    pq.write_table(pa.Table.from_batches([targets]), tmp_path / "targets.parquet", use_dictionary=["contig"])
    lt = pq.read_table(tmp_path / "reads.parquet", read_dictionary=["contig"])
    rt = pq.read_table(tmp_path / "targets.parquet", read_dictionary=["contig"])
    f = IV.parse_condition_sql(Q1, "a", COLS, "b", COLS)
    plan = optimize(HashJoinDesc(lt.schema, rt.schema, [("contig", "contig")], f), cuda_config())
    out = list(plan.execute(lt.to_batches(), rt.to_batches()))
    assert sort_rows(rows_of(out)) == sort_rows(golden["equi_rows"])
    plan.close()

    b, p = sn.synth.cfg5(scale=0.001)
    names = np.array(sn.synth.CONTIG_NAMES)
    for name, side in (("b", b), ("p", p)):
        t = pa.table([pa.array(names[side["contig"]]), pa.array(side["start"]), pa.array(side["end"])], names=COLS)
        pq.write_table(t, tmp_path / f"{name}.parquet", row_group_size=8192, use_dictionary=["contig"])
    lt = pq.ParquetFile(tmp_path / "b.parquet", read_dictionary=["contig"])
    rt = pq.ParquetFile(tmp_path / "p.parquet", read_dictionary=["contig"])
    lb = list(lt.iter_batches(batch_size=8192))
    rb = list(rt.iter_batches(batch_size=8192))
    plan = optimize(HashJoinDesc(lb[0].schema, rb[0].schema, [("contig", "contig")], f), cuda_config())
    got = sorted(tuple(r) for r in rows_of(list(plan.execute(lb, rb))))
    ol, orr, _ = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    want = sorted(zip(names[b["contig"]][ol].tolist(), b["start"][ol].tolist(), b["end"][ol].tolist(),
                      names[p["contig"]][orr].tolist(), p["start"][orr].tolist(), p["end"][orr].tolist()))
    assert got == want
    plan.close()
