"""CPU: the selection surface around the hot path, mirrored from the reference's own tests:
Algorithm / SequilaConfig / SET parsing (session_context.rs:50-136), the join-filter parser
(intervals.rs:30-232; spellings of intervals.rs:258-505), the optimizer rule
(sequila_physical_planner.rs:28-148) and the EXPLAIN line (integration_test.rs:110, 204)."""
import pyarrow as pa
import pytest

import sequila_native_b200 as sn
from sequila_native_b200 import intervals as IV
from sequila_native_b200.interval_join import AUTO, HashJoinDesc, IntervalJoinExec, PlanError, optimize
from sequila_native_b200.intervals import BinaryExpr, Column, Literal

A_COLS = ["contig", "l_start", "l_end"]            # CREATE TABLE a (contig, l_start, l_end)   intervals.rs:250
B_COLS = ["contig", "name", "r_end", "r_start"]    # CREATE TABLE b (contig, name, r_end, r_start)


def extract(cond):
    return IV.try_parse(IV.parse_condition_sql(cond, "a", A_COLS, "b", B_COLS))


L_START, L_END = Column("l_start", 1), Column("l_end", 2)
R_START, R_END = Column("r_start", 3), Column("r_end", 2)


@pytest.mark.parametrize("cond", [
    "b.r_end >= a.l_start AND a.l_end >= b.r_start", "b.r_end >= a.l_start AND b.r_start <= a.l_end",
    "a.l_start <= b.r_end AND a.l_end >= b.r_start", "a.l_start <= b.r_end AND b.r_start <= a.l_end",
    "a.l_end >= b.r_start AND b.r_end >= a.l_start", "a.l_end >= b.r_start AND a.l_start <= b.r_end",
    "b.r_start <= a.l_end AND b.r_end >= a.l_start", "b.r_start <= a.l_end AND a.l_start <= b.r_end"])
def test_all_comp_combinations_for_gteq_lteq(cond):
    ci = extract(cond)
    assert (ci.left_interval.start, ci.left_interval.end) == (L_START, L_END)
    assert (ci.right_interval.start, ci.right_interval.end) == (R_START, R_END)


@pytest.mark.parametrize("cond", [
    "b.r_end > a.l_start AND a.l_end > b.r_start", "b.r_end > a.l_start AND b.r_start < a.l_end",
    "a.l_start < b.r_end AND a.l_end > b.r_start", "a.l_start < b.r_end AND b.r_start < a.l_end",
    "a.l_end > b.r_start AND b.r_end > a.l_start", "a.l_end > b.r_start AND a.l_start < b.r_end",
    "b.r_start < a.l_end AND b.r_end > a.l_start", "b.r_start < a.l_end AND a.l_start < b.r_end"])
def test_all_comp_combinations_for_gt_lt(cond):
    ci = extract(cond)
    assert ci.left_interval.start == L_START and ci.right_interval.start == R_START
    assert ci.left_interval.end == BinaryExpr(L_END, "-", Literal(1))     # strict => end - 1, intervals.rs:67-69
    assert ci.right_interval.end == BinaryExpr(R_END, "-", Literal(1))


def test_mixed_strictness():
    ci = extract("b.r_start < a.l_end AND a.l_start <= b.r_end")
    assert ci.left_interval.end == BinaryExpr(L_END, "-", Literal(1)) and ci.right_interval.end == R_END
    ci = extract("b.r_start <= a.l_end AND a.l_start < b.r_end")
    assert ci.left_interval.end == L_END and ci.right_interval.end == BinaryExpr(R_END, "-", Literal(1))


def test_unparsable_filters():
    f = IV.parse_condition_sql("b.r_start <= a.l_end OR a.l_start <= b.r_end", "a", A_COLS, "b", B_COLS)
    with pytest.raises(ValueError):
        IV.try_parse(f)
    assert IV.parse(f) is None and IV.parse(None) is None
    # two columns under one operand: the reference panics (intervals.rs:53-55, test at :494-499)
    f = IV.parse_condition_sql("(b.r_end - a.l_start) >= a.l_start AND a.l_end >= b.r_start", "a", A_COLS, "b", B_COLS)
    with pytest.raises(IV.PanicError):
        IV.try_parse(f)
    # the same slot twice panics too (intervals.rs:158-183)
    f = IV.parse_condition_sql("a.l_start <= b.r_end AND a.l_start <= b.r_start", "a", A_COLS, "b", B_COLS)
    with pytest.raises(IV.PanicError):
        IV.try_parse(f)
    # equality is not a range operator
    f = IV.parse_condition_sql("a.l_start = b.r_end AND a.l_end >= b.r_start", "a", A_COLS, "b", B_COLS)
    assert IV.parse(f) is None


def test_algorithm_from_str_and_display():
    for text, alg in [("coitrees", sn.Algorithm.Coitrees), ("IntervalTree", sn.Algorithm.IntervalTree),
                      ("ARRAYINTERVALTREE", sn.Algorithm.ArrayIntervalTree), ("lapper", sn.Algorithm.Lapper),
                      ("superintervals", sn.Algorithm.SuperIntervals), ("CoitreesNearest", sn.Algorithm.CoitreesNearest),
                      ("coitreescountoverlaps", sn.Algorithm.CoitreesCountOverlaps), ("cuda", sn.Algorithm.Cuda),
                      ("CUDA", sn.Algorithm.Cuda)]:
        assert sn.Algorithm.from_str(text) is alg
    assert str(sn.Algorithm.Cuda) == "Cuda" and str(sn.Algorithm.CoitreesCountOverlaps) == "CoitreesCountOverlaps"
    assert sn.Algorithm.default() is sn.Algorithm.Coitrees
    with pytest.raises(sn.ParseAlgorithmError) as e:
        sn.Algorithm.from_str("gpu")
    assert str(e.value) == "Can't parse 'gpu' as Algorithm"  # session_context.rs:98-101


def test_set_statements():
    cfg = sn.SequilaConfig()
    assert cfg.prefer_interval_join and cfg.interval_join_algorithm is sn.Algorithm.Coitrees and not cfg.interval_join_low_memory
    sn.apply_set(cfg, "SET sequila.interval_join_algorithm TO cuda")
    assert cfg.interval_join_algorithm is sn.Algorithm.Cuda
    sn.apply_set(cfg, "set sequila.prefer_interval_join to false;")
    sn.apply_set(cfg, "SET sequila.interval_join_low_memory = 'true'")
    assert not cfg.prefer_interval_join and cfg.interval_join_low_memory
    with pytest.raises(sn.ParseAlgorithmError):
        sn.apply_set(cfg, "SET sequila.interval_join_algorithm TO nothing")
    with pytest.raises(KeyError):
        sn.apply_set(cfg, "SET sequila.no_such_option TO 1")
    with pytest.raises(ValueError):
        sn.apply_set(cfg, "SET sequila.prefer_interval_join TO maybe")


SCH = pa.schema([("contig", pa.string()), ("pos_start", pa.int32()), ("pos_end", pa.int32())])
COLS = ["contig", "pos_start", "pos_end"]


def q1_filter():
    return IV.parse_condition_sql("reads.pos_start <= targets.pos_end AND reads.pos_end >= targets.pos_start",
                                  "reads", COLS, "targets", COLS)


def test_rule_and_explain_lines():
    cfg = sn.SequilaConfig()
    sn.apply_set(cfg, "SET sequila.interval_join_algorithm TO cuda")
    plan = optimize(HashJoinDesc(SCH, SCH, [("contig", "contig")], q1_filter()), cfg)
    assert isinstance(plan, IntervalJoinExec)
    # integration_test.rs:110 with alg = Cuda
    assert plan.display() == ("IntervalJoinExec: mode=CollectLeft, join_type=Inner, on=[(contig@0, contig@0)], "
                              "filter=pos_start@0 <= pos_end@3 AND pos_end@1 >= pos_start@2, alg=Cuda")
    # nested-loop (range only) join: on=[(1, 1)], integration_test.rs:204
    plan = optimize(HashJoinDesc(SCH, SCH, [], q1_filter()), cfg)
    assert plan.display() == ("IntervalJoinExec: mode=CollectLeft, join_type=Inner, on=[(1, 1)], "
                              "filter=pos_start@0 <= pos_end@3 AND pos_end@1 >= pos_start@2, alg=Cuda")
    assert plan.null_equals_null and plan.projection is None
    # rule skipped: switched off, or the filter is not an interval predicate
    off = sn.SequilaConfig(prefer_interval_join=False)
    d = HashJoinDesc(SCH, SCH, [("contig", "contig")], q1_filter())
    assert optimize(d, off) is d
    d2 = HashJoinDesc(SCH, SCH, [("contig", "contig")], None)
    assert optimize(d2, cfg) is d2


def test_try_new_validation():
    ci = IV.parse(q1_filter())
    with pytest.raises(PlanError, match="On constraints in HashJoinExec should be non-empty"):
        IntervalJoinExec.try_new(SCH, SCH, [], q1_filter(), ci)
    with pytest.raises(PlanError):
        IntervalJoinExec.try_new(SCH, SCH, [("contig", "contig")], q1_filter(), ci, projection=[0, 9])
    x = IntervalJoinExec.try_new(SCH, SCH, [("contig", "contig")], q1_filter(), ci, projection=[1, 4])
    assert x.schema().names == ["pos_start", "pos_start"]
    assert "projection=[pos_start@1, pos_start@4]" in x.display()
    assert IntervalJoinExec.maintains_input_order("Inner") == [False, True]
    # PartitionMode::Auto is rejected at execute (interval_join.rs:504-509)
    y = IntervalJoinExec.try_new(SCH, SCH, [("contig", "contig")], q1_filter(), ci, partition_mode=AUTO)
    with pytest.raises(PlanError, match="unsupported PartitionMode Auto"):
        y._open()


def test_cuda_keys_travel_with_the_session_config():
    """`SET sequila.cuda_* TO ..` lands in SequilaConfig like the reference's own keys (session_context.rs:50-60, 127-131)
    and reaches the exec nodes the rule creates; unknown keys are rejected with DataFusion's message"""
    import pyarrow as pa
    from sequila_native_b200 import intervals as IV
    from sequila_native_b200.interval_join import HashJoinDesc, optimize
    cfg = sn.SequilaConfig()
    sn.apply_set(cfg, "SET sequila.interval_join_algorithm TO cuda")
    sn.apply_set(cfg, "SET sequila.cuda_coalesce_rows TO 65536")
    sn.apply_set(cfg, "set sequila.cuda_probe_layout = 'soa'")
    assert cfg.cuda == {"cuda_coalesce_rows": "65536", "cuda_probe_layout": "soa"}
    with pytest.raises(KeyError) as e:
        sn.apply_set(cfg, "SET sequila.cuda_warp_speed TO 9")
    assert "not found on SequilaConfig" in str(e.value)
    cols = ["contig", "pos_start", "pos_end"]
    sch = pa.schema([("contig", pa.string()), ("pos_start", pa.int32()), ("pos_end", pa.int32())])
    f = IV.parse_condition_sql("a.pos_start <= b.pos_end AND a.pos_end >= b.pos_start", "a", cols, "b", cols)
    plan = optimize(HashJoinDesc(sch, sch, [("contig", "contig")], f), cfg)
    assert plan.cuda_options == cfg.cuda
