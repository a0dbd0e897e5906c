"""`IntervalJoinExec` with ``alg=Cuda`` over Arrow RecordBatches.

Python mirror of the reference's operator surface (sequila/sequila-core/src/physical_planner/joins/
interval_join.rs, "IJ"; sequila_physical_planner.rs, "PP") on top of the exec-node C ABI
(``include/sequila_exec.h``, implemented in C++ in ``csrc/sq_exec.cpp``):

* :meth:`IntervalJoinExec.try_new`   IJ:112-172 — same arguments, same validation order and messages
* :meth:`IntervalJoinExec.display`   IJ:323-365 — the ``EXPLAIN`` line, ``alg=Cuda``
* :meth:`IntervalJoinExec.execute`   IJ:449-557 + the stream state machine IJ:1054-1167: build side
  collected once (``PartitionMode::CollectLeft``), one output batch per probe batch, probe order kept
* :func:`optimize`                   PP:28-103 — the ``sequila.prefer_interval_join`` rule: a hash /
  nested-loop join whose filter parses as an interval predicate becomes an IntervalJoinExec

All join work (key hashing and i32 marshalling on the host in C++, index build / probe / take on the
GPU) happens below the C ABI; nothing here computes a join on the CPU and there is no fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

from . import _native as N
from .intervals import BinaryExpr, ColIntervals, Column, Expr, Literal, parse
from .session import Algorithm, SequilaConfig

COLLECT_LEFT, PARTITIONED, AUTO = "CollectLeft", "Partitioned", "Auto"


class PlanError(ValueError):
    """DataFusionError::Plan"""


class ExecutionError(RuntimeError):
    """DataFusionError::Execution / External"""


class _ArrowSchema(C.Structure):
    _fields_ = [("format", C.c_char_p), ("name", C.c_char_p), ("metadata", C.c_char_p), ("flags", C.c_int64),
                ("n_children", C.c_int64), ("children", C.c_void_p), ("dictionary", C.c_void_p),
                ("release", C.c_void_p), ("private_data", C.c_void_p)]


class _ArrowArray(C.Structure):
    _fields_ = [("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64), ("n_buffers", C.c_int64),
                ("n_children", C.c_int64), ("buffers", C.c_void_p), ("children", C.c_void_p),
                ("dictionary", C.c_void_p), ("release", C.c_void_p), ("private_data", C.c_void_p)]


_RELEASE_ARRAY = C.CFUNCTYPE(None, C.c_void_p)


def _col_and_minus_one(e: Expr, what: str) -> Tuple[int, bool]:
    """The interval expressions the parser produces are `col` or `col - 1` (intervals.rs:67-69)."""
    if isinstance(e, Column):
        return e.index, False
    if isinstance(e, BinaryExpr) and e.op == "-" and isinstance(e.left, Column) and e.right == Literal(1):
        return e.left.index, True
    raise PlanError(f"{what}: only `column` or `column - 1` interval expressions are supported, got {e}")


def format_filter(e: Expr) -> str:
    """A JoinFilter prints its columns with their index in the filter's own intermediate schema: the
    build-side columns the filter uses (in source order), then the probe-side ones — e.g.
    ``pos_start@0 <= pos_end@3 AND pos_end@1 >= pos_start@2`` (the reference's integration_test.rs:110)."""
    used = {"Left": set(), "Right": set()}

    def walk(x):
        if isinstance(x, Column):
            used[x.side].add(x.index)
        elif isinstance(x, BinaryExpr):
            walk(x.left)
            walk(x.right)

    walk(e)
    order = [("Left", i) for i in sorted(used["Left"])] + [("Right", i) for i in sorted(used["Right"])]
    pos = {k: n for n, k in enumerate(order)}

    def fmt(x):
        if isinstance(x, Column):
            return f"{x.name}@{pos[(x.side, x.index)]}"
        if isinstance(x, BinaryExpr):
            return f"{fmt(x.left)} {x.op} {fmt(x.right)}"
        return str(x)

    return fmt(e)


@dataclass
class JoinMetrics:
    """BuildProbeJoinMetrics (the reference's utils.rs:441-495)"""
    build_input_batches: int = 0
    build_input_rows: int = 0
    build_mem_used: int = 0
    input_batches: int = 0
    input_rows: int = 0
    output_batches: int = 0
    output_rows: int = 0
    build_time_ns: int = 0
    join_time_ns: int = 0
    index_bytes: int = 0
    keys: int = 0


class IntervalJoinExec:
    def __init__(self):
        raise TypeError("use IntervalJoinExec.try_new")

    @classmethod
    def try_new(cls, left_schema, right_schema, on: Sequence[Tuple[object, object]], filter: Optional[Expr],
                intervals: ColIntervals, join_type: str = "Inner", projection: Optional[Sequence[int]] = None,
                partition_mode: str = COLLECT_LEFT, null_equals_null: bool = False,
                algorithm: Algorithm = Algorithm.Cuda, low_memory: bool = False, device: int = 0) -> "IntervalJoinExec":
        """`on` = [(left column, right column)] as names, indices or :class:`Column`; the constant pair
        ``[(1, 1)]`` given as ``[(Literal(1), Literal(1))]`` marks a range-only join (PP:127-148)."""
        self = object.__new__(cls)
        if not on:
            raise PlanError("On constraints in HashJoinExec should be non-empty")  # IJ:128-130
        self.left_schema, self.right_schema = left_schema, right_schema
        self.on_display = []
        self.on_left, self.on_right = [], []
        for l, r in on:
            if isinstance(l, Literal) and isinstance(r, Literal):
                self.on_display.append((str(l.value), str(r.value)))
                continue
            li, ri = self._resolve(left_schema, l, "left"), self._resolve(right_schema, r, "right")
            self.on_left.append(li)
            self.on_right.append(ri)
            self.on_display.append((f"{left_schema.field(li).name}@{li}", f"{right_schema.field(ri).name}@{ri}"))
        if join_type != "Inner":
            # the reference stores join_type but emits inner-join rows regardless (IJ:1064, 1580-1640)
            raise PlanError(f"IntervalJoinExec emits Inner joins only, got {join_type}")
        self.filter = filter
        self.intervals = intervals
        self.join_type = join_type
        n_l, n_r = len(left_schema), len(right_schema)
        # build_join_schema(left, right, Inner): left columns then right columns (IJ:133-134)
        self.join_fields = [left_schema.field(i) for i in range(n_l)] + [right_schema.field(i) for i in range(n_r)]
        if projection is not None:
            for p in projection:  # can_project (IJ:141)
                if not (0 <= p < n_l + n_r):
                    raise PlanError(f"project index {p} out of bounds, max field {n_l + n_r}")
        self.projection = list(projection) if projection is not None else None
        self.mode = partition_mode
        self.null_equals_null = null_equals_null
        self.algorithm = algorithm
        self.low_memory = low_memory
        # the reference caps a low-memory output batch at 1M rows, overridable with the environment variable
        # SEQUILA_MAX_OUTPUT_BATCH_SIZE of the HOST process (interval_join.rs:551-555, 1439): same name, same place — the
        # exec-node mirror; the library itself reads no environment
        import os
        self.max_output_rows = int(os.environ.get("SEQUILA_MAX_OUTPUT_BATCH_SIZE", "1000000"))
        self.device = device
        self.cuda_options = {}  # sequila.cuda_* keys of the session, applied when the node opens
        self._exec = None
        self._lib = None
        ls, lm = _col_and_minus_one(intervals.left_interval.start, "left start")
        le, lem = _col_and_minus_one(intervals.left_interval.end, "left end")
        rs, rm = _col_and_minus_one(intervals.right_interval.start, "right start")
        re_, rem = _col_and_minus_one(intervals.right_interval.end, "right end")
        if lm or rm:
            raise PlanError("`start - 1` is not an expression the interval parser produces")
        self._cols = (ls, le, lem, rs, re_, rem)
        return self

    @staticmethod
    def _resolve(schema, c, which: str) -> int:
        if isinstance(c, Column):
            return c.index
        if isinstance(c, int):
            return c
        i = schema.get_field_index(c)
        if i < 0:
            raise PlanError(f"The {which} or right side of the join does not have all columns on \"on\": missing {c}")
        return i

    # ---- planning surface ---------------------------------------------------------------------------
    def schema(self):
        import pyarrow as pa
        f = self.join_fields if self.projection is None else [self.join_fields[i] for i in self.projection]
        return pa.schema(f)

    @staticmethod
    def maintains_input_order(join_type: str = "Inner") -> List[bool]:
        return [False, join_type in ("Inner", "RightAnti", "RightSemi")]  # IJ:210-218

    def display(self) -> str:
        """IJ:323-365 (DisplayFormatType::Default)"""
        flt = f", filter={format_filter(self.filter)}" if self.filter is not None else ""
        proj = ""
        if self.projection is not None:
            proj = ", projection=[" + ", ".join(f"{self.join_fields[i].name}@{i}" for i in self.projection) + "]"
        on = ", ".join(f"({a}, {b})" for a, b in self.on_display)
        return f"IntervalJoinExec: mode={self.mode}, join_type={self.join_type}, on=[{on}]{flt}{proj}, alg={self.algorithm}"

    # ---- execution -----------------------------------------------------------------------------------
    def _err(self) -> str:
        return (self._lib.sq_exec_last_error(self._exec) or b"").decode()

    def _open(self):
        if self.algorithm not in (Algorithm.Cuda, Algorithm.CudaNearest):
            raise ExecutionError(f"algorithm {self.algorithm} is the reference's CPU code; this build executes alg=Cuda / CudaNearest only")
        if self.mode == AUTO:
            # IJ:504-509
            raise PlanError("Invalid IntervalJoinExec, unsupported PartitionMode Auto in execute()")
        lib = self._lib = N.lib()
        ls, le, lem, rs, re_, rem = self._cols
        n_on = len(self.on_left)
        onl = (C.c_int32 * max(n_on, 1))(*self.on_left)
        onr = (C.c_int32 * max(n_on, 1))(*self.on_right)
        proj = self.projection
        parr = (C.c_int32 * max(len(proj) if proj is not None else 1, 1))(*(proj or []))
        cfg = N.SqExecConfig(self.device, n_on, onl, onr, ls, le, rs, re_, int(lem), int(rem),
                             -1 if proj is None else len(proj), parr, 1 if self.algorithm == Algorithm.CudaNearest else 0,
                             int(self.low_memory), int(self.max_output_rows))
        lsch, rsch = _ArrowSchema(), _ArrowSchema()
        self.left_schema._export_to_c(C.addressof(lsch))
        self.right_schema._export_to_c(C.addressof(rsch))
        h = C.c_void_p()
        rc = lib.sq_exec_create(C.byref(cfg), C.addressof(lsch), C.addressof(rsch), C.byref(h))
        for sch in (lsch, rsch):
            if sch.release:
                _RELEASE_ARRAY(sch.release)(C.addressof(sch))
        if rc != N.SQ_OK:
            msg = lib.sq_exec_last_error(h).decode() if h.value else lib.sq_last_error(None).decode()
            if h.value:
                lib.sq_exec_free(h)
            raise ExecutionError(msg)
        self._exec = h
        for k, v in self.cuda_options.items():
            if lib.sq_exec_set_option(h, k.encode(), str(v).encode()) != N.SQ_OK:
                raise ExecutionError(self._err())

    def close(self):
        if self._exec is not None:
            self._lib.sq_exec_free(self._exec)
            self._exec = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def collect_build(self, build_batches: Iterable) -> None:
        """collect_left_input (IJ:597-689): fold the build batches, build the index once."""
        if self._exec is None:
            self._open()
        lib = self._lib
        for b in build_batches:
            arr = _ArrowArray()
            b._export_to_c(C.addressof(arr))
            rc = lib.sq_exec_push_build(self._exec, C.addressof(arr))
            if rc != N.SQ_OK:
                if arr.release:
                    _RELEASE_ARRAY(arr.release)(C.addressof(arr))
                raise ExecutionError(self._err())
        if lib.sq_exec_finish_build(self._exec) != N.SQ_OK:
            raise ExecutionError(self._err())

    def probe_batch(self, batch, partition: int = 0):
        """process_probe_batch, full mode (IJ:1580-1640): one output RecordBatch per probe batch."""
        import pyarrow as pa
        lib = self._lib
        arr, out = _ArrowArray(), _ArrowArray()
        batch._export_to_c(C.addressof(arr))
        try:
            rc = lib.sq_exec_probe(self._exec, partition, C.addressof(arr), C.addressof(out))
        finally:
            if arr.release:  # borrowed for the call only
                _RELEASE_ARRAY(arr.release)(C.addressof(arr))
        if rc != N.SQ_OK:
            raise ExecutionError(self._err())
        return pa.RecordBatch._import_from_c(C.addressof(out), self.schema())

    def probe_batches(self, batches: Iterable, partition: int = 0) -> Iterator:
        """The probe side of a partition as a stream of batches -> a stream of output batches (full mode).  Batches are
        coalesced into tiles of `sequila.cuda_coalesce_rows` rows (sq_exec_probe_push / _pop): the reference's child yields
        <= 8192-row batches (IJ:1192-1233), far too small for a launch chain + PCIe round trip each; row order is kept."""
        import pyarrow as pa
        lib = self._lib

        def pop(flush):
            out = _ArrowArray()
            has = C.c_int32(0)
            if lib.sq_exec_probe_pop(self._exec, partition, int(flush), C.addressof(out), C.byref(has)) != N.SQ_OK:
                raise ExecutionError(self._err())
            return pa.RecordBatch._import_from_c(C.addressof(out), self.schema()) if has.value else None
        for b in batches:
            arr = _ArrowArray()
            b._export_to_c(C.addressof(arr))
            ready = C.c_int32(0)
            if lib.sq_exec_probe_push(self._exec, partition, C.addressof(arr), C.byref(ready)) != N.SQ_OK:
                if arr.release:
                    _RELEASE_ARRAY(arr.release)(C.addressof(arr))
                raise ExecutionError(self._err())
            if ready.value:
                out = pop(False)
                if out is not None:
                    yield out
        while True:  # the tile in flight, then whatever is still waiting (tiles are pipelined two deep)
            out = pop(True)
            if out is None:
                break
            yield out

    def set_option(self, key: str, value) -> None:
        """`SET sequila.cuda_<name> TO <value>` for this node (sq_exec_set_option)."""
        if self._exec is None:
            self._open()
        if self._lib.sq_exec_set_option(self._exec, str(key).encode(), str(value).encode()) != N.SQ_OK:
            raise ExecutionError(self._err())

    def probe_batch_low_memory(self, batch, partition: int = 0) -> Iterator:
        """process_probe_batch, low-memory mode (IJ:1433-1530): output batches of at most
        `max_output_rows` rows, cut at probe-row boundaries."""
        import pyarrow as pa
        lib = self._lib
        arr = _ArrowArray()
        batch._export_to_c(C.addressof(arr))
        try:
            if lib.sq_exec_probe_begin(self._exec, partition, C.addressof(arr)) != N.SQ_OK:
                raise ExecutionError(self._err())
            more = C.c_int32(1)
            while more.value:
                out = _ArrowArray()
                if lib.sq_exec_probe_next(self._exec, partition, C.addressof(out), C.byref(more)) != N.SQ_OK:
                    raise ExecutionError(self._err())
                yield pa.RecordBatch._import_from_c(C.addressof(out), self.schema())
        finally:
            if arr.release:
                _RELEASE_ARRAY(arr.release)(C.addressof(arr))

    def execute(self, build_batches: Iterable, probe_batches: Iterable, partition: int = 0, coalesce: bool = True) -> Iterator:
        """IJ:449-557: await the build side, then stream the probe side.  Full mode coalesces the probe batches into tiles
        of `sequila.cuda_coalesce_rows` rows (one output batch per tile); coalesce=False joins every probe batch on its
        own, one output batch each, as the reference does (IJ:1580-1640)."""
        self.collect_build(build_batches)
        if not self.low_memory and coalesce:
            yield from self.probe_batches(probe_batches, partition)
            return
        for b in probe_batches:
            if self.low_memory:
                yield from self.probe_batch_low_memory(b, partition)
            else:
                yield self.probe_batch(b, partition)

    def metrics(self) -> JoinMetrics:
        out = (C.c_uint64 * 16)()
        if self._exec is None or self._lib.sq_exec_metrics(self._exec, out) != N.SQ_OK:
            return JoinMetrics()
        return JoinMetrics(*[int(out[i]) for i in range(11)])


# ---------------------------------------------------------------------------------------------
@dataclass
class HashJoinDesc:
    """What the rule reads off a HashJoinExec / NestedLoopJoinExec (PP:42-92)."""
    left_schema: object
    right_schema: object
    on: Sequence[Tuple[object, object]]   # empty for a NestedLoopJoinExec
    filter: Optional[Expr]
    join_type: str = "Inner"
    projection: Optional[Sequence[int]] = None
    partition_mode: str = COLLECT_LEFT
    null_equals_null: bool = False


def optimize(join: HashJoinDesc, config: SequilaConfig, device: int = 0):
    """IntervalJoinPhysicalOptimizationRule::optimize (PP:28-103).  Returns an IntervalJoinExec when the
    rule fires, else the unchanged join description."""
    if not config.prefer_interval_join:
        return join  # PP:36-39
    intervals = parse(join.filter)
    if intervals is None:
        return join  # "Could not build range filter", PP:55-60
    if join.on:      # from_hash_join, PP:105-125
        node = IntervalJoinExec.try_new(join.left_schema, join.right_schema, join.on, join.filter, intervals,
                                        join.join_type, join.projection, join.partition_mode, join.null_equals_null,
                                        config.interval_join_algorithm, config.interval_join_low_memory, device)
    else:
        # from_nested_loop_join, PP:127-148: on = [(lit(1), lit(1))], no projection, CollectLeft, null_equals_null
        node = IntervalJoinExec.try_new(join.left_schema, join.right_schema, [(Literal(1), Literal(1))], join.filter,
                                        intervals, join.join_type, None, COLLECT_LEFT, True,
                                        config.interval_join_algorithm, config.interval_join_low_memory, device)
    node.cuda_options = dict(getattr(config, "cuda", {}))
    return node
