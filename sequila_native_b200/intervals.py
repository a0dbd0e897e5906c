"""Join-filter parser: which columns are the build ("left") and probe ("right") interval bounds.

Mirror of the reference's ``intervals::parse`` (sequila/sequila-core/src/physical_planner/intervals.rs,
"IV"): the same parse table (IV:95-137), the same ``end - 1`` rewrite of strict comparisons
(IV:67-69), exactly two comparisons under one top-level ``AND`` (IV:212-231), and the same failure
modes — ``None`` when the filter is not an interval predicate (IV:30-38), :class:`PanicError` where the
reference panics (a slot set twice IV:158-183, two columns under one operand IV:53-55).

DataFusion's physical expressions are replaced by a minimal expression model (:class:`Column`,
:class:`Literal`, :class:`BinaryExpr`); :func:`parse_condition_sql` turns the SQL spellings the
reference's tests use (IV:258-505) into that model so the tests here read like the reference's.
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Optional, Union

LEFT, RIGHT = "Left", "Right"  # JoinSide


class PanicError(RuntimeError):
    """Situations in which the reference panics instead of returning an error."""


@dataclass(frozen=True)
class Column:
    name: str
    index: int            # index in the source (side) schema, as after map_column_to_source_schema
    side: Optional[str] = None  # join side; not part of equality (the reference compares name@index)

    def __eq__(self, other):
        return isinstance(other, Column) and (self.name, self.index) == (other.name, other.index)

    def __hash__(self):
        return hash((self.name, self.index))

    def __str__(self):
        return f"{self.name}@{self.index}"


@dataclass(frozen=True)
class Literal:
    value: int

    def __str__(self):
        return str(self.value)


@dataclass(frozen=True)
class BinaryExpr:
    left: "Expr"
    op: str  # one of < <= > >= - + AND OR
    right: "Expr"

    def __str__(self):
        return f"{self.left} {self.op} {self.right}"


Expr = Union[Column, Literal, BinaryExpr]


@dataclass(frozen=True)
class ColInterval:
    start: Expr
    end: Expr


@dataclass(frozen=True)
class ColIntervals:
    left_interval: ColInterval
    right_interval: ColInterval


def minus_one(e: Expr) -> Expr:
    return BinaryExpr(e, "-", Literal(1))  # IV:67-69


def _side_of(e: Expr) -> str:
    """IV:40-65: the operand must contain exactly one column; its side is the operand's side."""
    cols = []

    def walk(x):
        if isinstance(x, Column):
            cols.append(x)
        elif isinstance(x, BinaryExpr):
            walk(x.left)
            walk(x.right)

    walk(e)
    if len(cols) > 1:
        raise PanicError(f"complex sub queries are not supported {e}")
    if not cols:
        raise PanicError("side not found")
    return cols[0].side


class _Builder:
    def __init__(self):
        self.slots = {}

    def put(self, name: str, e: Expr):
        if name in self.slots:
            raise PanicError(f"{name} must not be called twice")  # IV:158-183
        self.slots[name] = e

    def finish(self) -> ColIntervals:
        for k in ("ls", "le", "rs", "re"):
            if k not in self.slots:
                raise PanicError(f"{k} must be set")  # IV:187-197
        s = self.slots
        return ColIntervals(ColInterval(s["ls"], s["le"]), ColInterval(s["rs"], s["re"]))


def _parse_condition(e: BinaryExpr, b: _Builder) -> None:
    """IV:71-138 — the four cases of the parse table."""
    op = e.op
    if op not in ("<", "<=", ">", ">="):
        raise ValueError(f"Unsupported operator: {op}")
    lt, gt = op in ("<", "<="), op in (">", ">=")
    strict = op in ("<", ">")
    lside = _side_of(e.left)
    if lside == RIGHT and lt:            # rs </<= le
        if _side_of(e.right) != LEFT:
            raise ValueError("couldn't parse as rs </<= le")
        b.put("rs", e.left)
        b.put("le", minus_one(e.right) if strict else e.right)
    elif lside == LEFT and lt:           # ls </<= re
        if _side_of(e.right) != RIGHT:
            raise ValueError("couldn't parse as ls </<= re")
        b.put("re", minus_one(e.right) if strict else e.right)
        b.put("ls", e.left)
    elif lside == RIGHT and gt:          # re >/>= ls
        if _side_of(e.right) != LEFT:
            raise ValueError("couldn't parse as re >/>= ls")
        b.put("re", minus_one(e.left) if strict else e.left)
        b.put("ls", e.right)
    elif lside == LEFT and gt:           # le >/>= rs
        if _side_of(e.right) != RIGHT:
            raise ValueError("couldn't parse as le >/>= rs")
        b.put("rs", e.right)
        b.put("le", minus_one(e.left) if strict else e.left)
    else:
        raise ValueError("couldn't parse left side")


def try_parse(filter_expr: Expr) -> ColIntervals:
    """IV:205-232"""
    if not isinstance(filter_expr, BinaryExpr):
        raise ValueError("couldn't cast filter to BinaryExpr")
    if filter_expr.op != "AND":
        raise ValueError("expr.op() is not AND")
    l, r = filter_expr.left, filter_expr.right
    if not isinstance(l, BinaryExpr):
        raise ValueError("couldn't cast left side to BinaryExpr")
    if not isinstance(r, BinaryExpr):
        raise ValueError("couldn't cast right side to BinaryExpr")
    b = _Builder()
    _parse_condition(l, b)
    _parse_condition(r, b)
    return b.finish()


def parse(filter_expr: Optional[Expr]) -> Optional[ColIntervals]:
    """IV:30-38: ``None`` (rule skipped) instead of an error; panics propagate."""
    if filter_expr is None:
        return None
    try:
        return try_parse(filter_expr)
    except ValueError:
        return None


# ---------------------------------------------------------------------------------------------
# SQL spelling -> expression model, for conditions of the form the reference tests use:
#   "b.r_end >= a.l_start AND a.l_end > b.r_start"
# ---------------------------------------------------------------------------------------------
_TOK = re.compile(r"\s*(?:(\d+)|([A-Za-z_]\w*)\.([A-Za-z_]\w*)|(<=|>=|<|>|=|\+|-|\(|\))|(AND|OR)\b)", re.I)


def parse_condition_sql(cond: str, left_alias: str, left_cols, right_alias: str, right_cols) -> Expr:
    """`left_cols` / `right_cols` are the column-name lists of the two source schemas (order = index)."""
    toks = []
    pos = 0
    cond = cond.strip()
    while pos < len(cond):
        m = _TOK.match(cond, pos)
        if not m:
            raise ValueError(f"cannot tokenize condition at {cond[pos:]!r}")
        pos = m.end()
        if m.group(1):
            toks.append(("num", int(m.group(1))))
        elif m.group(2):
            toks.append(("col", (m.group(2), m.group(3))))
        elif m.group(4):
            toks.append(("op", m.group(4)))
        else:
            toks.append(("bool", m.group(5).upper()))
    i = 0

    def peek():
        return toks[i] if i < len(toks) else (None, None)

    def take():
        nonlocal i
        t = toks[i]
        i += 1
        return t

    def atom():
        k, v = take()
        if k == "num":
            return Literal(v)
        if k == "col":
            alias, name = v
            if alias == left_alias:
                return Column(name, list(left_cols).index(name), LEFT)
            if alias == right_alias:
                return Column(name, list(right_cols).index(name), RIGHT)
            raise ValueError(f"unknown table alias {alias!r}")
        if (k, v) == ("op", "("):
            e = boolean()
            if take() != ("op", ")"):
                raise ValueError("expected )")
            return e
        raise ValueError(f"unexpected token {v!r}")

    def arith():
        e = atom()
        while peek() in (("op", "+"), ("op", "-")):
            _, o = take()
            e = BinaryExpr(e, o, atom())
        return e

    def comparison():
        e = arith()
        if peek()[0] == "op" and peek()[1] in ("<", "<=", ">", ">=", "="):
            _, o = take()
            e = BinaryExpr(e, o, arith())
        return e

    def boolean():
        e = comparison()
        while peek()[0] == "bool":
            _, o = take()
            e = BinaryExpr(e, o, comparison())
        return e

    out = boolean()
    if i != len(toks):
        raise ValueError("trailing tokens in condition")
    return out
