#!/usr/bin/env python
"""Extract the golden vectors the reference's OWN tests hold for the interval-join path and write
them to tests/golden/reference_tables.json.  Run in the build container only (it reads
/root/reference, which does not exist on the GPU box); the JSON it writes is committed.

Sources (reference file:line):
  testing/data/interval/reads.csv, targets.csv                 the only shipped fixtures
  sequila/sequila-core/tests/integration_test.rs:42-65         16-row equi+range result
  sequila/sequila-core/tests/integration_test.rs:122-161       32-row range-only result
  sequila/sequila-core/tests/integration_test.rs:216-291       closed boundaries: 12 b-rows -> 10 rows
  sequila/sequila-core/tests/integration_test.rs:295-350       strict boundaries -> 6 rows
  sequila/sequila-core/src/physical_planner/joins/interval_join.rs:1959-1965  cast-overflow text
  sequila/sequila-core/tests/integration_test.rs:352-399       CoitreesNearest: 1 a-row, 4 b-rows -> 4 rows
"""
import csv
import json
import os
import re

REF = "/root/reference"
IT = os.path.join(REF, "sequila/sequila-core/tests/integration_test.rs")
IJ = os.path.join(REF, "sequila/sequila-core/src/physical_planner/joins/interval_join.rs")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_tables.json")


def read_csv(name):
    with open(os.path.join(REF, "testing/data/interval", name)) as f:
        rows = list(csv.reader(f))
    assert rows[0] == ["contig", "pos_start", "pos_end"]
    return [[r[0], int(r[1]), int(r[2])] for r in rows[1:] if r]


def table_rows(text):
    """rows of an `assert_batches_sorted_eq!` ASCII table: lines `"| a | 1 | ... |",` minus the header"""
    rows = []
    for line in re.findall(r'"\|([^"]*)\|"', text):
        cells = [c.strip() for c in line.split("|")]
        rows.append(cells)
    body = [r for r in rows[1:]]  # first row is the column header
    return [[int(c) if re.fullmatch(r"-?\d+", c) else c for c in r] for r in body]


def fn_body(src, name):
    i = src.index(name)
    j = src.find("\n}\n", i)
    return src[i:j]


def values_rows(sql_block):
    return [[m[0], int(m[1]), int(m[2])] for m in re.findall(r"\('(\w+)',\s*(-?\d+),\s*(-?\d+)\)", sql_block)]


def main():
    it = open(IT).read()
    ij = open(IJ).read()
    gteq = fn_body(it, "async fn test_all_gteq_lteq_conditions")
    gtlt = fn_body(it, "async fn test_all_gt_lt_conditions")
    a_sql = gteq[gteq.index("let a ="):gteq.index("let b =")]
    b_sql = gteq[gteq.index("let b ="):gteq.index("let q0")]
    m = re.search(r'"(Arrow error: Cast error: Can\'t cast value \{\} to type Int32)"', ij)
    near = fn_body(it, "async fn test_nearest")
    near_a = near[near.index("let a ="):near.index("let b =")]
    near_b = near[near.index("let b ="):near.index("ctx.sql(\"SET")]
    quad = lambda block: [[m[0], m[1], int(m[2]), int(m[3])]
                          for m in re.findall(r"\('(\w+)',\s*'(\w+)',\s*(-?\d+),\s*(-?\d+)\)", block)]
    near_rows = []
    for line in re.findall(r'"\|([^"]*)\|"', near[near.index("let expected"):])[1:]:
        cells = [c.strip() for c in line.split("|")]
        near_rows.append([None if c == "" else (int(c) if re.fullmatch(r"-?\d+", c) else c) for c in cells])
    golden = {
        "_generated_by": "tests/golden/make_golden.py from /root/reference (see docstring for file:line)",
        "reads": read_csv("reads.csv"),
        # the fixture files byte for byte (scan-side parity: the device scanner and its oracle parse THESE bytes)
        "reads_csv_text": open(os.path.join(REF, "testing/data/interval/reads.csv"), "rb").read().decode("ascii"),
        "targets_csv_text": open(os.path.join(REF, "testing/data/interval/targets.csv"), "rb").read().decode("ascii"),
        "targets": read_csv("targets.csv"),
        "equi_rows": table_rows(fn_body(it, "fn expected_equi()")),
        "range_rows": table_rows(fn_body(it, "fn expected_range()")),
        "boundary_a": values_rows(a_sql),
        "boundary_b": values_rows(b_sql),
        "closed_rows": table_rows(gteq[gteq.index("let expected"):]),
        "strict_rows": table_rows(gtlt[gtlt.index("let expected"):]),
        "cast_error_format": m.group(1),
        "cast_error_value": 2 ** 31,
        "nearest_a": quad(near_a),
        "nearest_b": quad(near_b),
        "nearest_rows": near_rows,
    }
    assert len(golden["nearest_a"]) == 1 and len(golden["nearest_b"]) == 4 and len(golden["nearest_rows"]) == 4
    assert len(golden["reads"]) == 12 and len(golden["targets"]) == 10
    assert len(golden["equi_rows"]) == 16 and len(golden["range_rows"]) == 32
    assert len(golden["boundary_a"]) == 1 and len(golden["boundary_b"]) == 12
    assert len(golden["closed_rows"]) == 10 and len(golden["strict_rows"]) == 6
    with open(OUT, "w") as f:
        json.dump(golden, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
