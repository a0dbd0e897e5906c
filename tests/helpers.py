"""Shared test helpers: turn the reference's test tables into the C-ABI argument arrays and back."""
import numpy as np

import sequila_native_b200 as sn


def encode_tables(left_rows, right_rows, equi=True):
    """rows = [contig, start, end].  Returns dict arrays for each side.  The key hash stands in for
    create_hashes(on): contig dictionary id when the join has the equi key, the constant of
    `on=[(lit(1), lit(1))]` (physical planner rule, sequila_physical_planner.rs:127-148) otherwise."""
    names = sorted({r[0] for r in left_rows} | {r[0] for r in right_rows})
    ids = {n: i for i, n in enumerate(names)}

    def side(rows):
        contig = np.array([ids[r[0]] for r in rows], dtype=np.int32)
        key = sn.synth.key_hash(contig) if equi else sn.synth.key_hash(np.ones(len(rows), dtype=np.int32))
        return {"contig": contig, "key": key, "start": np.array([r[1] for r in rows], dtype=np.int32),
                "end": np.array([r[2] for r in rows], dtype=np.int32), "rows": rows}

    return side(left_rows), side(right_rows)


def rows_from_pairs(left, right, l_idx, r_idx):
    out = [left["rows"][int(a)] + right["rows"][int(b)] for a, b in zip(l_idx, r_idx)]
    return sorted(out, key=lambda r: tuple(str(x) if isinstance(x, str) else x for x in r))


def sort_rows(rows):
    return sorted([list(r) for r in rows], key=lambda r: tuple(str(x) if isinstance(x, str) else x for x in r))


def canon(l_idx, r_idx):
    l = np.asarray(l_idx, dtype=np.uint64)
    r = np.asarray(r_idx, dtype=np.uint64)
    return np.sort((r << np.uint64(32)) | l)
