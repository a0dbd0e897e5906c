"""GPU: bounded output = `sequila.interval_join_low_memory` (interval_join.rs:1433-1530).  The tile is
joined once on the device and handed back in windows cut at probe-row boundaries; concatenating the
windows gives exactly the full-mode result."""
import numpy as np
import pyarrow as pa
import pytest

import sequila_native_b200 as sn
from sequila_native_b200 import intervals as IV
from sequila_native_b200.interval_join import HashJoinDesc, optimize
from helpers import canon

pytestmark = pytest.mark.gpu


def test_windows_of_the_pair_sequence(cuda_ctx, oracle):
    b, p = sn.synth.cfg3(scale=0.01)
    idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
    col = idx.add_column(b["end"])
    st = sn.CudaStream(cuda_ctx)
    n = st.probe_count(idx, p["key"], p["start"], p["end"])
    full_l, full_r, counts = st.emit_pairs()
    assert np.array_equal(st.counts(), counts)
    cap = 5000
    got_l, got_r, sizes = [], [], []
    off = cur = 0
    bounds = []
    for c in counts:                      # the reference's cut rule: whole probe rows, cap exceeded only by a lone row
        if cur + c > cap and cur > 0:
            bounds.append((off, cur)); off += cur; cur = 0
        cur += int(c)
    bounds.append((off, cur))
    for o, k in bounds:
        st.set_window(o, k)
        l, r = st.fetch_pairs(k)
        st.n_pairs = k                    # the gather below sizes its output from the window
        v = st.gather_build(col, np.int32)
        assert np.array_equal(v, b["end"][l])
        got_l.append(l); got_r.append(r); sizes.append(k)
    assert sum(sizes) == n and max(sizes) <= max(cap, counts.max())
    assert np.array_equal(np.concatenate(got_l), full_l) and np.array_equal(np.concatenate(got_r), full_r)
    with pytest.raises(sn.SequilaCudaError):
        st.set_window(n, 1)


def test_exec_low_memory_batches(oracle, monkeypatch):
    monkeypatch.setenv("SEQUILA_MAX_OUTPUT_BATCH_SIZE", "700")
    rng = np.random.default_rng(3)
    nb, npq = 3000, 1500
    bs = rng.integers(0, 20000, nb).astype(np.int32); be = (bs + rng.integers(0, 300, nb)).astype(np.int32)
    ps = rng.integers(0, 20000, npq).astype(np.int32); pe = (ps + rng.integers(0, 300, npq)).astype(np.int32)
    names = np.array(["chr1", "chr2"])
    bc, pc = rng.integers(0, 2, nb), rng.integers(0, 2, npq)
    cols = ["contig", "pos_start", "pos_end"]
    left = pa.record_batch([pa.array(names[bc]), pa.array(bs), pa.array(be)], names=cols)
    right = pa.record_batch([pa.array(names[pc]), pa.array(ps), pa.array(pe)], names=cols)
    cfg = sn.SequilaConfig()
    sn.apply_set(cfg, "SET sequila.interval_join_algorithm TO cuda")
    sn.apply_set(cfg, "SET sequila.interval_join_low_memory TO true")
    f = IV.parse_condition_sql("a.pos_start <= b.pos_end AND a.pos_end >= b.pos_start", "a", cols, "b", cols)
    plan = optimize(HashJoinDesc(left.schema, right.schema, [("contig", "contig")], f), cfg)
    assert plan.low_memory
    out = list(plan.execute([left], [right]))
    ol, orr, oc = oracle.join(bc.astype(np.uint64), bs, be, pc.astype(np.uint64), ps, pe)
    assert len(out) > 1 and sum(b.num_rows for b in out) == len(ol)
    assert all(b.num_rows <= max(700, int(oc.max())) for b in out)
    got = sorted(r for b in out for r in zip(*[c.to_pylist() for c in b.columns]))
    want = sorted(zip(names[bc][ol].tolist(), bs[ol].tolist(), be[ol].tolist(), names[pc][orr].tolist(), ps[orr].tolist(), pe[orr].tolist()))
    assert got == want
    m = plan.metrics()
    assert m.output_batches == len(out) and m.output_rows == len(ol) and m.input_batches == 1
