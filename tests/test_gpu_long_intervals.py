"""GPU: a few chromosome-long build intervals among millions of short ones (a GFF `region` row, a large SV).  Behind such
a row the running max end stays high for the rest of its key segment, so candidate ranges cover O(segment) rows; the
walks descend the block-max pyramid (32 / 1024 / 32768 rows) and touch only the blocks that hold a hit, as coitrees prunes
through subtree_last (CT/nosimd.rs:343-384).  Parity with the oracle AND a time bound: without the pyramid these cases
run for minutes."""
import time

import numpy as np
import pytest

import sequila_native_b200 as sn
from sequila_native_b200 import _native as N
from helpers import canon

pytestmark = pytest.mark.gpu


def workload(nb=2_000_000, npq=200_000, span=100_000_000, keys=2, seed=3):
    rng = np.random.default_rng(seed)
    k = rng.integers(0, keys, nb)
    s = rng.integers(0, span, nb).astype(np.int64)
    e = s + rng.integers(50, 150, nb)
    # per key: one interval spanning everything, one spanning the second half, one 5 Mbp interval
    for key in range(keys):
        rows = np.flatnonzero(k == key)[:3]
        s[rows] = [0, span // 2, span // 4]
        e[rows] = [span, span, span // 4 + 5_000_000]
    b = {"key": sn.synth.key_hash(k), "start": s.astype(np.int32), "end": e.astype(np.int32)}
    pk = rng.integers(0, keys + 1, npq)  # one key absent from the build side
    ps = rng.integers(0, span, npq).astype(np.int64)
    p = {"key": sn.synth.key_hash(pk), "start": ps.astype(np.int32), "end": (ps + rng.integers(0, 300, npq)).astype(np.int32)}
    return b, p


@pytest.mark.parametrize("rank", ["force", "off"])
def test_join_behind_chromosome_long_intervals(cuda_ctx, oracle, rank):
    b, p = workload()
    cuda_ctx.set_option("cuda_rank_count", rank)
    try:
        idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
        assert not idx.uses_packed and idx.uses_rank == (rank == "force")  # a width >= 65536: the SoA kernels serve it
        st = sn.CudaStream(cuda_ctx)
        st.probe(idx, p["key"][:1000], p["start"][:1000], p["end"][:1000])  # warm-up (allocations)
        t0 = time.perf_counter()
        n = st.probe_count(idx, p["key"], p["start"], p["end"])
        l, r, c = st.emit_pairs()
        dt = time.perf_counter() - t0
    finally:
        cuda_ctx.set_option("cuda_rank_count", "on")
    ol, orr, oc = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    assert n == len(ol) and np.array_equal(c, oc)
    assert np.all(np.diff(r.astype(np.int64)) >= 0)
    assert np.array_equal(canon(l, r), canon(ol, orr))
    assert dt < 2.0, f"{dt:.2f} s for 200k probe rows: the long candidate ranges are being walked row by row"


def test_nearest_behind_a_chromosome_long_interval(cuda_ctx, oracle):
    b, p = workload(nb=1_000_000, npq=100_000, keys=1)
    idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
    st = sn.CudaStream(cuda_ctx)
    st.probe_nearest(idx, p["key"][:100], p["start"][:100], p["end"][:100])
    t0 = time.perf_counter()
    got = st.probe_nearest(idx, p["key"], p["start"], p["end"])
    dt = time.perf_counter() - t0
    want, is_overlap = oracle.nearest(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    # every probe row of a present key overlaps the spanning interval: the answer must be SOME overlapping row
    present = want != oracle.NULL_INDEX
    assert np.array_equal(got == N.NULL_INDEX, ~present)
    g = got[present].astype(np.int64)
    assert np.all((b["start"][g] <= p["end"][present]) & (b["end"][g] >= p["start"][present]))
    assert is_overlap[present].all()
    assert dt < 2.0, f"{dt:.2f} s: the walk back to the long interval is linear"
