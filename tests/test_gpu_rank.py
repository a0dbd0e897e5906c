"""GPU: the rank-difference probe kernel (sq_probe_rank.cu) against the oracle and against the count / scan / write
kernels — configs 2-4, inverted probe rows (counted by the walk), an inverted BUILD row (no rank structure: the old
chain serves the index), extreme coordinates, candidate lists around the 32-candidate flattening limit, count-only."""
import numpy as np
import pytest
import torch

import sequila_native_b200 as sn
from helpers import canon

pytestmark = pytest.mark.gpu


@pytest.fixture()
def soa(cuda_ctx):
    cuda_ctx.set_option("cuda_probe_layout", "soa")
    cuda_ctx.set_option("cuda_rank_count", "force")
    yield cuda_ctx
    cuda_ctx.set_option("cuda_probe_layout", "auto")
    cuda_ctx.set_option("cuda_rank_count", "on")


def check(oracle, ctx, b, p, want_rank=True):
    idx = sn.CudaIndex.build(ctx, b["key"], b["start"], b["end"])
    assert idx.uses_rank == want_rank and not idx.uses_packed
    st = sn.CudaStream(ctx)
    n = st.probe_count(idx, p["key"], p["start"], p["end"])
    l, r, c = st.emit_pairs()
    ol, orr, oc = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    assert n == len(ol) and np.array_equal(c, oc)
    assert np.all(np.diff(r.astype(np.int64)) >= 0)
    assert np.array_equal(canon(l, r), canon(ol, orr))
    return idx


@pytest.mark.parametrize("name,scale", [("cfg2", 0.1), ("cfg3", 0.02), ("cfg4", 0.05), ("cfg5", 0.002)])
def test_configs(soa, oracle, name, scale):
    b, p = sn.synth.CONFIGS[name](scale=scale)
    check(oracle, soa, b, p)


def test_inverted_probe_rows_and_an_inverted_build_row(soa, oracle):
    rng = np.random.default_rng(4)
    nb, npq = 40_000, 30_000
    bs = rng.integers(0, 1_000_000, nb).astype(np.int32)
    b = {"key": sn.synth.key_hash(rng.integers(0, 3, nb)), "start": bs, "end": (bs + rng.integers(0, 5000, nb)).astype(np.int32)}
    ps = rng.integers(-1000, 1_001_000, npq).astype(np.int32)
    pe = (ps + rng.integers(-3000, 4000, npq)).astype(np.int32)  # ~40 % of the probe rows inverted (end < start)
    p = {"key": sn.synth.key_hash(rng.integers(0, 4, npq)), "start": ps, "end": pe}
    check(oracle, soa, b, p)
    b2 = {k: v.copy() for k, v in b.items()}
    b2["end"][777] = b2["start"][777] - 5  # one inverted build row: the rank identity no longer holds, no rank structure
    check(oracle, soa, b2, p, want_rank=False)


def test_candidate_lists_around_the_flattening_limit_and_extremes(soa, oracle):
    rng = np.random.default_rng(8)
    # every probe row meets exactly k candidates, k = 0..70: nested build intervals around one point per key
    keys, starts, ends = [], [], []
    for k in range(71):
        for d in range(k):
            keys.append(k); starts.append(1000 - d); ends.append(1000 + d)
    b = {"key": sn.synth.key_hash(np.array(keys)), "start": np.array(starts, np.int32), "end": np.array(ends, np.int32)}
    pk = np.repeat(np.arange(72), 5)
    p = {"key": sn.synth.key_hash(pk), "start": np.full(len(pk), 1000, np.int32) + rng.integers(-3, 4, len(pk)).astype(np.int32),
         "end": np.full(len(pk), 1000, np.int32) + rng.integers(0, 40, len(pk)).astype(np.int32)}
    check(oracle, soa, b, p)
    lo, hi = -2_147_483_647, 2_147_483_646
    b = {"key": sn.synth.key_hash(np.zeros(6)), "start": np.array([lo, lo, 0, hi - 1, hi, -5], np.int32),
         "end": np.array([lo, hi, 0, hi, hi, 5], np.int32)}
    p = {"key": sn.synth.key_hash(np.zeros(5)), "start": np.array([lo, hi, -1, lo, 6], np.int32), "end": np.array([lo, hi, 1, hi, hi], np.int32)}
    check(oracle, soa, b, p)


def test_rank_kernel_equals_the_two_walk_chain_on_the_device(cuda_ctx, oracle):
    """cfg3 at 1/5 scale on the device: same counts, same pair digest, count-only launch touches no candidate"""
    b, p = sn.synth.cfg3(scale=0.2)
    dev = torch.device("cuda", 0)
    to = lambda s: {k: torch.from_numpy(s[k].view(np.int64) if k == "key" else s[k]).to(dev) for k in ("key", "start", "end")}
    bd, pd = to(b), to(p)
    ts = torch.cuda.current_stream().cuda_stream
    res = []
    for rank in ("force", "off"):
        cuda_ctx.set_option("cuda_rank_count", rank)
        idx = sn.CudaIndex.build_device(cuda_ctx, bd["key"], bd["start"], bd["end"], ts)
        assert idx.uses_rank == (rank == "force")
        st = sn.CudaStream(cuda_ctx, cuda_stream=ts)
        n = st.probe_count_device(idx, pd["key"], pd["start"], pd["end"])
        counts = st.counts()
        left = torch.empty(n, dtype=torch.int32, device=dev)
        right = torch.empty(n, dtype=torch.int32, device=dev)
        assert st.probe_join_device(idx, pd["key"], pd["start"], pd["end"], left, right) == n
        assert bool((right[1:] >= right[:-1]).all())
        res.append((n, st.digest_device(left, right, n), counts, st.counts()))
    cuda_ctx.set_option("cuda_rank_count", "on")
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1]
    assert np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][3], res[1][3])
    oidx = oracle.OracleIndex(b["key"], b["start"], b["end"])
    assert np.array_equal(res[0][2][:100000], oidx.counts(p["key"][:100000], p["start"][:100000], p["end"][:100000]))


def test_depth_heuristic_picks_the_kernel(cuda_ctx):
    """option on (the default): deep indexes (cfg4: ~100 rows reach every start) get the rank kernel, shallow ones keep the
    count / scan / write chain (measured crossover, sq_probe_rank.cu::use_rank)"""
    cuda_ctx.set_option("cuda_rank_count", "on")
    b4, _ = sn.synth.cfg4(scale=0.05)
    b2, _ = sn.synth.cfg2(scale=0.05)
    assert sn.CudaIndex.build(cuda_ctx, b4["key"], b4["start"], b4["end"]).uses_rank
    assert not sn.CudaIndex.build(cuda_ctx, b2["key"], b2["start"], b2["end"]).uses_rank
