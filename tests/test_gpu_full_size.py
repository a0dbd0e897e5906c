"""GPU, BASELINE.json's full sizes, against the ORACLE: cfg2-cfg4 whole, and the bench workload (cfg5: 100M build rows,
one 12.5M-row probe launch) whole — pair count, order-independent pair digest (multi-threaded CPU pass of the coitrees
restatement), per-row counts, right_idx non-decreasing — for both CUDA implementations (fused kernel over packed lines,
count / scan / write over the SoA arrays)."""
import os

import numpy as np
import pytest
import torch

import sequila_native_b200 as sn

pytestmark = pytest.mark.gpu


def device_join(ctx, idx, p, packed):
    ctx.set_option("cuda_probe_layout", "packed" if packed else "soa")
    try:
        st = sn.CudaStream(ctx, cuda_stream=torch.cuda.current_stream().cuda_stream)
        n = st.probe_count_device(idx, p["key"], p["start"], p["end"])
        left = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
        right = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
        assert st.probe_join_device(idx, p["key"], p["start"], p["end"], left, right) == n
    finally:
        ctx.set_option("cuda_probe_layout", "auto")
    return st, n, left, right


def counts_of(st, n_rows):
    assert st.n_rows == n_rows
    return st.counts()


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
def test_full_config_matches_oracle_digest(cuda_ctx, oracle, name):
    b, p = sn.synth.CONFIGS[name]()
    dev = torch.device("cuda", 0)
    to = lambda s: {"key": torch.from_numpy(s["key"].view(np.int64)).to(dev), "start": torch.from_numpy(s["start"]).to(dev),
                    "end": torch.from_numpy(s["end"]).to(dev)}
    bd, pd = to(b), to(p)
    idx = sn.CudaIndex.build_device(cuda_ctx, bd["key"], bd["start"], bd["end"], torch.cuda.current_stream().cuda_stream)
    oidx = oracle.OracleIndex(b["key"], b["start"], b["end"])
    _, want_pairs, want_digest = oidx.time_probe(p["key"], p["start"], p["end"], threads=os.cpu_count() or 1, digest=True)
    want_counts = None
    for packed in (True, False):
        st, n, left, right = device_join(cuda_ctx, idx, pd, packed)
        assert n == want_pairs
        dg = st.digest_device(left, right, n)
        assert dg[0] == want_pairs and dg[1] == want_digest
        assert bool((right[1:n] >= right[:n - 1]).all())
        c = counts_of(st, len(p["key"]))
        assert int(c.sum(dtype=np.uint64)) == n
        if want_counts is None:
            want_counts = oidx.counts(p["key"][:200000], p["start"][:200000], p["end"][:200000])
        assert np.array_equal(c[:200000], want_counts)


def test_cfg5_shard_whole_launch_matches_oracle(cuda_ctx, oracle):
    """The bench workload at bench size — 100M build rows, the 12.5M-row probe shard of rank 0 — against the ORACLE:
    a coitrees restatement index over all 100M build rows (interval_join.rs:662-683), the whole shard probed through it
    on every host core (interval_join.rs:1582-1618); pair count, order-independent pair digest and all 12.5M per-row
    counts must equal the output of the fused packed-line kernel.  The SoA kernels must produce the same."""
    import bench

    class A:
        build_rows, shard_rows, workload, scaling, parallelism, probe_order = 100_000_000, 12_500_000, "cfg5_shard", "weak", "replicated", "random"
        total_probe_rows = 0
    dev = torch.device("cuda", 0)
    build, probe, _, _ = bench.make_workload(A, 0, 1, dev)
    idx = sn.CudaIndex.build_device(cuda_ctx, build["key"], build["start"], build["end"], torch.cuda.current_stream().cuda_stream)
    assert idx.uses_packed
    bh, ph = bench.to_host(build), bench.to_host(probe)
    # the host generator yields the same rows (what the reference arm of bench.py probes)
    chk = sn.synth.counter_side(1_000_000, bench.PROBE_SEED, first_row=5_000_000)
    assert np.array_equal(chk["start"], ph["start"][5_000_000:6_000_000]) and np.array_equal(chk["key"], ph["key"][5_000_000:6_000_000])
    oidx = oracle.OracleIndex(bh["key"], bh["start"], bh["end"])
    threads = os.cpu_count() or 1
    _, want_pairs, want_digest = oidx.time_probe(ph["key"], ph["start"], ph["end"], threads=threads, digest=True)
    import concurrent.futures as cf
    cuts = np.linspace(0, A.shard_rows, threads + 1).astype(np.int64)
    with cf.ThreadPoolExecutor(threads) as pool:  # ctypes releases the GIL inside the oracle
        parts = list(pool.map(lambda i: oidx.counts(ph["key"][cuts[i]:cuts[i + 1]], ph["start"][cuts[i]:cuts[i + 1]],
                                                    ph["end"][cuts[i]:cuts[i + 1]]), range(threads)))
    want_counts = np.concatenate(parts)
    assert int(want_counts.sum(dtype=np.uint64)) == want_pairs
    del oidx
    for packed in (True, False):
        st, n, left, right = device_join(cuda_ctx, idx, probe, packed)
        assert n == want_pairs
        dg = st.digest_device(left, right, n)
        assert dg[0] == want_pairs and dg[1] == want_digest
        assert bool((right[1:n] >= right[:n - 1]).all())
        assert np.array_equal(counts_of(st, A.shard_rows), want_counts)
        del left, right
    # the same rows position-sorted through the staged (TMA) kernel: same multiset of (left, original probe row) pairs
    order = torch.argsort(probe["contig"].to(torch.int64) * (1 << 32) + probe["start"].to(torch.int64))
    sp = {k: probe[k][order].contiguous() for k in ("key", "start", "end")}
    cuda_ctx.set_option("cuda_staged_probe", "on")
    try:
        st, n, left, right = device_join(cuda_ctx, idx, sp, True)
    finally:
        cuda_ctx.set_option("cuda_staged_probe", "auto")
    assert n == want_pairs
    assert bool((right[1:n] >= right[:n - 1]).all())
    assert np.array_equal(counts_of(st, A.shard_rows), want_counts[order.cpu().numpy()])
    orig = order.to(torch.int32)[right[:n].long()].contiguous()  # right_idx back in the unsorted numbering
    dg = st.digest_device(left, orig, n)
    assert dg[0] == want_pairs and dg[1] == want_digest


def test_two_rank_shards_on_one_gpu_cover_the_join(cuda_ctx, oracle):
    """The N>1 plans of sequila_native_b200.sharding driven through the CUDA path (ranks emulated one after the other on
    this GPU; the gloo test covers the rendezvous and the reductions): replicated index + probe shards, and contig-sharded
    (Partitioned) — the union of the ranks' pairs equals the oracle's whole join."""
    from sequila_native_b200 import sharding
    from helpers import canon
    b, p = sn.synth.cfg5(scale=0.002)
    ol, orr, _ = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    want = canon(ol, orr)
    world = 2
    # replicated
    idx = sn.CudaIndex.build(cuda_ctx, b["key"], b["start"], b["end"])
    ls, rs = [], []
    for rank in range(world):
        lo, hi = sharding.shard_bounds(len(p["key"]), world)[rank]
        st = sn.CudaStream(cuda_ctx)
        l, r, _ = st.probe(idx, p["key"][lo:hi], p["start"][lo:hi], p["end"][lo:hi])
        ls.append(l)
        rs.append(r.astype(np.int64) + lo)
    assert np.array_equal(canon(np.concatenate(ls), np.concatenate(rs)), want)
    assert np.all(np.diff(np.concatenate(rs)) >= 0)  # rank-order concatenation keeps probe order
    # contig-sharded
    plan = sharding.assign_keys_lpt(sn.synth.HG38, world)
    ls, rs = [], []
    for rank in range(world):
        brows = sharding.route_rows(b["contig"].astype(np.int64), plan)[rank]
        prows = sharding.route_rows(p["contig"].astype(np.int64), plan)[rank]
        idx_r = sn.CudaIndex.build(cuda_ctx, b["key"][brows], b["start"][brows], b["end"][brows])
        st = sn.CudaStream(cuda_ctx)
        l, r, _ = st.probe(idx_r, p["key"][prows], p["start"][prows], p["end"][prows])
        ls.append(brows[l])
        rs.append(prows[r])
    assert np.array_equal(canon(np.concatenate(ls), np.concatenate(rs)), want)
