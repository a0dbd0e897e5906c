"""GPU, BASELINE.json's full sizes: parity through properties that do not need a CPU pass over the whole
input — (i) the two independent CUDA implementations (fused kernel over packed lines, count/scan/write
over the SoA arrays) emit the same pair multiset (order-independent digest) and the same per-row counts,
(ii) sum(counts) == n_pairs and right_idx is non-decreasing, (iii) a sampled slice of probe rows equals
the oracle's answer row by row.  cfg2-cfg4 are additionally checked against the oracle's digest of the
WHOLE join (multi-threaded CPU pass, seconds)."""
import os

import numpy as np
import pytest
import torch

import sequila_native_b200 as sn

pytestmark = pytest.mark.gpu


def device_join(ctx, idx, p, packed):
    os.environ["SQ_PACKED"] = "1" if packed else "0"
    try:
        st = sn.CudaStream(ctx, cuda_stream=torch.cuda.current_stream().cuda_stream)
        n = st.probe_count_device(idx, p["key"], p["start"], p["end"])
        left = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
        right = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
        assert st.probe_join_device(idx, p["key"], p["start"], p["end"], left, right) == n
    finally:
        os.environ.pop("SQ_PACKED", None)
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


def test_cfg5_shard_two_implementations_agree(cuda_ctx, oracle):
    """100M build rows, 12.5M probe rows (the bench workload): packed-line kernel vs SoA kernels."""
    import bench

    class A:
        build_rows, shard_rows, workload = 100_000_000, 12_500_000, "cfg5_shard"
    dev = torch.device("cuda", 0)
    build, probe, _ = bench.make_workload(A, 0, 1, dev)
    idx = sn.CudaIndex.build_device(cuda_ctx, build["key"], build["start"], build["end"], torch.cuda.current_stream().cuda_stream)
    res = []
    for packed in (True, False):
        st, n, left, right = device_join(cuda_ctx, idx, probe, packed)
        dg = st.digest_device(left, right, n)
        assert bool((right[1:n] >= right[:n - 1]).all())
        c = counts_of(st, A.shard_rows)
        assert int(c.sum(dtype=np.uint64)) == n
        res.append((n, dg, c))
        del left, right
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1] and np.array_equal(res[0][2], res[1][2])
    # a slice of probe rows against the oracle, on the contigs those rows touch
    sl = slice(0, 20000)
    pk = probe["key"][sl].cpu().numpy().view(np.uint64)
    ps, pe = probe["start"][sl].cpu().numpy(), probe["end"][sl].cpu().numpy()
    small = torch.isin(build["contig"], torch.tensor([21, 23], dtype=torch.int32, device=dev))  # chr22, chrY
    bk = build["key"][small].cpu().numpy().view(np.uint64)
    bs, be = build["start"][small].cpu().numpy(), build["end"][small].cpu().numpy()
    mine = np.isin(pk, np.unique(bk))
    want = oracle.OracleIndex(bk, bs, be).counts(pk[mine], ps[mine], pe[mine])
    assert mine.sum() > 300 and np.array_equal(res[0][2][sl][mine], want)
