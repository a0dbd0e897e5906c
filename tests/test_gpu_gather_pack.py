"""GPU: the several-columns-per-pass gathers (sq_index_pack_columns / sq_gather_pack_device /
sq_gather_probe_columns_device) against numpy `take` on the oracle-checked pairs — the materialise step of
process_probe_batch (interval_join.rs:1620-1632) for the six output columns of SURVEY §8(d)."""
import numpy as np
import pytest

import sequila_native_b200 as sn

pytestmark = pytest.mark.gpu


def device_side(side, dev):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v.view(np.int64) if v.dtype == np.uint64 else v)).to(dev) for k, v in side.items()}


@pytest.mark.parametrize("n_cols", [1, 2, 3, 4])
def test_packed_gathers_equal_take(cuda_ctx, oracle, n_cols):
    import torch
    dev = torch.device("cuda", 0)
    b, p = sn.synth.cfg2(scale=0.05)
    bd, pd = device_side(b, dev), device_side(p, dev)
    idx = sn.CudaIndex.build_device(cuda_ctx, bd["key"], bd["start"], bd["end"])
    st = sn.CudaStream(cuda_ctx)
    n = st.probe_count_device(idx, pd["key"], pd["start"], pd["end"])
    left = torch.empty(n, dtype=torch.int32, device=dev)
    right = torch.empty(n, dtype=torch.int32, device=dev)
    assert st.probe_join_device(idx, pd["key"], pd["start"], pd["end"], left, right) == n
    ol, orr, _ = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    l, r = left.cpu().numpy().view(np.uint32), right.cpu().numpy().view(np.uint32)
    assert np.array_equal(oracle.sorted_pairs(l, r), oracle.sorted_pairs(ol, orr))
    extra = np.arange(len(b["start"]), dtype=np.int32) * 7 - 3
    bcols = [b["contig"].astype(np.int32), b["start"], b["end"], extra][:n_cols]
    ids = [idx.add_column_device(torch.from_numpy(c).to(dev)) for c in bcols]
    pack = idx.pack_columns(ids)
    outs = [torch.full((n,), -1, dtype=torch.int32, device=dev) for _ in range(n_cols)]
    st.gather_pack_device(pack, outs)
    for c, o in zip(bcols, outs):
        assert np.array_equal(o.cpu().numpy(), c[l])
    pextra = (np.arange(len(p["start"]), dtype=np.int32) ^ 0x5A5A)
    pcols = [p["contig"].astype(np.int32), p["start"], p["end"], pextra][:n_cols]
    pouts = [torch.full((n,), -1, dtype=torch.int32, device=dev) for _ in range(n_cols)]
    st.gather_probe_columns_device([torch.from_numpy(c).to(dev) for c in pcols], pouts)
    for c, o in zip(pcols, pouts):
        assert np.array_equal(o.cpu().numpy(), c[r])
    # fewer outputs than the pack holds: the leading columns
    one = [torch.empty(n, dtype=torch.int32, device=dev)]
    st.gather_pack_device(pack, one)
    assert np.array_equal(one[0].cpu().numpy(), bcols[0][l])


def test_pack_argument_errors(cuda_ctx):
    import torch
    dev = torch.device("cuda", 0)
    b, p = sn.synth.cfg2(scale=0.001)
    bd, pd = device_side(b, dev), device_side(p, dev)
    idx = sn.CudaIndex.build_device(cuda_ctx, bd["key"], bd["start"], bd["end"])
    c4 = idx.add_column_device(bd["start"])
    c8 = idx.add_column_device(bd["key"])
    with pytest.raises(sn.SequilaCudaError):
        idx.pack_columns([c4, c8])  # 8-byte column
    with pytest.raises(sn.SequilaCudaError):
        idx.pack_columns([c4] * 5)
    with pytest.raises(sn.SequilaCudaError):
        idx.pack_columns([99])
    pack = idx.pack_columns([c4])
    st = sn.CudaStream(cuda_ctx)
    with pytest.raises(sn.SequilaCudaError):  # no emit yet
        st.gather_pack_device(pack, [torch.empty(1, dtype=torch.int32, device=dev)])
