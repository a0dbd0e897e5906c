"""for ncu: one packed build-column gather of the cfg5 shard's pairs with row ids (payload in build-row order) and one with
position ids (payload in the index's order)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sequila_native_b200 as sn
import bench

class A:
    build_rows = int(os.environ.get("BUILD_ROWS", 100_000_000)); shard_rows = int(os.environ.get("SHARD_ROWS", 12_500_000))
    workload, scaling, parallelism, probe_order, total_probe_rows = "cfg5_shard", "weak", "replicated", "random", 0
dev = torch.device("cuda", 0)
ctx = sn.CudaContext(0)
build, probe, _, _ = bench.make_workload(A, 0, 1, dev)
ts = torch.cuda.current_stream().cuda_stream
for ids in ("rows", "positions"):
    ctx.set_option("cuda_build_ids", ids)
    idx = sn.CudaIndex.build_device(ctx, build["key"], build["start"], build["end"], ts)
    st = sn.CudaStream(ctx, cuda_stream=ts)
    n = st.probe_count_device(idx, probe["key"], probe["start"], probe["end"])
    left = torch.empty(n, dtype=torch.int32, device=dev); right = torch.empty(n, dtype=torch.int32, device=dev)
    assert st.probe_join_device(idx, probe["key"], probe["start"], probe["end"], left, right) == n
    pack = idx.pack_columns([idx.add_column_device(build[k]) for k in ("contig", "start", "end")])
    outs = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(3)]
    st.gather_pack_device(pack, outs)
    torch.cuda.synchronize()
    print(ids, n, file=sys.stderr)
    del idx, st, left, right, outs
