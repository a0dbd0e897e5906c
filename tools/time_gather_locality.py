"""What a position-local build-side gather would buy the materialise step: the same cfg5 shard joined against the build side
in its given (random) row order and against the SAME rows pre-sorted by (contig, start) — then left_idx is the sorted
position, the hits of a probe row are neighbours in the row-wise pack and one pair read touches a shared sector."""
import json, os, sys
import numpy as np, torch
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for order in ("given", "position_sorted"):
    b = build if order == "given" else bench.sort_by_position(build)
    idx = sn.CudaIndex.build_device(ctx, b["key"], b["start"], b["end"], ts)
    st = sn.CudaStream(ctx, cuda_stream=ts)
    n = st.probe_count_device(idx, probe["key"], probe["start"], probe["end"])
    left = torch.empty(n, dtype=torch.int32, device=dev); right = torch.empty(n, dtype=torch.int32, device=dev)
    assert st.probe_join_device(idx, probe["key"], probe["start"], probe["end"], left, right) == n
    cols = [idx.add_column_device(b[k]) for k in ("contig", "start", "end")]
    pack = idx.pack_columns(cols)
    outs = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(3)]
    fn = lambda: st.gather_pack_device(pack, outs)
    for _ in range(3):
        fn()
    ms = bench.timed_steps(torch, fn, 10, flush)
    li = left.long()[::997]
    assert torch.equal(outs[1][::997], b["start"][li])
    out[order] = {"pairs": n, "gather_pack_ms": ms}
    print(order, n, "%.4f ms" % ms, file=sys.stderr)
    del idx, st, left, right, outs
print(json.dumps(out))
