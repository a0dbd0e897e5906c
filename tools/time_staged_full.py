"""A/B on the full config: all 100M probe rows position-sorted (8 launches of 12.5M rows) against the 100M-row index,
k_probe_staged vs k_probe_packed.  MODE=profile runs ONE staged launch of the shard-sorted and one of the full-sorted case."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sequila_native_b200 as sn
import bench

class A:
    build_rows = 100_000_000; shard_rows = 12_500_000
    workload, scaling, parallelism, probe_order, total_probe_rows = "cfg5_shard", "weak", "replicated", "random", 0
dev = torch.device("cuda", 0)
ctx = sn.CudaContext(0)
profile = os.environ.get("MODE") == "profile"
build = sn.synth.counter_side_torch(A.build_rows, bench.BUILD_SEED, dev)
ts = torch.cuda.current_stream().cuda_stream
idx = sn.CudaIndex.build_device(ctx, build["key"], build["start"], build["end"], ts)
st = sn.CudaStream(ctx, cuda_stream=ts)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pf = bench.sort_by_position(sn.synth.counter_side_torch(100_000_000, bench.PROBE_SEED, dev))
tiles = bench.Tiles(pf, A.shard_rows)
shard = bench.sort_by_position(sn.synth.counter_side_torch(A.shard_rows, bench.PROBE_SEED, dev))
ctx.set_option("cuda_staged_probe", "off")
pairs = [st.probe_count_device(idx, *c) for c in tiles.cols]
cap = max(pairs)
left = torch.empty(cap, dtype=torch.int32, device=dev); right = torch.empty(cap, dtype=torch.int32, device=dev)
out = {}
if profile:
    ctx.set_option("cuda_staged_probe", "on")
    st.probe_join_device(idx, shard["key"], shard["start"], shard["end"], left, right)
    c = tiles.cols[3]
    st.probe_join_device(idx, c[0], c[1], c[2], left, right)
    ctx.set_option("cuda_staged_probe", "off")
    st.probe_join_device(idx, c[0], c[1], c[2], left, right)
    torch.cuda.synchronize()
    sys.exit(0)
for staged in ("off", "on"):
    ctx.set_option("cuda_staged_probe", staged)
    for what, right_ in (("join", right), ("join_noright", None)):
        def fn():
            return sum(st.probe_join_device(idx, c[0], c[1], c[2], left, right_) for c in tiles.cols)
        assert fn() == sum(pairs)
        ms = bench.timed_steps(torch, fn, 5, flush)
        out[f"full_sorted.{what}.staged_{staged}"] = ms
        print("full sorted", what, "staged", staged, "%.4f ms" % ms, "frac %.3f" % ((16e8 + 12.0 * sum(pairs)) / (ms * 1e-3) / 1e9 / 6460.5), file=sys.stderr)
    def fc():
        return sum(st.probe_count_device(idx, *c) for c in tiles.cols)
    assert fc() == sum(pairs)
    ms = bench.timed_steps(torch, fc, 5, flush)
    out[f"full_sorted.count.staged_{staged}"] = ms
    print("full sorted count staged", staged, "%.4f ms" % ms, file=sys.stderr)
print(json.dumps(out))
