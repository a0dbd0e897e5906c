"""A/B: k_probe_staged vs k_probe_packed on the cfg5 shard (12.5M probe rows, 100M-row index) — position-sorted and
random probe order, join and count-only — and on the full 100M-row sorted probe side."""
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
idx = sn.CudaIndex.build_device(ctx, build["key"], build["start"], build["end"], ts)
st = sn.CudaStream(ctx, cuda_stream=ts)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
sp = bench.sort_by_position(probe)
n = st.probe_count_device(idx, probe["key"], probe["start"], probe["end"])
left = torch.empty(n, dtype=torch.int32, device=dev); right = torch.empty(n, dtype=torch.int32, device=dev)
for order, p in (("sorted", sp), ("random", probe)):
    for staged in ("off", "on"):
        ctx.set_option("cuda_staged_probe", staged)
        for what, fn in (("join", lambda: st.probe_join_device(idx, p["key"], p["start"], p["end"], left, right)),
                         ("join_noright", lambda: st.probe_join_device(idx, p["key"], p["start"], p["end"], left, None)),
                         ("count", lambda: st.probe_count_device(idx, p["key"], p["start"], p["end"]))):
            for _ in range(3):
                assert fn() == n
            ms = bench.timed_steps(torch, fn, 10, flush)
            out[f"{order}.{what}.staged_{staged}"] = ms
            print(order, what, "staged", staged, "%.4f ms" % ms, file=sys.stderr)
print(json.dumps(out))
