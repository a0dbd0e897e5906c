"""A/B: one vs two tiles per CTA of the emitting packed-line kernel (option cuda_probe_tiles) on the cfg5 shard, random
and position-sorted probe order, block 64 and 128; pair digest compared between the variants."""
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
ctx.set_option("cuda_staged_probe", "off")
build, probe, _, _ = bench.make_workload(A, 0, 1, dev)
ts = torch.cuda.current_stream().cuda_stream
idx = sn.CudaIndex.build_device(ctx, build["key"], build["start"], build["end"], ts)
st = sn.CudaStream(ctx, cuda_stream=ts)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
sp = bench.sort_by_position(probe)
n = st.probe_count_device(idx, probe["key"], probe["start"], probe["end"])
left = torch.empty(n, dtype=torch.int32, device=dev); right = torch.empty(n, dtype=torch.int32, device=dev)
ref = {}
for order, p in (("random", probe), ("sorted", sp)):
    for block in (128, 64):
        ctx.set_option("cuda_probe_block", block)
        for tiles in (1, 2):
            ctx.set_option("cuda_probe_tiles", tiles)
            for what, r in (("join", right), ("join_noright", None)):
                fn = lambda: st.probe_join_device(idx, p["key"], p["start"], p["end"], left, r)
                for _ in range(3):
                    assert fn() == n
                dig = (int(left.to(torch.int64).sum()), int((left.to(torch.int64) * (torch.arange(n, device=dev) % 1021)).sum()))
                if r is not None:
                    dig += (int((r.to(torch.int64) * (torch.arange(n, device=dev) % 1019)).sum()),)
                assert ref.setdefault((order, what), dig) == dig, (order, block, tiles, what)
                ms = bench.timed_steps(torch, fn, 10, flush)
                out[f"{order}.{what}.block{block}.tiles{tiles}"] = ms
                print(order, what, "block", block, "tiles", tiles, "%.4f ms" % ms, file=sys.stderr)
print(json.dumps(out))
