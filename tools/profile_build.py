"""for ncu: one warm build, then one build with 32-bit and one with 64-bit sort keys of the cfg5 index (100M rows)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sequila_native_b200 as sn
import bench

class A:
    build_rows = int(os.environ.get("BUILD_ROWS", 100_000_000)); shard_rows = 1000
    workload, scaling, parallelism, probe_order, total_probe_rows = "cfg5_shard", "weak", "replicated", "random", 0
dev = torch.device("cuda", 0)
ctx = sn.CudaContext(0)
build, probe, _, _ = bench.make_workload(A, 0, 1, dev)
ts = torch.cuda.current_stream().cuda_stream
for sort in ("auto", "auto", "wide"):
    ctx.set_option("cuda_build_sort", sort)
    idx = sn.CudaIndex.build_device(ctx, build["key"], build["start"], build["end"], ts)
    print(sort, idx.sort_key_bits, idx.build_ms, file=sys.stderr)
    del idx
