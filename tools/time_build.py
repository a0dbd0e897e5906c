"""Build time of the cfg5 index (100M rows) with 32-bit and 64-bit sort keys, rows / position ids: sq_index_build_ms, best of 4."""
import json, os, sys
import numpy as np, torch
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
out = {}
for sort in ("auto", "wide"):
    for ids in ("rows", "positions"):
        ctx.set_option("cuda_build_sort", sort); ctx.set_option("cuda_build_ids", ids)
        ms = []
        for _ in range(4):
            idx = sn.CudaIndex.build_device(ctx, build["key"], build["start"], build["end"], ts)
            ms.append(idx.build_ms); bits = idx.sort_key_bits; by = idx.bytes
            del idx
        out[f"{sort}.{ids}"] = {"best_ms": min(ms), "all_ms": ms, "sort_key_bits": bits, "index_bytes": by}
        print(sort, ids, bits, ["%.3f" % m for m in ms], file=sys.stderr)
print(json.dumps(out))
