"""Experiment helper: device time of the index build on the cfg5 build side (100M rows)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import sequila_native_b200 as sn, bench
class A: build_rows = int(os.environ.get("BUILD_ROWS", 100_000_000)); shard_rows = 1000; workload = "cfg5_shard"
dev = torch.device("cuda", 0); ctx = sn.CudaContext(0)
build, probe, _ = bench.make_workload(A, 0, 1, dev)
ts = torch.cuda.current_stream().cuda_stream
ms = []
for _ in range(5):
    idx = sn.CudaIndex.build_device(ctx, build["key"], build["start"], build["end"], ts); ms.append(idx.build_ms); del idx
print("build_ms", [round(x, 2) for x in ms])
