"""Experiment helper (not part of the product): time count-only vs fused join on the cfg5 shard."""
import os, sys, json
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sequila_native_b200 as sn
import bench

class A: build_rows = int(os.environ.get("BUILD_ROWS", 100_000_000)); shard_rows = int(os.environ.get("SHARD_ROWS", 12_500_000)); workload = "cfg5_shard"
dev = torch.device("cuda", 0)
ctx = sn.CudaContext(0)
build, probe, _ = bench.make_workload(A, 0, 1, dev)
ts = torch.cuda.current_stream().cuda_stream
idx = sn.CudaIndex.build_device(ctx, build["key"], build["start"], build["end"], ts)
st = sn.CudaStream(ctx, cuda_stream=ts)
n_pairs = st.probe_count_device(idx, probe["key"], probe["start"], probe["end"])
left = torch.empty(n_pairs, dtype=torch.int32, device=dev); right = torch.empty_like(left)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    ts_ = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts_.append(a.elapsed_time(b))
    return float(np.median(ts_))
print("pairs", n_pairs, "build_ms", idx.build_ms, "index MB", idx.bytes >> 20)
print("count_only_ms", timeit(lambda: st.probe_count_device(idx, probe["key"], probe["start"], probe["end"])))
print("join_ms", timeit(lambda: st.probe_join_device(idx, probe["key"], probe["start"], probe["end"], left, right)))
print("join_noright_ms", timeit(lambda: st.probe_join_device(idx, probe["key"], probe["start"], probe["end"], left, None)))
