"""one launch of each probe kernel of the default bench workload (cfg5 shard), for ncu: packed join, packed count-only,
staged join on the position-sorted shard, rank join on cfg4"""
import os, sys
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
build, probe, _, _ = bench.make_workload(A, 0, 1, dev)
ts = torch.cuda.current_stream().cuda_stream
idx = sn.CudaIndex.build_device(ctx, build["key"], build["start"], build["end"], ts)
st = sn.CudaStream(ctx, cuda_stream=ts)
n = st.probe_count_device(idx, probe["key"], probe["start"], probe["end"])
left = torch.empty(n, dtype=torch.int32, device=dev); right = torch.empty(n, dtype=torch.int32, device=dev)
st.probe_join_device(idx, probe["key"], probe["start"], probe["end"], left, right)
sp = bench.sort_by_position(probe)
ctx.set_option("cuda_staged_probe", "on")
st.probe_join_device(idx, sp["key"], sp["start"], sp["end"], left, right)
st.probe_count_device(idx, sp["key"], sp["start"], sp["end"])
ctx.set_option("cuda_staged_probe", "auto")
del idx, build, probe, sp, left, right
b, p = sn.synth.cfg4()
bd, pd = bench.to_device(b, dev), bench.to_device(p, dev)
idx = sn.CudaIndex.build_device(ctx, bd["key"], bd["start"], bd["end"], ts)
n = st.probe_count_device(idx, pd["key"], pd["start"], pd["end"])
left = torch.empty(n, dtype=torch.int32, device=dev); right = torch.empty(n, dtype=torch.int32, device=dev)
st.probe_join_device(idx, pd["key"], pd["start"], pd["end"], left, right)
torch.cuda.synchronize()
