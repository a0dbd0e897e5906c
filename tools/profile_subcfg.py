"""one count-only and one join launch of a sub-config (for ncu)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sequila_native_b200 as sn
import bench
dev = torch.device("cuda", 0)
ctx = sn.CudaContext(0)
name = os.environ.get("CFG", "cfg3")
b, p = sn.synth.CONFIGS[name]()
bd, pd = bench.to_device(b, dev), bench.to_device(p, dev)
ts = torch.cuda.current_stream().cuda_stream
idx = sn.CudaIndex.build_device(ctx, bd["key"], bd["start"], bd["end"], ts)
st = sn.CudaStream(ctx, cuda_stream=ts)
n = st.probe_count_device(idx, pd["key"], pd["start"], pd["end"])
left = torch.empty(n, dtype=torch.int32, device=dev); right = torch.empty(n, dtype=torch.int32, device=dev)
st.probe_join_device(idx, pd["key"], pd["start"], pd["end"], left, right)
torch.cuda.synchronize()
