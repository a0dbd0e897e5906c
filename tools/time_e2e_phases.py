"""Experiment helper: per-phase device times of the host entry point sq_probe_join on one sq_stream."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sequila_native_b200 as sn
import bench

class A: build_rows = 100_000_000; shard_rows = int(os.environ.get("ROWS", 390_625)); workload = "cfg5_shard"
dev = torch.device("cuda", 0)
ctx = sn.CudaContext(0)
build, probe, _ = bench.make_workload(A, 0, 1, dev)
idx = sn.CudaIndex.build_device(ctx, build["key"], build["start"], build["end"], torch.cuda.current_stream().cuda_stream)
hk = ctx.pinned_copy(probe["key"].cpu().numpy().view(np.uint64)); hs = ctx.pinned_copy(probe["start"].cpu().numpy()); he = ctx.pinned_copy(probe["end"].cpu().numpy())
st = sn.CudaStream(ctx)
n = st.probe_count(idx, hk, hs, he)
out = (ctx.pinned_empty(n, np.uint32), ctx.pinned_empty(n, np.uint32), None)
for _ in range(3): st.probe_join(idx, hk, hs, he, out)
st.set_profiling(True)
t0 = time.perf_counter()
for _ in range(20): st.probe_join(idx, hk, hs, he, out)
dt = (time.perf_counter() - t0) / 20 * 1e3
print("rows", A.shard_rows, "pairs", n, "wall_ms/tile", round(dt, 3), {k: round(v, 3) for k, v in st.phase_ms().items()})
print("bytes in", 16 * A.shard_rows / 1e6, "MB; out", (4 * n + 4 * A.shard_rows) / 1e6, "MB")
