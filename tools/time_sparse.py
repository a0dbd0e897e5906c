"""Experiment helper: a LARGE sparse index (clusters of starts separated by gaps wider than 65536) — the case
the window-cut packed lines exist for: packed vs SoA kernels, digest equality and time."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import sequila_native_b200 as sn, bench
class A: build_rows = 100_000_000; shard_rows = 12_500_000; workload = "cfg5_shard"
dev = torch.device("cuda", 0); ctx = sn.CudaContext(0)
build, probe, _ = bench.make_workload(A, 0, 1, dev)
def spread(s):  # 2000-bp clusters every 140000 bp: same local density, gaps > one window
    st = s["start"].long(); w = (s["end"].long() - st)
    st2 = (st // 2000) * 140000 + st % 2000
    st2 = st2 % (2 ** 31 - 200000)
    return {"key": s["key"], "start": st2.int(), "end": (st2 + w).int()}
b, p = spread(build), spread(probe)
ts = torch.cuda.current_stream().cuda_stream
idx = sn.CudaIndex.build_device(ctx, b["key"], b["start"], b["end"], ts)
print("uses_packed", idx.uses_packed, "index MB", idx.bytes >> 20, "build_ms", round(idx.build_ms, 1))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
for packed in ("1", "0"):
    os.environ["SQ_PACKED"] = packed
    st = sn.CudaStream(ctx, cuda_stream=ts)
    n = st.probe_count_device(idx, p["key"], p["start"], p["end"])
    left = torch.empty(n, dtype=torch.int32, device=dev); right = torch.empty_like(left)
    fn = lambda: st.probe_join_device(idx, p["key"], p["start"], p["end"], left, right)
    for _ in range(3): fn()
    tt = []
    for _ in range(8):
        flush.fill_(1)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize(); tt.append(a.elapsed_time(e))
    res[packed] = st.digest_device(left, right, n)
    print("packed" if packed == "1" else "soa", "pairs", n, "join_ms", round(float(np.median(tt)), 3))
    del left, right
print("digests equal:", res["1"] == res["0"])
