"""Experiment helper: fused packed kernel vs SoA kernels on an L2-resident, high-fan-out narrow index."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sequila_native_b200 as sn

dev = torch.device("cuda", 0)
ctx = sn.CudaContext(0)
g = torch.Generator(device=dev); g.manual_seed(1)
nb, npq, L = 1_200_000, 10_000_000, int(os.environ.get("SPAN", 416_000))
keys = torch.tensor(sn.synth.key_hash(np.arange(24)).view(np.int64), device=dev)
def side(n):
    c = torch.randint(0, 24, (n,), generator=g, device=dev)
    w = torch.randint(50, 151, (n,), generator=g, device=dev)
    s = torch.randint(0, L, (n,), generator=g, device=dev)
    return {"key": keys[c].contiguous(), "start": s.int(), "end": (s + w - 1).int()}
b, p = side(nb), side(npq)
ts = torch.cuda.current_stream().cuda_stream
idx = sn.CudaIndex.build_device(ctx, b["key"], b["start"], b["end"], ts)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
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
    print("packed" if packed == "1" else "soa", "pairs", n, "hits/row", n / npq, "join_ms", float(np.median(tt)),
          "roof", (16 * npq + 12 * n) / (np.median(tt) * 1e-3) / 1e9 / 6460.5)
    del left, right
