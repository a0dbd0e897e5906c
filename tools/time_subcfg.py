"""cfg2 / cfg3 / cfg4 on one GPU: rank-difference kernel vs the count / scan / write chain (join and count-only)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sequila_native_b200 as sn
import bench
dev = torch.device("cuda", 0)
ctx = sn.CudaContext(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
if os.environ.get("RPB"):
    ctx.set_option("cuda_rows_per_bin", os.environ["RPB"])
out = {}
for name in os.environ.get("CFGS", "cfg2 cfg3 cfg4").split():
    for rank in ("off", "force", "on"):
        ctx.set_option("cuda_rank_count", rank)
        r = bench.sub_config(sn, torch, ctx, name, dev, 6460.5, 10, flush)
        out[f"{name}.rank_{rank}"] = r
        print(name, "rank", rank, "join %.4f ms frac %.3f count %.4f ms build %.3f ms %s" % (r["ms_per_step"], r["roofline_frac"], r["count_only_ms"], r["build_ms"], r["kernels"]), file=sys.stderr)
ctx.set_option("cuda_rank_count", "on")
print(json.dumps(out))
