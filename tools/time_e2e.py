"""Sweep of the end-to-end host path (sq_driver_run over submit / collect): partitions x tiles x pipeline depth, cfg5 shard."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sequila_native_b200 as sn
from sequila_native_b200.cuda_join import CudaDriver
from sequila_native_b200 import _native as N
import bench

class A:
    build_rows = 100_000_000; shard_rows = int(os.environ.get("SHARD_ROWS", 12_500_000))
    workload, scaling, parallelism, probe_order, total_probe_rows = "cfg5_shard", "weak", "replicated", "random", 0
dev = torch.device("cuda", 0)
ctx = sn.CudaContext(0)
build, probe, _, _ = bench.make_workload(A, 0, 1, dev)
idx = sn.CudaIndex.build_device(ctx, build["key"], build["start"], build["end"], torch.cuda.current_stream().cuda_stream)
del build
hk = ctx.pinned_copy(probe["key"].cpu().numpy().view(np.uint64)); hs = ctx.pinned_copy(probe["start"].cpu().numpy()); he = ctx.pinned_copy(probe["end"].cpu().numpy())
n = len(hk)
out = []
for depth in (2, 3, 4):
    ctx.set_option("cuda_pipeline_depth", depth)
    for T in (1, 2, 4, 8):
        drv = CudaDriver(ctx, T)
        for tiles in (8, 16, 32, 64, 128):
            if tiles < T: continue
            for flags, name in ((0, "join"), (N.TILE_COUNT_ONLY | N.TILE_NO_COUNTS, "count")):
                if name == "count" and depth != 3: continue
                for _ in range(2): r = drv.run(idx, hk, hs, he, tiles, flags)
                t0 = time.perf_counter(); reps = 4
                for _ in range(reps): r = drv.run(idx, hk, hs, he, tiles, flags)
                ms = (time.perf_counter() - t0) * 1e3 / reps
                rec = {"depth": depth, "T": T, "tiles": tiles, "what": name, "ms": ms, "G_rows_s": n / ms / 1e6, "pairs": r["n_pairs"],
                       "d2h_GBps": r["d2h_bytes"] / ms / 1e6, "h2d_GBps": r["h2d_bytes"] / ms / 1e6, "regrown": r["regrown_tiles"]}
                out.append(rec)
                print(rec, file=sys.stderr)
        del drv
print(json.dumps(out))
