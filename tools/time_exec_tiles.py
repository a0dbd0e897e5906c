"""exec-node line of bench.py for several tile sizes (sequila.cuda_coalesce_rows) and probe-side sizes: the tiles are
pipelined two deep, so the overlap needs a few tiles to show"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sequila_native_b200 as sn
import bench

out = {}
for n_probe in (2_000_000, 8_000_000):
    for rows in (131072, 262144, 524288, 1048576):
        r = bench.exec_node_line(sn, None, n_probe=n_probe, options={"cuda_coalesce_rows": rows})
        out[f"{n_probe}.{rows}"] = {"seconds": r["seconds"], "value": r["value"], "library_ms": r["library_ms"],
                                    "partitions_4_value": r["partitions_4"]["value"]}
        print(n_probe, rows, out[f"{n_probe}.{rows}"], file=sys.stderr)
print(json.dumps(out))
