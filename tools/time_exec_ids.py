"""A/B on one box: the exec-node line of bench.py with the node's default (position ids: build payload in the index's
sorted order) and with `SET sequila.cuda_build_ids TO rows`, alternating, three rounds."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sequila_native_b200 as sn
import bench

out = {"positions": [], "rows": []}
for rnd in range(3):
    for mode in ("positions", "rows"):
        r = bench.exec_node_line(sn, None, build_ids=mode)
        out[mode].append({k: r[k] for k in ("seconds", "library_ms", "collect_build_seconds", "value")} | {"partitions_4_seconds": r["partitions_4"]["seconds"]})
        print(mode, out[mode][-1], file=sys.stderr)
print(json.dumps(out))
