"""print the interesting numbers of a bench.py JSON line"""
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.3f G rows/s  ms/step %.4f  launches %s  clocks %s" % (d["value"] / 1e9, d["ms_per_step"], d["gpu_launches"], d.get("clocks")))
r = d["roofline"]; print("roofline frac %.4f  kernel %s  avg_launch_ms %.4f traffic %s (%s)" % (r["frac"], r["kernel"], r["avg_launch_ms"], r["traffic"], r.get("traffic_source")))
print("build", d["build"])
e = d["e2e"]; print("e2e %.3f G rows/s  ms %.3f  T=%s tiles=%s link %s phase %s" % (e["value"] / 1e9, e["ms_per_step"], e["partitions"], e["tiles"], e["link_GBps"], e["phase_ms"]))
if e.get("counts_u8"): print("e2e counts_u8 %.3f G rows/s ms %.3f" % (e["counts_u8"]["value"] / 1e9, e["counts_u8"]["ms_per_step"]))
c = e["count_only"]; print("e2e count-only %.3f G rows/s ms %.3f link %s" % (c["value"] / 1e9, c["ms_per_step"], c["link_GBps"]))
for k in ("locality", "count_only", "materialise", "full_config", "exec_node", "oracle_check", "balance"):
    if d.get(k): print(k, json.dumps(d[k])[:600])
for k, v in (d.get("sub_configs") or {}).items():
    print(k, "ms %.4f frac %.3f count_only_ms %.4f kernels %s pairs %d" % (v["ms_per_step"], v["roofline_frac"], v["count_only_ms"], v["kernels"], v["pairs"]))
if d.get("cpu_baseline"):
    c = d["cpu_baseline"]; print("cpu %.2f M rows/s on %d cores; si %s" % (c["value"] / 1e6, c["cores"], c["reference_superintervals"] and c["reference_superintervals"]["value"]))
