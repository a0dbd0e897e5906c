#!/bin/bash
# usage: tools_bench_all.sh [extra bench args]; prints one summary line per workload
for w in cfg5_shard cfg2 cfg3 cfg4; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 "$@" > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  tail -2 gpurun_out/bench_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json"))
    print("$w", "ms/step", round(d["ms_per_step"],4), "kern_ms", round(d["roofline"]["avg_launch_ms"],4), "roof", round(d["roofline"]["frac"],4), "Gprobes/s", round(d["value"]/1e9,3), "Gpairs/s", round(d["pairs_per_s"]/1e9,2), "e2e_ms", round(d["e2e"]["ms_per_step"],3), "build_ms", round(d["build"]["ms"],3))
except Exception as e:
    print("$w FAILED", e)
PY
done
