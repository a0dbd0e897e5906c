"""per-phase wall times of the exec node (option cuda_exec_trace) for a few coalesced tiles of 8192-row batches"""
import os, sys, time
import numpy as np, pyarrow as pa
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sequila_native_b200 as sn
from sequila_native_b200 import intervals as IV, synth
from sequila_native_b200.interval_join import HashJoinDesc, optimize
n_build = n_probe = 2_000_000
L_len = np.maximum((synth.HG38 * (n_build / 100_000_000)).astype(np.int64), 1000)
b = synth.counter_side(n_build, 5001, lengths=L_len); p = synth.counter_side(n_probe, 5002, lengths=L_len)
names = np.array(synth.CONTIG_NAMES); cols = ["contig", "pos_start", "pos_end"]
tab = lambda s: pa.record_batch([pa.array(names[s["contig"]]), pa.array(s["start"]), pa.array(s["end"])], names=cols)
L, R = tab(b), tab(p)
cfg = sn.SequilaConfig(); sn.apply_set(cfg, "SET sequila.interval_join_algorithm TO cuda")
f = IV.parse_condition_sql("a.pos_start <= b.pos_end AND a.pos_end >= b.pos_start", "a", cols, "b", cols)
plan = optimize(HashJoinDesc(L.schema, R.schema, [("contig", "contig")], f), cfg)
plan.collect_build([L])
batches = [R.slice(i, 8192) for i in range(0, R.num_rows, 8192)]
for rep in range(2):
    rows = sum(o.num_rows for o in plan.probe_batches(batches))
plan.set_option("cuda_exec_trace", 1)
t0 = time.perf_counter()
rows = sum(o.num_rows for o in plan.probe_batches(batches))
print("total %.2f ms for %d probe rows -> %d rows" % ((time.perf_counter() - t0) * 1e3, n_probe, rows), file=sys.stderr)
