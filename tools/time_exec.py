"""Experiment helper: throughput of the exec-node layer (Arrow RecordBatches in, joined RecordBatches out)."""
import os, sys, time
import numpy as np
import pyarrow as pa
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sequila_native_b200 as sn
from sequila_native_b200 import intervals as IV
from sequila_native_b200.interval_join import HashJoinDesc, optimize

name = os.environ.get("CFG", "cfg2")
scale = float(os.environ.get("SCALE", "1.0"))
batch_rows = int(os.environ.get("BATCH", "1000000"))
b, p = sn.synth.CONFIGS[name](scale=scale)
names = np.array(["chr%d" % (i + 1) for i in range(22)] + ["chrX", "chrY"])
COLS = ["contig", "pos_start", "pos_end"]
DICT = os.environ.get("DICT") is not None  # contig as DictionaryArray<int32, Utf8> (what a dictionary-encoding scan hands over)
def table(s):
    contig = (pa.DictionaryArray.from_arrays(pa.array(s["contig"].astype(np.int32)), pa.array(names)) if DICT
              else pa.array(names[s["contig"]]))
    return pa.record_batch([contig, pa.array(s["start"]), pa.array(s["end"])], names=COLS)
L, R = table(b), table(p)
cfg = sn.SequilaConfig(); sn.apply_set(cfg, "SET sequila.interval_join_algorithm TO cuda")
f = IV.parse_condition_sql("a.pos_start <= b.pos_end AND a.pos_end >= b.pos_start", "a", COLS, "b", COLS)
proj = None if os.environ.get("PROJ") is None else [int(x) for x in os.environ["PROJ"].split(",")]
plan = optimize(HashJoinDesc(L.schema, R.schema, [("contig", "contig")], f, projection=proj), cfg)
t0 = time.perf_counter(); plan.collect_build([L]); t1 = time.perf_counter()
print(f"{name}{' (dictionary-encoded contig)' if DICT else ''}: build {L.num_rows} rows: {(t1 - t0) * 1e3:.1f} ms")
batches = [R.slice(i, batch_rows) for i in range(0, R.num_rows, batch_rows)]
for rep in range(int(os.environ.get("REPS", "2"))):
    t0 = time.perf_counter(); rows = 0
    for rb in batches:
        out = plan.probe_batch(rb); rows += out.num_rows
    dt = time.perf_counter() - t0
    print(f"  probe {R.num_rows} rows in {len(batches)} batches -> {rows} rows: {dt * 1e3:.1f} ms  "
          f"({R.num_rows / dt / 1e6:.1f} M probe rows/s, {rows / dt / 1e6:.1f} M out rows/s)")
