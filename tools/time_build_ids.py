"""Where the build of a position-id index spends its extra time: sq_index_build and the add_* calls, one by one, in a fresh
context per round (what an exec node does), rows vs positions."""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sequila_native_b200 as sn
from sequila_native_b200 import synth, _native as N

n = int(os.environ.get("ROWS", 2_000_000))
L_len = np.maximum((synth.HG38 * (n / 100_000_000)).astype(np.int64), 1000)
b = synth.counter_side(n, 5001, lengths=L_len)
names = [s.encode() for s in synth.CONTIG_NAMES]
lens = np.array([len(s) for s in names], np.int64)[b["contig"]]
off = np.zeros(n + 1, np.int64); np.cumsum(lens, out=off[1:])
data = np.frombuffer(b"".join(names[c] for c in b["contig"][:0]), np.uint8)  # placeholder
table = np.frombuffer(b"".join(s.ljust(8, b"\0") for s in names), np.uint8).reshape(len(names), 8)
rows8 = table[b["contig"]]
mask = np.arange(8)[None, :] < lens[:, None]
data = np.ascontiguousarray(rows8[mask])
assert len(data) == off[-1]
lib = N.lib()
out = {}
for rnd in range(3):
    for mode in ("positions", "rows"):
        ctx = sn.CudaContext(0)
        ctx.set_option("cuda_build_ids", mode)
        t = [time.perf_counter()]
        idx = sn.CudaIndex.build(ctx, b["key"], b["start"], b["end"]); t.append(time.perf_counter())
        idx.add_column(b["start"]); t.append(time.perf_counter())
        idx.add_column(b["end"]); t.append(time.perf_counter())
        cid = C.c_int32(-1)
        rc = lib.sq_index_add_utf8_column(idx._h, off.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p), len(data), C.byref(cid))
        assert rc == 0; t.append(time.perf_counter())
        rec = dict(zip(("build", "add4_a", "add4_b", "add_utf8"), [round((y - x) * 1e3, 3) for x, y in zip(t, t[1:])]))
        out.setdefault(mode, []).append(rec)
        print(rnd, mode, rec, file=sys.stderr)
        del idx, ctx
print(json.dumps(out))
