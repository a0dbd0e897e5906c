"""One scan of a device-resident BED text (for `ncu --metrics gpu__time_duration.sum` launch lists)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import sequila_native_b200 as sn
from time_scan import bed_bytes
rows = int(os.environ.get("ROWS", 20_000_000))
b, _ = sn.synth.cfg5(nb=rows, np_=1000)
d_text = torch.frombuffer(bytearray(bed_bytes(b)), dtype=torch.uint8).cuda()
ctx = sn.CudaContext(0); st = sn.CudaStream(ctx)
for _ in range(2):
    sc = sn.CudaScan.from_text(st, d_text)
    print(sc.rows, sc.timing_ms)
