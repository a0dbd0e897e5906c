"""Host link ceiling of the box: pinned H2D, D2H and both at once (two CUDA streams), per copy size.
The end-to-end numbers of bench.py are bounded by these."""
import json
import sys
import time

import torch

dev = torch.device("cuda", 0)
out = {}
for mb in (1, 4, 16, 64, 256):
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    reps = max(4, 1024 // mb)

    def run(h2d, d2h):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    for _ in range(2):
        run(True, True)
    a, b, c = run(True, False), run(False, True), run(True, True)
    out[f"{mb}MiB"] = {"h2d_GBps": n / a / 1e9, "d2h_GBps": n / b / 1e9, "duplex_each_GBps": n / c / 1e9}
    print(mb, "MiB", out[f"{mb}MiB"], file=sys.stderr)
print(json.dumps(out))
