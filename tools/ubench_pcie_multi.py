"""Aggregate host-link ceiling of the box: k = 1, 2, 4, .. WORLD_SIZE ranks copy pinned buffers H2D and D2H at the same time
(one process per GPU, bound to the CPUs next to its GPU like bench.py), the others idle.  Run under torch.distributed.run.
The end-to-end numbers of bench.py at N GPUs are bounded by the k = N row."""
import json, os, sys, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
try:
    import pynvml
    pynvml.nvmlInit()
    cpus = os.sched_getaffinity(0)
    words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local), (max(cpus) // 64) + 1)
    near = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1} & cpus
    if near and world > 1:
        os.sched_setaffinity(0, near)
except Exception:
    pass
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")
n = 64 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device=dev); d_b = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def run(h2d, d2h, reps=24):
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
    return n * reps / (time.perf_counter() - t0) / 1e9

out = {}
k = 1
while k <= world:
    for mode, (a, b) in (("duplex", (True, True)), ("h2d", (True, False)), ("d2h", (False, True))):
        if world > 1:
            dist.barrier()
        g = run(a, b) if rank < k else 0.0
        if rank < k:
            run(a, b, 4)
        if world > 1:
            t = torch.tensor([g], dtype=torch.float64)
            dist.all_reduce(t)
            tot = float(t.item())
        else:
            tot = g
        out[f"{k}_gpus.{mode}"] = {"aggregate_GBps_per_direction": tot, "per_gpu_GBps": tot / k}
    k *= 2
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
