#!/usr/bin/env python
"""bench.py — interval-overlap join probe throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One "step" = one pass of the probe hot path (count kernel -> emit kernel: interval_join.rs:1582-1618)
over this rank's probe batch against the resident build index.  Workload (default ``cfg5_shard``) =
BASELINE.json configs[4], the configuration the metric "at 1/2/4/8 B200" is quoted on: 100M build
intervals (hg38-weighted contigs, widths U{50..150}) replicated on every GPU, probes of the same
distribution sharded 12.5M per GPU (weak scaling; N=8 is exactly the 100M x 100M join).  Other
workloads (cfg2/cfg3/cfg4) are selectable with --workload.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHARD_ROWS = 12_500_000
BUILD_ROWS = 100_000_000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="cfg5_shard", choices=["cfg5_shard", "cfg2", "cfg3", "cfg4"])
    ap.add_argument("--build-rows", type=int, default=BUILD_ROWS)
    ap.add_argument("--shard-rows", type=int, default=SHARD_ROWS)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-partitions", type=int, default=8, help="host threads, one sq_stream each")
    ap.add_argument("--e2e-tiles", type=int, default=64, help="probe sub-tiles per step (all partitions)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-locality", action="store_true", help="skip the position-sorted side measurement")
    ap.add_argument("--no-materialise", action="store_true", help="skip the six-column gather measurement")
    ap.add_argument("--no-count-only", action="store_true", help="skip the count(1) measurement")
    ap.add_argument("--cpu-sample-contigs", type=int, default=8)
    ap.add_argument("--cpu-sample-probes", type=int, default=2_000_000)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# data
# ------------------------------------------------------------------------------------------------
def torch_uniform_side(n, seed, device, wlo=50, whi=150):
    """cfg5 distribution generated on the device (same parameters as synth._uniform_side)."""
    import torch
    from sequila_native_b200 import synth
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lengths = torch.tensor(synth.HG38, device=device)
    cum = torch.cumsum(lengths.double() / float(synth.HG38.sum()), 0)
    keys = torch.tensor(synth.key_hash(np.arange(24)).view(np.int64), device=device)
    contig = torch.searchsorted(cum, torch.rand(n, generator=g, device=device, dtype=torch.float64)).clamp_(max=23)
    L = lengths[contig]
    w = torch.randint(wlo, whi + 1, (n,), generator=g, device=device)
    start = (torch.rand(n, generator=g, device=device, dtype=torch.float64) * (L - w + 1).double()).long()
    end = start + w - 1
    return {"contig": contig.int(), "key": keys[contig].contiguous(), "start": start.int(), "end": end.int()}


def to_device(side, device):
    import torch
    return {"contig": torch.from_numpy(side["contig"]).to(device),
            "key": torch.from_numpy(side["key"].view(np.int64)).to(device),
            "start": torch.from_numpy(side["start"]).to(device), "end": torch.from_numpy(side["end"]).to(device)}


def to_host(side):
    return {"contig": side["contig"].cpu().numpy(), "key": side["key"].cpu().numpy().view(np.uint64),
            "start": side["start"].cpu().numpy(), "end": side["end"].cpu().numpy()}


def make_workload(args, rank, world, device):
    from sequila_native_b200 import synth
    if args.workload == "cfg5_shard":
        build = torch_uniform_side(args.build_rows, 5001, device)
        probe = torch_uniform_side(args.shard_rows, 5002 + 7919 * rank + 104729 * world, device)
        name = (f"cfg5 100Mx100M hg38-weighted U{{50..150}}: {args.build_rows} build rows replicated per GPU, "
                f"{args.shard_rows}-probe shard per GPU (N=8 == the full config)")
        return build, probe, name
    b, p = synth.CONFIGS[args.workload]()
    names = {"cfg2": "cfg2 1Mx1M 24 contigs x 10Mbp uniform U{50..150}",
             "cfg3": "cfg3 databio-shaped 1.2M build x 10M probe, skewed lengths",
             "cfg4": "cfg4 high fan-out 1Mx1M, build widths U{100k..500k}"}
    if world > 1:  # weak scaling: every rank probes its own re-seeded copy of the probe side
        rng = np.random.default_rng(rank + 1)
        perm = rng.permutation(len(p["key"]))
        p = {k: v[perm] for k, v in p.items()}
    return to_device(b, device), to_device(p, device), names[args.workload]


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every 5 ms from a thread (the
    timed region of the default run lasts tens of milliseconds, too short for `nvidia-smi -lms`)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, gpu_index):
        self.rows = []
        self.stop_flag = False
        self.h = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            idx = gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[gpu_index])
                except (ValueError, IndexError):
                    idx = gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.h = None

    def _pump(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    why = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    why = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.time(), sm, why))
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self, t0, t1):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.stop_flag = True
        self.t.join(timeout=1.0)
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows
        reasons = set()
        for _, _, why in rows:
            for bit, name in self.REASONS.items():
                if why & bit:
                    reasons.add(name)
        return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(reasons), "samples": len(rows)}


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference's coitrees path) on a bounded contig-subset sample
# ------------------------------------------------------------------------------------------------
def cpu_sample(args, build_h, probe_h):
    """Bounded sample of the workload for the CPU arm: every build AND probe row of the
    `cpu_sample_contigs` smallest contigs (same density / fan-out as the whole workload, since the join
    never crosses contigs), probe rows capped at --cpu-sample-probes."""
    from sequila_native_b200 import synth
    if args.workload == "cfg5_shard":
        sel = np.argsort(synth.HG38)[:args.cpu_sample_contigs]
        bm = np.isin(build_h["contig"], sel)
        pm = np.isin(probe_h["contig"], sel)
        what = f"contigs {sorted(int(c) for c in sel)} of the workload"
    else:
        bm = np.ones(len(build_h["key"]), bool)
        pm = np.ones(len(probe_h["key"]), bool)
        what = "whole build side"
    cap = args.cpu_sample_probes
    return ({k: build_h[k][bm] for k in ("key", "start", "end")},
            {k: probe_h[k][pm][:cap] for k in ("key", "start", "end")}, what)


def cpu_baseline(args, build_h, probe_h, threads, min_seconds=1.0, index=None):
    """Times the reference's CPU path (oracle/: C++ restatement of coitrees 0.4.0 AVX2 tree +
    interval_join.rs probe loop) on the box's host cores: 8192-row probe batches dealt round-robin to
    `threads` threads over one shared index (= DataFusion CollectLeft with target_partitions = threads).
    Passes over the sample are repeated until `min_seconds` of wall time (x threads = CPU work)."""
    from oracle import oracle as O
    b, p, what = cpu_sample(args, build_h, probe_h)
    idx = index or O.OracleIndex(b["key"], b["start"], b["end"], variant=8)
    secs, pairs = [], 0
    idx.time_probe(p["key"], p["start"], p["end"], threads=threads, batch_rows=8192)  # warm-up pass
    while sum(secs) < min_seconds and len(secs) < 64:
        sec, pairs, _ = idx.time_probe(p["key"], p["start"], p["end"], threads=threads, batch_rows=8192)
        secs.append(sec)
    sec = float(np.mean(secs))
    n = len(p["key"])
    # T = 1, the setting of the reference's own benchmarks (benches/databio_benchmark.rs:125,141, README.md:39-41:
    # target_partitions = 1), on a slice of the same sample so that it stays around a second
    one = None
    if threads > 1 and not index:
        m = max(min(n // max(threads // 2, 1), n), 1)
        s1, _, _ = idx.time_probe(p["key"][:m], p["start"][:m], p["end"][:m], threads=1, batch_rows=8192)
        one = {"value": m / s1, "cores": 1, "target_partitions": 1, "probe_rows": int(m), "seconds": s1}
    # the reference's OWN superintervals.hpp (its `SuperIntervals` arm, which its tests require to return the same rows
    # as `Coitrees`, IJ:1752-1758), compiled from /root/reference into oracle/_ref: a reference-authored anchor next to
    # the coitrees restatement, same sample, same threads, same batch dealing
    ref_si = None
    if not index and O.ref_timing_available():
        try:
            si = O.RefSuperIntervalsIndex(b["key"], b["start"], b["end"])
            si.time_probe(p["key"], p["start"], p["end"], threads=threads, batch_rows=8192)  # warm-up
            s_sec, s_pairs = si.time_probe(p["key"], p["start"], p["end"], threads=threads, batch_rows=8192)
            assert s_pairs == pairs, (s_pairs, pairs)  # the two reference algorithms agree on the sample
            ref_si = {"value": n / s_sec, "unit": "probe intervals/s", "kind": "reference", "cores": threads,
                      "what": "superintervals.hpp from the reference tree (Algorithm::SuperIntervals arm), oracle/_ref/libsi_ref.so",
                      "seconds": s_sec, "index_build_seconds": si.build_seconds}
            del si
        except Exception as ex:  # the anchor is optional; the coitrees restatement above is the baseline
            ref_si = {"unavailable": repr(ex)}
    return {"reference_superintervals": ref_si, "one_thread": one, "value": n / sec, "unit": "probe intervals/s", "cores": threads, "target_partitions": threads, "kind": "port",
            "sample": f"{what}: {len(b['key'])} build rows, {n} probe rows, {pairs} pairs per pass, {len(secs)} passes of "
                      f"{sec:.3f}s ({sum(secs) * threads:.0f} core-seconds), index build {idx.build_seconds:.2f}s (1 thread); "
                      f"coitrees 0.4.0 AVX2-layout restatement, 8192-row batches dealt to {threads} thread(s)",
            "pairs_per_s": pairs / sec, "seconds": sec, "_index": idx}


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Rust crate
    cannot be built here: no cargo/rustc in the image) with all host threads, bounded sample."""
    if rank != 0:
        return
    from sequila_native_b200 import synth
    threads = os.cpu_count() or 1
    if args.workload == "cfg5_shard":
        # generate only the sample's contigs on the host (the GPU is not needed for this arm)
        sel = np.argsort(synth.HG38)[:args.cpu_sample_contigs]
        frac = synth.HG38[sel].sum() / synth.HG38.sum()
        w = synth.HG38[sel] / synth.HG38[sel].sum()

        def side(n, seed):
            rng = np.random.default_rng(seed)
            c = sel[rng.choice(len(sel), size=n, p=w)]
            L = synth.HG38[c]
            wd = rng.integers(50, 151, n)
            st = (rng.random(n) * (L - wd + 1)).astype(np.int64)
            return synth._table(c, st, st + wd - 1)
        build_h = side(int(args.build_rows * frac), 5001)
        probe_h = side(min(int(args.shard_rows * frac), args.cpu_sample_probes), 5002)
        name = (f"cfg5 100Mx100M hg38-weighted U{{50..150}}: {args.build_rows} build rows replicated per GPU, "
                f"{args.shard_rows}-probe shard per GPU (N=8 == the full config)")
    else:
        build_h, probe_h = synth.CONFIGS[args.workload]()
        name = args.workload
    vals = []
    base = None
    index = None
    one_thread = None
    ref_si = None
    t_all = time.time()
    for it in range(args.warmup + args.steps):  # one step = one timed pass set over the bounded sample
        base = cpu_baseline(args, build_h, probe_h, threads, min_seconds=0.5, index=index)
        index = base.pop("_index")
        one_thread = base.get("one_thread") or (one_thread if it else None)  # measured with the first pass set
        ref_si = base.get("reference_superintervals") or (ref_si if it else None)
        if it >= args.warmup:
            vals.append(base)
        if len(vals) >= 3 and time.time() - t_all > 150:
            break
    v = float(np.mean([b["value"] for b in vals]))
    ms = float(np.mean([b["seconds"] for b in vals])) * 1e3
    base["value"] = v
    base["one_thread"] = one_thread
    base["reference_superintervals"] = ref_si
    print(json.dumps({
        "impl": "reference", "metric": "probe_intervals_per_s", "value": v, "unit": "probe intervals/s",
        "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": name, "note": "CPU reference arm: oracle port of coitrees path, bounded sample"},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "probe intervals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pairs_per_s": base["pairs_per_s"], "gpu_launches": 0,
    }))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import sequila_native_b200 as sn

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the cuda interval join has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    # NUMA locality of the pinned staging buffers: run this rank on the CPUs next to its GPU (what a
    # launcher with --bind-to would do); the CPU baseline below widens the mask again
    all_cpus = os.sched_getaffinity(0)
    if world > 1 and not os.environ.get("SQ_NO_AFFINITY"):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(all_cpus) // 64) + 1)
            near = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1} & all_cpus
            if near:
                os.sched_setaffinity(0, near)
        except Exception:
            pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when NCCL_DEBUG is set on the box; stdout carries the
        # ONE JSON line of the contract, so the banner is routed to stderr while the communicator forms
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"

    ctx = sn.CudaContext(local_rank)
    build, probe, wname = make_workload(args, rank, world, device)
    torch.cuda.synchronize()
    n_build, n_probe = build["key"].numel(), probe["key"].numel()

    # ---- build (once per query: collect_left_input, interval_join.rs:597-689); timed on its own
    tstream = torch.cuda.current_stream().cuda_stream
    idx = sn.CudaIndex.build_device(ctx, build["key"], build["start"], build["end"], tstream)
    build_ms = [idx.build_ms]
    for _ in range(2):
        idx = None
        idx = sn.CudaIndex.build_device(ctx, build["key"], build["start"], build["end"], tstream)
        build_ms.append(idx.build_ms)
    build_best = min(build_ms)
    index_bytes = idx.bytes  # before the materialise step below registers payload columns and their row-wise pack

    st = sn.CudaStream(ctx, cuda_stream=tstream)
    n_pairs = st.probe_count_device(idx, probe["key"], probe["start"], probe["end"])
    left = torch.empty(max(n_pairs, 1), dtype=torch.int32, device=device)
    right = torch.empty(max(n_pairs, 1), dtype=torch.int32, device=device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2

    def step():
        # count -> look-back scan -> write as ONE fused kernel pass (sq_probe_join_device)
        return st.probe_join_device(idx, probe["key"], probe["start"], probe["end"], left, right)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # parity guard inside the bench: digest of the emitted pairs must be reproducible and the pair count
    # must equal the sum of per-row counts (the oracle comparison itself lives in tests/ and smoke())
    dg = st.digest_device(left, right, n_pairs)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    st.set_profiling(True)
    launches0 = st.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0 = time.time()
    for a, b in evs:
        flush.fill_(1)  # L2 flush between timed iterations (outside the event pair)
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.time()
    phases = st.phase_ms()
    st.set_profiling(False)
    launches = st.launches - launches0
    step_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    clocks = sampler.stop(t0, t1) if sampler else None

    from sequila_native_b200.sharding import reduce_step
    step_ms_max, probes_total, pairs_total = reduce_step(step_ms, n_probe, n_pairs, device)
    value = probes_total / (step_ms_max * 1e-3)

    # ---- end to end through the host C ABI: pinned host inputs -> H2D -> kernels -> D2H pairs.
    # The step's probe batch is cut into sub-tiles dealt to `--e2e-partitions` host threads, each
    # with its own sq_stream (= one IntervalJoinStream per DataFusion partition over one shared
    # index, PartitionMode::CollectLeft): H2D of one partition overlaps kernels and D2H of others.
    import concurrent.futures as cf
    T = max(1, args.e2e_partitions)
    tiles_per = max(1, args.e2e_tiles // T)
    n_tiles = T * tiles_per
    bounds = np.linspace(0, n_probe, n_tiles + 1).astype(np.int64)
    hk = ctx.pinned_copy(probe["key"].cpu().numpy().view(np.uint64))
    hs = ctx.pinned_copy(probe["start"].cpu().numpy())
    he = ctx.pinned_copy(probe["end"].cpu().numpy())
    # per-tile pair counts size each partition's pinned output buffers (DataFusion would size them
    # from the previous batch and retry on SQ_ECAPACITY)
    tile_pairs = []
    tmp_st = sn.CudaStream(ctx)
    for t in range(n_tiles):
        lo, hi = int(bounds[t]), int(bounds[t + 1])
        tile_pairs.append(tmp_st.probe_count(idx, hk[lo:hi], hs[lo:hi], he[lo:hi]))
    del tmp_st
    assert sum(tile_pairs) == n_pairs
    workers = []
    for w in range(T):
        mine = list(range(w, n_tiles, T))
        cap = max(max(tile_pairs[t] for t in mine), 1)
        workers.append({"st": sn.CudaStream(ctx), "tiles": mine,
                        "out": (ctx.pinned_empty(cap, np.uint32), ctx.pinned_empty(cap, np.uint32))})

    def run_partition(wk):
        got = 0
        for t in wk["tiles"]:
            lo, hi = int(bounds[t]), int(bounds[t + 1])
            out = (wk["out"][0], wk["out"][1], None)  # index pairs; the optional per-row counts stay on the device
            got += wk["st"].probe_join(idx, hk[lo:hi], hs[lo:hi], he[lo:hi], out)
        return got

    pool = cf.ThreadPoolExecutor(T)

    def e2e_step():
        return sum(pool.map(run_partition, workers))

    for _ in range(2):
        assert e2e_step() == n_pairs
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e_ms = (time.perf_counter() - e0) * 1e3 / args.e2e_steps
    e_ms_max, _, _ = reduce_step(e_ms, 0, 0, device)
    e2e_value = probes_total / (e_ms_max * 1e-3)

    # the same through sq_probe_count: `select count(1)` end to end (host columns in, one number out) — the query
    # shape of the reference's own benchmarks (queries/q1-coitrees.sql:16-19); only the H2D copy is left on the wire
    def count_partition(wk):
        got = 0
        for t in wk["tiles"]:
            lo, hi = int(bounds[t]), int(bounds[t + 1])
            got += wk["st"].probe_count(idx, hk[lo:hi], hs[lo:hi], he[lo:hi])
        return got

    assert sum(pool.map(count_partition, workers)) == n_pairs
    torch.cuda.synchronize()
    c0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        sum(pool.map(count_partition, workers))
    torch.cuda.synchronize()
    ec_ms = (time.perf_counter() - c0) * 1e3 / args.e2e_steps
    ec_ms_max, _, _ = reduce_step(ec_ms, 0, 0, device)
    pool.shutdown()

    # ---- count only: `select count(1) from a join b on ...` is what the reference's own benchmark queries run
    # (queries/q1-coitrees.sql:16-19, benches/databio_benchmark.rs); no pair is written
    count_only = None
    if not args.no_count_only:
        for _ in range(3):
            assert st.probe_count_device(idx, probe["key"], probe["start"], probe["end"]) == n_pairs
        torch.cuda.synchronize()
        st.set_profiling(True)
        for _ in range(min(args.steps, 10)):
            flush.fill_(1)
            st.probe_count_device(idx, probe["key"], probe["start"], probe["end"])
        torch.cuda.synchronize()
        c_ms = st.phase_ms()["join"]
        st.set_profiling(False)
        count_only = {"query": "count(1) of the join", "avg_launch_ms": c_ms, "value": n_probe / (c_ms * 1e-3),
                      "unit": "probe intervals/s", "algorithmic_bytes": 16.0 * n_probe + 4.0 * n_pairs,
                      "roofline_frac": (16.0 * n_probe + 4.0 * n_pairs) / (c_ms * 1e-3) / 1e9 / hbm_peak}

    # ---- materialise (process_probe_batch's `take` per output column, interval_join.rs:1620-1632): the six output
    # columns of SURVEY §8(d) — contig (dictionary id, int32), pos_start, pos_end of both sides — gathered on the
    # device from the pairs of the last step; B_gather = 8 B pair read + 2 x 4 B per column = 56 B per pair.
    materialise = None
    if not args.no_materialise:
        assert step() == n_pairs
        cols = [idx.add_column_device(build[k]) for k in ("contig", "start", "end")]
        outs = [torch.empty(max(n_pairs, 1), dtype=torch.int32, device=device) for _ in range(6)]

        pack = idx.pack_columns(cols)  # the three build columns row-wise: one random read per pair serves all
        pvals = [probe[k] for k in ("contig", "start", "end")]

        def gather_step():  # two launches: build pack, probe columns
            st.gather_pack_device(pack, outs[:3])
            st.gather_probe_columns_device(pvals, outs[3:])

        def take_step():  # `take` column by column, what the reference's loop does (six launches)
            for c, o in zip(cols, outs[:3]):
                st.gather_build_device(c, o)
            for v, o in zip(pvals, outs[3:]):
                st.gather_probe_device(v, o)

        def timed(fn):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            evs_ = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 10))]
            for a, b in evs_:
                flush.fill_(1)
                a.record()
                fn()
                b.record()
            torch.cuda.synchronize()
            return float(np.mean([a.elapsed_time(b) for a, b in evs_]))

        take_ms = timed(take_step)
        g_ms = timed(gather_step)
        # spot check against torch indexing (same device data)
        li = left[:n_pairs].long()[:: max(n_pairs // 100000, 1)]
        assert torch.equal(outs[1][:n_pairs][:: max(n_pairs // 100000, 1)], build["start"][li])
        ri = right[:n_pairs].long()[:: max(n_pairs // 100000, 1)]
        assert torch.equal(outs[5][:n_pairs][:: max(n_pairs // 100000, 1)], probe["end"][ri])
        g_bytes = 56.0 * n_pairs
        materialise = {"columns": 6, "ms": g_ms, "pairs_per_s": n_pairs / (g_ms * 1e-3) if g_ms else None,
                       "algorithmic_bytes": g_bytes, "roofline_frac": g_bytes / (g_ms * 1e-3) / 1e9 / hbm_peak if g_ms else None,
                       "launches": 2, "how": "build columns packed row-wise (16 B per build row): one random read per pair; "
                       "probe columns in one pass over right_idx", "take_per_column_ms": take_ms}
        del outs

    # ---- same rows in position order (what BAM / BED inputs look like): an explanatory side line, N=1 only.
    # The headline workload probes in random order on purpose; with locality the walk's line reads hit L2
    # instead of costing one DRAM request each, and this shows how much of the gap to the
    # streaming roofline is the access pattern rather than the kernel.
    locality = None
    if world == 1 and not args.no_locality:
        order = torch.argsort(probe["contig"].to(torch.int64) * (1 << 32) + probe["start"].to(torch.int64))
        sp = {k: v[order].contiguous() for k, v in probe.items()}
        del order

        def sorted_step():
            return st.probe_join_device(idx, sp["key"], sp["start"], sp["end"], left, right)

        for _ in range(3):
            assert sorted_step() == n_pairs
        torch.cuda.synchronize()
        st.set_profiling(True)
        for _ in range(min(args.steps, 10)):
            flush.fill_(1)
            sorted_step()
        torch.cuda.synchronize()
        s_ms = st.phase_ms()["join"]
        st.set_profiling(False)
        locality = {"probe_order": "position-sorted (contig, start)", "avg_launch_ms": s_ms,
                    "value": n_probe / (s_ms * 1e-3), "unit": "probe intervals/s",
                    "roofline_frac": (16.0 * n_probe + 12.0 * n_pairs) / (s_ms * 1e-3) / 1e9 / hbm_peak}
        del sp

    if rank == 0:
        # ---- roofline of the dominant kernel (algorithmic bytes) -----------------
        # B_probe of SURVEY.md §8(d) with u64 key hashes consumed on the device:
        # 16 B per probe row (key hash 8 + start 4 + end 4) + 12 B per emitted pair
        # (read the hit's build row id 4, write (left,right) 8).  One launch = the whole tile.
        dom = "k_probe_packed" if idx.uses_packed else "k_probe_count+k_tile_scan+k_probe_write"
        b_dom = 16.0 * n_probe + 12.0 * n_pairs
        t_dom = phases["join"]
        achieved = b_dom / (t_dom * 1e-3) / 1e9 if t_dom > 0 else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        except Exception:
            pass
        result = {
            "metric": "probe_intervals_per_s", "value": value, "unit": "probe intervals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms_max,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": wname, "build_rows": n_build, "probe_rows_per_gpu": n_probe,
                       "pairs_per_gpu": n_pairs, "l2": "256 MiB flush write between timed steps; inputs also > L2",
                       "parallelism": f"probe shards x{world}, build index replicated, no collective on the data path"},
            "pairs_per_s": pairs_total / (step_ms_max * 1e-3),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": b_dom, "avg_launch_ms": t_dom,
                         # DRAM bytes the launch really moves (ncu) over its duration: how busy HBM is
                         "traffic_rate": (traffic / (t_dom * 1e-3) / 1e9) if traffic and t_dom > 0 else None,
                         "traffic_frac": (traffic / (t_dom * 1e-3) / 1e9 / hbm_peak) if traffic and t_dom > 0 else None},
            "build": {"ms": build_best, "rows_per_s": n_build / (build_best * 1e-3) if build_best else None,
                      "roofline_frac": (24.0 * n_build / (build_best * 1e-3) / 1e9 / hbm_peak) if build_best else None,
                      "index_bytes": index_bytes, "keys": idx.keys},
            "e2e": {"value": e2e_value, "unit": "probe intervals/s", "h2d_bytes_per_step": 16 * n_probe,
                    "d2h_bytes_per_step": (4 * n_pairs + 4 * n_probe + 16 * n_tiles) if os.environ.get("SQ_RLE_WIRE", "1") != "0"
                    else 8 * n_pairs + 16 * n_tiles,
                    "wire": "left_idx u32 per pair + per-row counts u32; right_idx is expanded from the counts into the caller's "
                            "host buffer by the calling thread (interval_join.rs:1611-1618)", "ms_per_step": e_ms_max,
                    "steps": args.e2e_steps, "api": "sq_probe_join (host C ABI), pinned host buffers", "partitions": T, "tiles": n_tiles,
                    "count_only": {"value": probes_total / (ec_ms_max * 1e-3), "unit": "probe intervals/s", "ms_per_step": ec_ms_max,
                                   "api": "sq_probe_count (host C ABI): count(1) of the join, host columns in, one number out",
                                   "h2d_bytes_per_step": 16 * n_probe, "d2h_bytes_per_step": 8 * n_tiles}},
            "gpu_launches": int(launches), "clocks": clocks,
            "digest": {"pairs": dg[0], "sum": dg[1], "xor": dg[2]},
        }
        if locality:
            result["locality"] = locality
        if materialise:
            result["materialise"] = materialise
        if count_only:
            result["count_only"] = count_only
        if not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)
            if args.workload == "cfg5_shard":
                bh, ph = _sample_to_host(build, args), _sample_to_host(probe, args)
            else:
                bh, ph = to_host(build), to_host(probe)
            result["cpu_baseline"] = cpu_baseline(args, bh, ph, threads=os.cpu_count() or 1, min_seconds=1.5)
            result["cpu_baseline"].pop("_index", None)
        print(json.dumps(result))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _sample_to_host(side, args):
    """copy only the CPU-baseline sample's contigs back to the host"""
    import torch
    from sequila_native_b200 import synth
    sel = torch.tensor(np.argsort(synth.HG38)[:args.cpu_sample_contigs].astype(np.int32), device=side["contig"].device)
    m = torch.isin(side["contig"], sel)
    return to_host({k: v[m] for k, v in side.items()})


if __name__ == "__main__":
    main()
