#!/usr/bin/env python
"""bench.py — interval-overlap join probe throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One "step" = one pass of the probe hot path (search -> count -> scan -> write: interval_join.rs:1582-1618) over
this rank's probe rows against the resident build index.  Workload (default) = BASELINE.json configs[4], the
configuration the metric "at 1/2/4/8 B200" is quoted on: 100M build intervals (hg38-weighted contigs, widths
U{50..150}) replicated on every GPU, the 100M probe rows of the same distribution sharded 12.5M per GPU (weak
scaling: rank r probes rows [r, r+1) x 12.5M of the probe side, so N=8 is exactly the 100M x 100M join).
`--scaling strong` fixes the total at 100M probe rows (N=1 runs all of them, in 12.5M-row launches);
`--parallelism contig` shards contigs instead of probe rows (PartitionMode::Partitioned, interval_join.rs:488-503).
Rows come from a counter-based generator (synth.counter_side*): both arms and every rank see the same rows.
At N=1 the line also carries the full 100M x 100M config, the other synthetic configs (cfg2/cfg3/cfg4), the
position-sorted variant, count-only, materialise, the exec-node layer and the CPU baseline.  ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import hashlib
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
BUILD_SEED, PROBE_SEED = 5001, 5002
CHECK_CONTIGS = (20, 21, 23)  # chr21, chr22, chrY: the rows every rank compares with the oracle (5 % of the workload)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="cfg5_shard", choices=["cfg5_shard", "cfg2", "cfg3", "cfg4"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--parallelism", default="replicated", choices=["replicated", "contig"])
    ap.add_argument("--probe-order", default="random", choices=["random", "sorted"])
    ap.add_argument("--build-rows", type=int, default=BUILD_ROWS)
    ap.add_argument("--shard-rows", type=int, default=SHARD_ROWS)
    ap.add_argument("--total-probe-rows", type=int, default=0, help="strong scaling / contig sharding: rows of the whole probe side (default 8 x shard rows)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-partitions", type=int, default=2, help="host threads, one sq_stream each")
    ap.add_argument("--e2e-tiles", type=int, default=8, help="probe tiles per step (all partitions): 1.56M rows each for the 12.5M-row shard, "
                    "what a host that coalesces 8192-row batches would submit (profiles/r02_e2e_sweep.json)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-locality", action="store_true", help="skip the position-sorted side measurement")
    ap.add_argument("--no-materialise", action="store_true", help="skip the six-column gather measurement")
    ap.add_argument("--no-count-only", action="store_true", help="skip the count(1) measurement")
    ap.add_argument("--no-sub-configs", action="store_true", help="skip the cfg2/cfg3/cfg4 sub-lines")
    ap.add_argument("--no-full-config", action="store_true", help="skip the 100M x 100M single-GPU sub-line")
    ap.add_argument("--no-exec", action="store_true", help="skip the exec-node (Arrow batches) sub-line")
    ap.add_argument("--no-oracle-check", action="store_true", help="skip the per-rank comparison with the oracle")
    ap.add_argument("--cpu-sample-contigs", type=int, default=8)
    ap.add_argument("--cpu-sample-probes", type=int, default=2_000_000)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# data
# ------------------------------------------------------------------------------------------------
def to_device(side, device):
    import torch
    return {"contig": torch.from_numpy(side["contig"]).to(device),
            "key": torch.from_numpy(side["key"].view(np.int64)).to(device),
            "start": torch.from_numpy(side["start"]).to(device), "end": torch.from_numpy(side["end"]).to(device)}


def to_host(side):
    return {"contig": side["contig"].cpu().numpy(), "key": side["key"].cpu().numpy().view(np.uint64),
            "start": side["start"].cpu().numpy(), "end": side["end"].cpu().numpy()}


def sort_by_position(side):
    import torch
    order = torch.argsort(side["contig"].to(torch.int64) * (1 << 32) + side["start"].to(torch.int64))
    return {k: v[order].contiguous() for k, v in side.items()}


def total_probe_rows(args):
    return args.total_probe_rows or 8 * args.shard_rows


def probe_range(args, rank, world):
    """rows [first, first + n) of the probe side this rank owns"""
    if args.scaling == "weak":
        return rank * args.shard_rows, args.shard_rows
    tot = total_probe_rows(args)
    lo, hi = tot * rank // world, tot * (rank + 1) // world
    return lo, hi - lo


def make_workload(args, rank, world, device):
    import torch
    from sequila_native_b200 import synth
    from sequila_native_b200.sharding import assign_keys_lpt
    if args.workload == "cfg5_shard":
        build = synth.counter_side_torch(args.build_rows, BUILD_SEED, device)
        info = {}
        if getattr(args, "parallelism", "replicated") == "contig":
            # PartitionMode::Partitioned: contigs are LPT-assigned to ranks by their build-row weight; a rank builds
            # only its contigs and receives only the probe rows of those contigs (DataFusion's RepartitionExec)
            mine = torch.tensor(sorted(assign_keys_lpt(synth.HG38, world)[rank]), dtype=torch.int32, device=device)
            probe = synth.counter_side_torch(total_probe_rows(args), PROBE_SEED, device)
            bm, pm = torch.isin(build["contig"], mine), torch.isin(probe["contig"], mine)
            build = {k: v[bm].contiguous() for k, v in build.items()}
            probe = {k: v[pm].contiguous() for k, v in probe.items()}
            info = {"contigs": [int(c) for c in mine.tolist()]}
            name = (f"cfg5 100Mx100M hg38-weighted U{{50..150}}, contig-sharded (PartitionMode::Partitioned): {world} ranks, "
                    f"LPT contig assignment, each rank builds and probes its contigs only")
        else:
            first, n = probe_range(args, rank, world)
            probe = synth.counter_side_torch(n, PROBE_SEED, device, first_row=first)
            if getattr(args, "scaling", "weak") == "weak":
                name = (f"cfg5 100Mx100M hg38-weighted U{{50..150}}: {args.build_rows} build rows replicated per GPU, "
                        f"{args.shard_rows}-probe shard per GPU (N=8 == the full config)")
            else:
                name = (f"cfg5 100Mx100M hg38-weighted U{{50..150}}: {args.build_rows} build rows replicated per GPU, "
                        f"{total_probe_rows(args)} probe rows in total split over the GPUs (strong scaling)")
        if getattr(args, "probe_order", "random") == "sorted":
            probe = sort_by_position(probe)
            name += ", probe rows position-sorted (contig, start)"
        return build, probe, name, info
    b, p = synth.CONFIGS[args.workload]()
    names = {"cfg2": "cfg2 1Mx1M 24 contigs x 10Mbp uniform U{50..150}",
             "cfg3": "cfg3 databio-shaped 1.2M build x 10M probe, skewed lengths",
             "cfg4": "cfg4 high fan-out 1Mx1M, build widths U{100k..500k}"}
    if world > 1:  # weak scaling: every rank probes its own permutation of the probe side
        rng = np.random.default_rng(rank + 1)
        perm = rng.permutation(len(p["key"]))
        p = {k: v[perm] for k, v in p.items()}
    return to_device(b, device), to_device(p, device), names[args.workload], {}


def host_digest(left, right_global):
    """order-independent multiset digest of (left, right) pairs = what sq_pairs_digest_device computes"""
    l = np.asarray(left, dtype=np.uint64)
    r = np.asarray(right_global, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (l << np.uint64(32)) | (r & np.uint64(0xFFFFFFFF))
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
        return int(l.size), int(np.add.reduce(x, dtype=np.uint64)) if x.size else 0


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
def sample_contigs(args):
    from sequila_native_b200 import synth
    return [int(c) for c in np.argsort(synth.HG38)[:args.cpu_sample_contigs]]


def cpu_sample_cfg5(args):
    """Bounded sample of the cfg5 workload for the CPU arm, generated on the host by the same counter-based
    generator the CUDA arm uses on the device (identical rows): every build row of the `cpu_sample_contigs`
    smallest contigs, and rank 0's probe rows of those contigs capped at --cpu-sample-probes (same density and
    fan-out as the whole workload: the join never crosses contigs)."""
    from sequila_native_b200 import synth
    sel = sample_contigs(args)
    b = synth.counter_side(args.build_rows, BUILD_SEED, keep_contigs=sel)
    p = synth.counter_side(args.shard_rows, PROBE_SEED, keep_contigs=sel)
    cap = args.cpu_sample_probes
    return ({k: b[k] for k in ("key", "start", "end")}, {k: p[k][:cap] for k in ("key", "start", "end")},
            f"contigs {sorted(sel)} of the workload")


def cpu_baseline(b, p, what, threads, min_seconds=1.0, index=None):
    """Times the reference's CPU path (oracle/: C++ restatement of coitrees 0.4.0 AVX2 tree +
    interval_join.rs probe loop) on the box's host cores: 8192-row probe batches dealt round-robin to
    `threads` threads over one shared index (= DataFusion CollectLeft with target_partitions = threads).
    Passes over the sample are repeated until `min_seconds` of wall time (x threads = CPU work)."""
    from oracle import oracle as O
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
        build_h, probe_h, what = cpu_sample_cfg5(args)
        name = (f"cfg5 100Mx100M hg38-weighted U{{50..150}}: {args.build_rows} build rows replicated per GPU, "
                f"{args.shard_rows}-probe shard per GPU (N=8 == the full config)")
    else:
        build_h, probe_h = synth.CONFIGS[args.workload]()
        name, what = args.workload, "whole build side"
    vals = []
    base = None
    index = None
    one_thread = None
    ref_si = None
    t_all = time.time()
    for it in range(args.warmup + args.steps):  # one step = one timed pass set over the bounded sample
        base = cpu_baseline(build_h, probe_h, what, threads, min_seconds=0.5, index=index)
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


# ------------------------------------------------------------------------------------------------
# pieces of the CUDA arm
# ------------------------------------------------------------------------------------------------
class Tiles:
    """this rank's probe rows cut into launches of at most `rows` rows (one launch = one probe tile)"""

    def __init__(self, probe, rows):
        n = probe["key"].numel()
        self.bounds = [(lo, min(lo + rows, n)) for lo in range(0, max(n, 1), rows)]
        self.cols = [tuple(probe[k][lo:hi] for k in ("key", "start", "end")) for lo, hi in self.bounds]


def timed_steps(torch, fn, steps, flush):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        if flush is not None:
            flush.fill_(1)  # L2 flush between timed iterations (outside the event pair)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in evs]))


def oracle_check(sn, st, ctx, build, probe, tiles, left, right, run_tile):
    """Every rank compares its CUDA output with the oracle on the rows of CHECK_CONTIGS (5 % of the workload):
    per-row counts and the digest of the emitted pairs restricted to those probe rows."""
    import torch
    from oracle import oracle as O  # checker only (allowed here: the comparison is not timed and not shipped)
    dev = probe["key"].device
    sel = torch.tensor(CHECK_CONTIGS, dtype=torch.int32, device=dev)
    bm = torch.isin(build["contig"], sel)
    brow = torch.nonzero(bm).squeeze(1)
    bh = {k: build[k][bm].cpu().numpy() for k in ("key", "start", "end")}
    oidx = O.OracleIndex(bh["key"].view(np.uint64), bh["start"], bh["end"])
    brow_h = brow.cpu().numpy().astype(np.uint64)
    rows = pairs = 0
    ok_counts = ok_digest = True
    for (lo, hi), cols in zip(tiles.bounds[:2], tiles.cols[:2]):  # at most two launches (25M probe rows) per rank
        n_pairs = run_tile(cols)
        pm = torch.isin(probe["contig"][lo:hi], sel)
        prow = torch.nonzero(pm).squeeze(1)
        ph = {k: probe[k][lo:hi][pm].cpu().numpy() for k in ("key", "start", "end")}
        ol, orr, oc = oidx.probe(ph["key"].view(np.uint64), ph["start"], ph["end"])
        got_counts = torch.from_numpy(st.counts().astype(np.int32)).to(dev)[prow].cpu().numpy().astype(np.uint32)
        ok_counts &= bool(np.array_equal(got_counts, oc))
        keep = pm[right[:n_pairs].long()]
        l_sel, r_sel = left[:n_pairs][keep].contiguous(), right[:n_pairs][keep].contiguous()
        dg = st.digest_device(l_sel, r_sel, int(l_sel.numel()))
        want = host_digest(brow_h[ol], prow.cpu().numpy().astype(np.uint64)[orr])
        ok_digest &= (dg[0], dg[1]) == want
        rows += int(pm.sum())
        pairs += len(ol)
    return {"contigs": list(CHECK_CONTIGS), "probe_rows_checked": rows, "pairs_checked": pairs,
            "counts_equal_oracle": ok_counts, "pair_digest_equals_oracle": ok_digest,
            "build_rows_in_oracle_index": int(len(bh["key"]))}


def e2e_pipeline(sn, ctx, idx, probe_h, n_pairs_expect, args, T, flags, steps, barrier):
    """End to end through the host C ABI, asynchronous tile pipeline (sq_stream_submit / sq_stream_collect): pinned
    host probe columns in, pinned host (left_idx, counts) out, fresh output buffers from the library's pool for every
    tile, no pre-pass, `cuda_pipeline_depth` tiles in flight per partition.  T host threads = DataFusion partitions
    over one shared index (PartitionMode::CollectLeft, interval_join.rs:473-487), run by the library's native partition
    loop (sq_driver_run, include/sequila_driver.h: what a compiled host does; a Python loop holds the GIL for
    ~100 us per tile, which at 64 tiles per step is most of the step).  The wire format is verified once, untimed, through
    the Python binding of the same two calls: the host-side pairs digest must equal the device-side one."""
    from sequila_native_b200.cuda_join import CudaDriver
    drv = CudaDriver(ctx, T)
    n_probe = len(probe_h["key"])
    n_tiles = max(T, (args.e2e_tiles // T) * T)
    hk, hs, he = probe_h["key"], probe_h["start"], probe_h["end"]
    digest = None
    if not flags & 1 and probe_h.get("ids") is None:
        bounds = np.linspace(0, n_probe, 9).astype(np.int64)
        st = sn.CudaStream(ctx)
        depth = int(ctx.get_option("cuda_pipeline_depth"))
        pend, cnt, tot = [], 0, 0

        def fold(tk, t):
            nonlocal cnt, tot
            n, left, right, counts = st.collect(tk)
            r = np.repeat(np.arange(int(bounds[t]), int(bounds[t + 1]), dtype=np.uint64), counts)
            c, s_ = host_digest(left, r)
            cnt += c
            tot = (tot + s_) & ((1 << 64) - 1)
        for t in range(8):
            if len(pend) == depth:
                fold(*pend.pop(0))
            lo, hi = int(bounds[t]), int(bounds[t + 1])
            pend.append((st.submit(idx, hk[lo:hi], hs[lo:hi], he[lo:hi], flags), t))
        for tk, t in pend:
            fold(tk, t)
        digest = (cnt, tot)
        del st
    for _ in range(2):
        got = drv.run(idx, hk, hs, he, n_tiles, flags)
        assert got["n_pairs"] == n_pairs_expect, (got, n_pairs_expect)
    ids = probe_h.get("ids")

    def one_pass():
        if ids is not None:  # 12 bytes per probe row: the contig column as dictionary ids (sq_stream_submit_ids)
            return drv.run_ids(idx, probe_h["dict"], ids, hs, he, n_tiles, flags)
        return drv.run(idx, hk, hs, he, n_tiles, flags)
    if ids is not None:
        assert one_pass()["n_pairs"] == n_pairs_expect
    barrier()
    e0 = time.perf_counter()
    agg = {"h2d_ms": 0.0, "kernel_ms": 0.0, "d2h_ms": 0.0, "tiles": 0, "regrown": 0, "native_seconds": 0.0, "d2h_bytes": 0, "h2d_bytes": 0}
    for _ in range(steps):
        r = one_pass()
        for k, f in (("h2d_ms", "h2d_ms"), ("kernel_ms", "kernel_ms"), ("d2h_ms", "d2h_ms"), ("tiles", "n_tiles"),
                     ("regrown", "regrown_tiles"), ("native_seconds", "seconds"), ("d2h_bytes", "d2h_bytes"), ("h2d_bytes", "h2d_bytes")):
            agg[k] += r[f]
    ms = (time.perf_counter() - e0) * 1e3 / steps
    return ms, n_tiles, digest, agg


def sub_config(sn, torch, ctx, name, device, hbm_peak, steps, flush):
    """one of the other synthetic configs on this GPU: kernel-only step, roofline fraction, pair digest"""
    b, p = sn.synth.CONFIGS[name]()
    bd, pd = to_device(b, device), to_device(p, device)
    ts = torch.cuda.current_stream().cuda_stream
    idx = sn.CudaIndex.build_device(ctx, bd["key"], bd["start"], bd["end"], ts)
    st = sn.CudaStream(ctx, cuda_stream=ts)
    n_pairs = st.probe_count_device(idx, pd["key"], pd["start"], pd["end"])
    left = torch.empty(max(n_pairs, 1), dtype=torch.int32, device=device)
    right = torch.empty(max(n_pairs, 1), dtype=torch.int32, device=device)

    def step():
        return st.probe_join_device(idx, pd["key"], pd["start"], pd["end"], left, right)
    for _ in range(3):
        assert step() == n_pairs
    ms = timed_steps(torch, step, steps, flush)
    dg = st.digest_device(left, right, n_pairs)
    n_probe = len(p["key"])
    bts = 16.0 * n_probe + 12.0 * n_pairs
    c_ms = timed_steps(torch, lambda: st.probe_count_device(idx, pd["key"], pd["start"], pd["end"]), steps, flush)
    return {"probe_rows": n_probe, "build_rows": len(b["key"]), "pairs": n_pairs, "ms_per_step": ms,
            "value": n_probe / (ms * 1e-3), "pairs_per_s": n_pairs / (ms * 1e-3), "unit": "probe intervals/s",
            "kernels": "k_probe_packed" if idx.uses_packed else ("k_probe_rank" if idx.uses_rank else "k_probe_count+k_tile_scan+k_probe_write"),
            "roofline_frac": bts / (ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": bts,
            "count_only_ms": c_ms, "build_ms": idx.build_ms, "digest": {"pairs": dg[0], "sum": dg[1]}}


def exec_node_line(sn, args, n_build=2_000_000, n_probe=2_000_000, batch_rows=8192, build_ids=None, options=None):
    """The layer a DataFusion user hits (IntervalJoinExec over Arrow RecordBatches, interval_join.rs:1192-1233,
    1580-1640): cfg5-shaped rows at 2 % scale, Utf8 contig column, probe side fed in 8192-row batches (DataFusion's
    default batch size), the six joined columns out."""
    import pyarrow as pa
    from sequila_native_b200 import intervals as IV
    from sequila_native_b200.interval_join import HashJoinDesc, optimize
    from sequila_native_b200 import synth
    scale = n_build / BUILD_ROWS
    L_len = np.maximum((synth.HG38 * scale).astype(np.int64), 1000)
    b = synth.counter_side(n_build, BUILD_SEED, lengths=L_len)
    p = synth.counter_side(n_probe, PROBE_SEED, lengths=L_len)
    names = np.array(synth.CONTIG_NAMES)
    cols = ["contig", "pos_start", "pos_end"]

    def table(s):
        # `take` of the 24 names: pa.array() of a big numpy string array comes back chunked
        contig = pa.array(names.tolist(), pa.string()).take(pa.array(s["contig"].astype(np.int32)))
        return pa.record_batch([contig, pa.array(s["start"]), pa.array(s["end"])], names=cols)
    L, R = table(b), table(p)
    cfg = sn.SequilaConfig()
    sn.apply_set(cfg, "SET sequila.interval_join_algorithm TO cuda")
    if build_ids:  # A/B runs (tools/time_exec_ids.py); the node's own default is position ids
        sn.apply_set(cfg, f"SET sequila.cuda_build_ids TO {build_ids}")
    for k, v in (options or {}).items():
        sn.apply_set(cfg, f"SET sequila.{k} TO {v}")
    f = IV.parse_condition_sql("a.pos_start <= b.pos_end AND a.pos_end >= b.pos_start", "a", cols, "b", cols)
    plan = optimize(HashJoinDesc(L.schema, R.schema, [("contig", "contig")], f), cfg)
    t0 = time.perf_counter()
    plan.collect_build([L])
    build_s = time.perf_counter() - t0
    batches = [R.slice(i, batch_rows) for i in range(0, R.num_rows, batch_rows)]
    best = None
    rows = 0
    lib_ms = None
    for _ in range(3):
        j0 = plan.metrics().join_time_ns
        t0 = time.perf_counter()
        rows = 0
        for out in plan.probe_batches(batches):
            rows += out.num_rows
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, lib_ms = dt, (plan.metrics().join_time_ns - j0) / 1e6
    # the same probe side dealt to 4 partitions of the one node (DataFusion's default is one partition per core; the
    # reference's own benchmarks set target_partitions = 1): 4 host threads, one sq_stream each, one shared index
    import concurrent.futures as cf
    P = 4
    parts = [batches[p::P] for p in range(P)]

    def drive(p):
        return sum(out.num_rows for out in plan.probe_batches(parts[p], partition=p))
    best_p = None
    with cf.ThreadPoolExecutor(P) as pool:
        for _ in range(3):
            t0 = time.perf_counter()
            rows_p = sum(pool.map(drive, range(P)))
            dt = time.perf_counter() - t0
            best_p = dt if best_p is None else min(best_p, dt)
    assert rows_p == rows
    plan.close()
    return {"api": "IntervalJoinExec (sq_exec_* over the Arrow C Data Interface), one host thread", "batch_rows": batch_rows,
            "probe_rows": n_probe, "build_rows": n_build, "output_rows": rows, "seconds": best,
            "value": n_probe / best, "unit": "probe intervals/s", "output_rows_per_s": rows / best,
            "columns_out": 6, "key_column": "Utf8",
            "partitions_4": {"seconds": best_p, "value": n_probe / best_p, "output_rows_per_s": rows / best_p},
            "library_ms": lib_ms, "collect_build_seconds": build_s,
            "note": "seconds = wall time of the Python host loop of the best one-thread pass (244 pushes through pyarrow's C Data "
                    "export); library_ms = join_time of the node's metrics for that pass (utils.rs:441-495): concat + hash + cast + "
                    "probe + take inside the library"}


def link_ceiling():
    """what the host link of a 1-GPU box of this pool moves with pinned memory (profiles/r02_pcie_ubench.json, 64 MiB copies,
    tools/ubench_pcie.py): the ceiling the end-to-end lines are read against"""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_pcie_ubench.json")) as f:
            r = json.load(f)["64MiB"]
        return {"h2d_alone": r["h2d_GBps"], "d2h_alone": r["d2h_GBps"], "each_way_under_duplex_load": r["duplex_each_GBps"],
                "source": "profiles/r02_pcie_ubench.json (tools/ubench_pcie.py, 64 MiB pinned copies)"}
    except Exception:
        return None


def source_sha(files):
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(ROOT, "sequila_native_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


KERNEL_SOURCES = {"k_probe_packed": ["sq_probe_packed.cu", "sq_packed_common.cuh", "sq_probe_common.cuh"],
                  "k_probe_staged": ["sq_probe_staged.cu", "sq_packed_common.cuh", "sq_probe_common.cuh"],
                  "k_probe_rank": ["sq_probe_rank.cu", "sq_soa_common.cuh", "sq_probe_common.cuh"],
                  "k_probe_soa": ["sq_probe.cu", "sq_soa_common.cuh", "sq_probe_common.cuh"]}


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json), valid only while
    the kernel's sources are byte-identical to the ones that were profiled"""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel)
        if not isinstance(rec, dict):
            return None, "no ncu capture recorded for this kernel"
        srcs = [f for f in KERNEL_SOURCES.get(kernel, []) if os.path.exists(os.path.join(ROOT, "sequila_native_b200", "csrc", f))]
        if rec.get("source_sha") != source_sha(srcs):
            return None, f"stale: the capture {rec.get('capture')} was taken on other kernel sources"
        return float(rec["dram_bytes"]), rec.get("capture")
    except Exception as ex:
        return None, repr(ex)


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
    from sequila_native_b200 import _native as N

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the cuda interval join has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    # NUMA locality of the pinned staging buffers: run this rank on the CPUs next to its GPU (what a
    # launcher with --bind-to would do); the CPU baseline below widens the mask again
    all_cpus = os.sched_getaffinity(0)
    near = set(all_cpus)
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
            near = set(all_cpus)
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"

    ctx = sn.CudaContext(local_rank)
    build, probe, wname, winfo = make_workload(args, rank, world, device)
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
    tiles = Tiles(probe, args.shard_rows)
    tile_pairs = [st.probe_count_device(idx, *c) for c in tiles.cols]
    n_pairs = sum(tile_pairs)
    cap = max(max(tile_pairs), 1)
    left = torch.empty(cap, dtype=torch.int32, device=device)
    right = torch.empty(cap, dtype=torch.int32, device=device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2

    def run_tile(cols):
        # search -> count -> look-back scan -> write as ONE fused kernel pass (sq_probe_join_device)
        return st.probe_join_device(idx, cols[0], cols[1], cols[2], left, right)

    def step():
        return sum(run_tile(c) for c in tiles.cols)

    for _ in range(max(args.warmup, 3)):
        assert step() == n_pairs
    torch.cuda.synchronize()

    # parity guard inside the bench: the emitted pairs of every launch are digested on the device (order-independent
    # multiset digest, right_idx made global) and — below — compared with the oracle on a contig sample, on EVERY rank
    dsum = dcount = 0
    for (lo, hi), c, tp in zip(tiles.bounds, tiles.cols, tile_pairs):
        assert run_tile(c) == tp
        d = st.digest_device(left, right, tp, right_offset=lo)
        dcount += d[0]
        dsum = (dsum + d[1]) & ((1 << 64) - 1)
    assert dcount == n_pairs
    check = None
    if args.workload == "cfg5_shard" and not args.no_oracle_check:
        check = oracle_check(sn, st, ctx, build, probe, tiles, left, right, run_tile)
        assert check["counts_equal_oracle"] and check["pair_digest_equals_oracle"], check

    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    st.set_profiling(True)
    launches0 = st.launches
    t0 = time.time()
    step_ms = timed_steps(torch, step, args.steps, flush)
    barrier()
    t1 = time.time()
    phases = st.phase_ms()
    st.set_profiling(False)
    launches = st.launches - launches0
    clocks = sampler.stop(t0, t1) if sampler else None

    from sequila_native_b200.sharding import reduce_step, gather_ranks
    step_ms_max, probes_total, pairs_total = reduce_step(step_ms, n_probe, n_pairs, device)
    value = probes_total / (step_ms_max * 1e-3)
    per_rank = gather_ranks({"rank": rank, "probe_rows": n_probe, "build_rows": n_build, "pairs": n_pairs, "ms_per_step": step_ms,
                             "build_ms": build_best, "index_bytes": index_bytes, "oracle_check": check, **winfo}, device)

    # ---- end to end through the host C ABI: pinned host inputs -> H2D -> kernels -> D2H pairs, asynchronous tiles
    T = max(1, args.e2e_partitions)
    e2e_rows = min(n_probe, args.shard_rows)  # one launch worth of this rank's rows
    probe_h = {"key": ctx.pinned_copy(probe["key"][:e2e_rows].cpu().numpy().view(np.uint64)),
               "start": ctx.pinned_copy(probe["start"][:e2e_rows].cpu().numpy()),
               "end": ctx.pinned_copy(probe["end"][:e2e_rows].cpu().numpy())}
    e_pairs = tile_pairs[0]
    e_ms, n_tiles, e_digest, e_stats = e2e_pipeline(sn, ctx, idx, probe_h, e_pairs, args, T, 0, args.e2e_steps, barrier)
    torch.cuda.synchronize()
    d0 = st.digest_device(left, right, run_tile(tiles.cols[0]), right_offset=0)
    assert e_digest == (d0[0], d0[1]), "host-side pairs of the pipeline differ from the device-side pairs"
    e_ms_max, e_rows_total, _ = reduce_step(e_ms, e2e_rows, 0, device)
    e2e_value = e_rows_total / (e_ms_max * 1e-3)
    # the same with the counts as one byte per probe row (SQ_TILE_COUNTS_U8: every count of a cfg5 tile fits; a tile with a
    # bigger one falls back to 4-byte counts by itself)
    eu_ms, _, eu_digest, eu_stats = e2e_pipeline(sn, ctx, idx, probe_h, e_pairs, args, T, N.TILE_COUNTS_U8, args.e2e_steps, barrier)
    assert eu_digest == (d0[0], d0[1]), "host-side pairs of the pipeline (byte counts) differ from the device-side pairs"
    eu_ms_max, _, _ = reduce_step(eu_ms, 0, 0, device)
    # the same as `select count(1)`: host columns in, one number per tile out — the query shape of the reference's own
    # benchmarks (queries/q1-coitrees.sql:16-19); only the H2D copy is left on the wire
    ec_ms, _, _, ec_stats = e2e_pipeline(sn, ctx, idx, probe_h, e_pairs, args, T, N.TILE_COUNT_ONLY | N.TILE_NO_COUNTS,
                                         args.e2e_steps, barrier)
    ec_ms_max, _, _ = reduce_step(ec_ms, 0, 0, device)
    # ... and with the key column as 4-byte dictionary ids (12 bytes per probe row on the wire), join and count(1)
    e_ids = None
    if args.workload == "cfg5_shard":
        probe_h["ids"] = ctx.pinned_copy(probe["contig"][:e2e_rows].cpu().numpy().astype(np.uint32))
        probe_h["dict"] = sn.synth.key_hash(np.arange(24))
        ei_ms, _, _, _ = e2e_pipeline(sn, ctx, idx, probe_h, e_pairs, args, T, N.TILE_COUNT_ONLY | N.TILE_NO_COUNTS, args.e2e_steps, barrier)
        ej_ms, _, _, _ = e2e_pipeline(sn, ctx, idx, probe_h, e_pairs, args, T, 0, args.e2e_steps, barrier)
        ej_ms_max, _, _ = reduce_step(ej_ms, 0, 0, device)
        ei_ms_max, _, _ = reduce_step(ei_ms, 0, 0, device)
        e_ids = {"api": "sq_stream_submit_ids: key column as u32 ids into a dictionary of key hashes, 12 B per probe row in",
                 "value": e_rows_total / (ej_ms_max * 1e-3), "unit": "probe intervals/s", "ms_per_step": ej_ms_max,
                 "h2d_bytes_per_step": 12 * e2e_rows,
                 "count_only": {"value": e_rows_total / (ei_ms_max * 1e-3), "unit": "probe intervals/s", "ms_per_step": ei_ms_max,
                                "h2d_bytes_per_step": 12 * e2e_rows, "link_GBps": {"h2d": 12 * e2e_rows / (ei_ms * 1e-3) / 1e9}}}
    del probe_h

    # ---- count only: `select count(1) from a join b on ...`; no pair is written
    count_only = None
    if not args.no_count_only:
        c0 = tiles.cols[0]
        for _ in range(3):
            assert st.probe_count_device(idx, *c0) == tile_pairs[0]
        c_ms = timed_steps(torch, lambda: st.probe_count_device(idx, *c0), min(args.steps, 10), flush)
        nr = c0[0].numel()
        count_only = {"query": "count(1) of the join", "avg_launch_ms": c_ms, "value": nr / (c_ms * 1e-3),
                      "unit": "probe intervals/s", "algorithmic_bytes": 16.0 * nr + 4.0 * tile_pairs[0],
                      "roofline_frac": (16.0 * nr + 4.0 * tile_pairs[0]) / (c_ms * 1e-3) / 1e9 / hbm_peak}

    # ---- materialise (process_probe_batch's `take` per output column, interval_join.rs:1620-1632): the six output
    # columns of SURVEY §8(d) — contig (dictionary id, int32), pos_start, pos_end of both sides — gathered on the
    # device from the pairs of the last launch; B_gather = 8 B pair read + 2 x 4 B per column = 56 B per pair.
    # Headline: an index built with position ids (cuda_build_ids = positions, what the exec node does): payload in the
    # index's sorted order, the hits of a probe row are neighbouring rows of the row-wise pack.  Beside it the same
    # gathers with row ids (payload in build-row order: one random row per pair) and `take` column by column.
    materialise = None
    if not args.no_materialise and rank == 0:
        c0 = tiles.cols[0]
        lo0, hi0 = tiles.bounds[0]
        np0 = tile_pairs[0]
        pvals = [probe[k][lo0:hi0] for k in ("contig", "start", "end")]
        ctx.set_option("cuda_build_ids", "positions")
        idx_p = sn.CudaIndex.build_device(ctx, build["key"], build["start"], build["end"], tstream)
        ctx.set_option("cuda_build_ids", "rows")
        st_p = sn.CudaStream(ctx, cuda_stream=tstream)
        left_p = torch.empty(max(np0, 1), dtype=torch.int32, device=device)

        def prepare(index, stream, lbuf):
            assert stream.probe_join_device(index, c0[0], c0[1], c0[2], lbuf, right) == np0
            cols = [index.add_column_device(build[k]) for k in ("contig", "start", "end")]
            pack = index.pack_columns(cols)  # the three build columns row-wise: one read per pair serves all
            outs = [torch.empty(max(np0, 1), dtype=torch.int32, device=device) for _ in range(6)]

            def gather_step():  # two launches: build pack, probe columns
                stream.gather_pack_device(pack, outs[:3])
                stream.gather_probe_columns_device(pvals, outs[3:])

            def take_step():  # `take` column by column, what the reference's loop does (six launches)
                for c, o in zip(cols, outs[:3]):
                    stream.gather_build_device(c, o)
                for v, o in zip(pvals, outs[3:]):
                    stream.gather_probe_device(v, o)

            def pack_only():
                stream.gather_pack_device(pack, outs[:3])
            return gather_step, take_step, pack_only, outs

        res = {}
        for mode, (index, stream, lbuf) in (("rows", (idx, st, left)), ("positions", (idx_p, st_p, left_p))):
            gather_step, take_step, pack_only, outs = prepare(index, stream, lbuf)
            for fn in (take_step, gather_step):
                fn()
            take_ms = timed_steps(torch, take_step, min(args.steps, 10), flush)
            pack_ms = timed_steps(torch, pack_only, min(args.steps, 10), flush)
            g_ms = timed_steps(torch, gather_step, min(args.steps, 10), flush)
            res[mode] = {"ms": g_ms, "take_per_column_ms": take_ms, "build_pack_ms": pack_ms, "outs": outs}
        # rows mode against torch indexing (same device data); positions mode against rows mode, every value
        stride = max(np0 // 100000, 1)
        li = left[:np0].long()[::stride]
        assert torch.equal(res["rows"]["outs"][1][:np0][::stride], build["start"][li])
        ri = right[:np0].long()[::stride]
        assert torch.equal(res["rows"]["outs"][5][:np0][::stride], pvals[2][ri])
        for a_, b_ in zip(res["rows"]["outs"], res["positions"]["outs"]):
            assert torch.equal(a_[:np0], b_[:np0])
        g_bytes = 56.0 * np0
        g_ms = res["positions"]["ms"]
        frac = lambda ms: g_bytes / (ms * 1e-3) / 1e9 / hbm_peak if ms else None
        materialise = {"columns": 6, "ms": g_ms, "pairs_per_s": np0 / (g_ms * 1e-3) if g_ms else None,
                       "algorithmic_bytes": g_bytes, "roofline_frac": frac(g_ms), "launches": 2,
                       "how": "index built with position ids (sequila.cuda_build_ids = positions, the exec node's default): build "
                       "columns kept in the index's sorted order and packed row-wise (16 B per build row), the hits of a probe "
                       "row are neighbouring rows; probe columns in one pass over right_idx",
                       "build_pack_ms": res["positions"]["build_pack_ms"], "take_per_column_ms": res["positions"]["take_per_column_ms"],
                       "values_equal_row_ids": True,
                       "row_ids": {"ms": res["rows"]["ms"], "roofline_frac": frac(res["rows"]["ms"]),
                                   "build_pack_ms": res["rows"]["build_pack_ms"], "take_per_column_ms": res["rows"]["take_per_column_ms"],
                                   "how": "payload in build-row order: one random 16-byte read per pair"}}
        del res, idx_p, st_p, left_p
        assert run_tile(c0) == np0  # the sections below read left / right of the first launch

    # ---- same rows in position order (what BAM / BED inputs look like), N=1 only
    locality = None
    if world == 1 and not args.no_locality and args.probe_order == "random":
        lo0, hi0 = tiles.bounds[0]
        sp = sort_by_position({k: v[lo0:hi0] for k, v in probe.items()})

        def sorted_step():
            return st.probe_join_device(idx, sp["key"], sp["start"], sp["end"], left, right)
        for _ in range(3):
            assert sorted_step() == tile_pairs[0]
        sd = st.digest_device(left, right, tile_pairs[0])
        s_ms = timed_steps(torch, sorted_step, min(args.steps, 10), flush)
        locality = {"probe_order": "position-sorted (contig, start)", "avg_launch_ms": s_ms,
                    "value": (hi0 - lo0) / (s_ms * 1e-3), "unit": "probe intervals/s", "pairs": sd[0],
                    "roofline_frac": (16.0 * (hi0 - lo0) + 12.0 * tile_pairs[0]) / (s_ms * 1e-3) / 1e9 / hbm_peak}
        del sp

    # ---- the full config on ONE GPU: all 100M probe rows against the 100M-row index, 12.5M-row launches
    full = None
    if (world == 1 and not args.no_full_config and args.workload == "cfg5_shard" and args.scaling == "weak"
            and args.parallelism == "replicated"):
        tot = total_probe_rows(args)
        pf = sn.synth.counter_side_torch(tot, PROBE_SEED, device)
        ft = Tiles(pf, args.shard_rows)
        f_pairs = [st.probe_count_device(idx, *c) for c in ft.cols]
        if max(f_pairs) > left.numel():
            left = torch.empty(max(f_pairs), dtype=torch.int32, device=device)
            right = torch.empty(max(f_pairs), dtype=torch.int32, device=device)

        def full_step():
            return sum(st.probe_join_device(idx, c[0], c[1], c[2], left, right) for c in ft.cols)
        assert full_step() == sum(f_pairs)
        f_ms = timed_steps(torch, full_step, 3, flush)
        full = {"probe_rows": tot, "build_rows": n_build, "pairs": sum(f_pairs), "launches": len(ft.cols), "ms": f_ms,
                "value": tot / (f_ms * 1e-3), "pairs_per_s": sum(f_pairs) / (f_ms * 1e-3), "unit": "probe intervals/s",
                "roofline_frac": (16.0 * tot + 12.0 * sum(f_pairs)) / (f_ms * 1e-3) / 1e9 / hbm_peak, "probe_order": "random"}
        # ... and position-sorted (one global sort of the probe side, as a BAM / BED file would arrive)
        ps = sort_by_position(pf)
        del pf, ft
        fs = Tiles(ps, args.shard_rows)
        s_pairs = [st.probe_count_device(idx, *c) for c in fs.cols]
        assert sum(s_pairs) == sum(f_pairs)
        if max(s_pairs) > left.numel():
            left = torch.empty(max(s_pairs), dtype=torch.int32, device=device)
            right = torch.empty(max(s_pairs), dtype=torch.int32, device=device)

        def full_sorted_step():
            return sum(st.probe_join_device(idx, c[0], c[1], c[2], left, right) for c in fs.cols)
        assert full_sorted_step() == sum(s_pairs)
        fs_ms = timed_steps(torch, full_sorted_step, 3, flush)
        full["position_sorted"] = {"ms": fs_ms, "value": tot / (fs_ms * 1e-3), "pairs_per_s": sum(s_pairs) / (fs_ms * 1e-3),
                                   "roofline_frac": (16.0 * tot + 12.0 * sum(s_pairs)) / (fs_ms * 1e-3) / 1e9 / hbm_peak}
        del ps, fs

    sub = None
    if world == 1 and not args.no_sub_configs and args.workload == "cfg5_shard":
        del left, right
        sub = {}
        for name in ("cfg2", "cfg3", "cfg4"):
            sub[name] = sub_config(sn, torch, ctx, name, device, hbm_peak, min(args.steps, 10), flush)

    exec_line = None
    if world == 1 and not args.no_exec:
        try:
            exec_line = exec_node_line(sn, args)
        except Exception as ex:  # pyarrow missing on a box: the line is explanatory
            exec_line = {"unavailable": repr(ex)}

    if rank == 0:
        # ---- roofline of the dominant kernel (algorithmic bytes) -----------------
        # B_probe of SURVEY.md §8(d) with u64 key hashes consumed on the device:
        # 16 B per probe row (key hash 8 + start 4 + end 4) + 12 B per emitted pair
        # (read the hit's build row id 4, write (left,right) 8).  One launch = one probe tile.
        dom = "k_probe_packed" if idx.uses_packed else ("k_probe_rank" if idx.uses_rank else "k_probe_soa")
        rows0 = tiles.bounds[0][1] - tiles.bounds[0][0]
        n_launch = len(tiles.cols)
        b_dom = (16.0 * n_probe + 12.0 * n_pairs) / n_launch
        t_dom = phases["join"]
        achieved = b_dom / (t_dom * 1e-3) / 1e9 if t_dom > 0 else 0.0
        traffic, traffic_note = ncu_traffic(dom)
        result = {
            "metric": "probe_intervals_per_s", "value": value, "unit": "probe intervals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms_max,
            "higher_is_better": True, "scaling": args.scaling if args.parallelism == "replicated" else "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": wname, "build_rows": n_build, "probe_rows_per_gpu": n_probe,
                       "pairs_per_gpu": n_pairs, "launches_per_step": n_launch, "rows_per_launch": rows0,
                       "l2": "256 MiB flush write between timed steps; inputs also > L2",
                       "generator": "counter-based (synth.counter_side*): identical rows on every arm and rank",
                       "parallelism": (f"probe shards x{world}, build index replicated, no collective on the data path"
                                       if args.parallelism == "replicated" else
                                       f"contigs LPT-sharded over {world} ranks (Partitioned), no collective on the data path")},
            "pairs_per_s": pairs_total / (step_ms_max * 1e-3),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": b_dom, "avg_launch_ms": t_dom,
                         # DRAM bytes the launch really moves (ncu) over its duration: how busy HBM is
                         "traffic_rate": (traffic / (t_dom * 1e-3) / 1e9) if traffic and t_dom > 0 else None,
                         "traffic_frac": (traffic / (t_dom * 1e-3) / 1e9 / hbm_peak) if traffic and t_dom > 0 else None},
            "build": {"ms": build_best, "rows_per_s": n_build / (build_best * 1e-3) if build_best else None,
                      "roofline_frac": (24.0 * n_build / (build_best * 1e-3) / 1e9 / hbm_peak) if build_best else None,
                      "index_bytes": index_bytes, "keys": idx.keys},
            "e2e": {"value": e2e_value, "unit": "probe intervals/s", "h2d_bytes_per_step": 16 * e2e_rows,
                    "d2h_bytes_per_step": 4 * e_pairs + 4 * e2e_rows + 16 * n_tiles,
                    # what the pipeline really moved (the speculative copy-out of a tile is sized 2 % above the previous
                    # tile's fan-out: a little more than the pairs there are)
                    "d2h_bytes_moved_per_step": e_stats["d2h_bytes"] // max(args.e2e_steps, 1),
                    "wire": "in: key hash u64 + start/end i32 per probe row; out: left_idx u32 per pair + per-row counts u32 "
                            "(= rle_right, interval_join.rs:1604; right_idx is their run-length expansion and does not cross PCIe)",
                    "ms_per_step": e_ms_max, "rows_per_step_per_gpu": e2e_rows, "steps": args.e2e_steps,
                    "api": "sq_stream_submit / sq_stream_collect (host C ABI) driven by sq_driver_run (native partition threads): "
                           "pinned host inputs, fresh pinned outputs from the library's pool per tile, no pre-pass",
                    "partitions": T, "tiles": n_tiles,
                    "pipeline_depth": int(ctx.get_option("cuda_pipeline_depth")),
                    "phase_ms": {"h2d_sum": e_stats["h2d_ms"], "kernels_sum": e_stats["kernel_ms"], "d2h_sum": e_stats["d2h_ms"],
                                 "tiles": e_stats["tiles"], "regrown": e_stats["regrown"]},
                    "link_GBps": {"h2d": 16 * e2e_rows / (e_ms * 1e-3) / 1e9,
                                  "d2h": (4 * e_pairs + 4 * e2e_rows) / (e_ms * 1e-3) / 1e9},
                    "pairs_digest_equals_device": True, "key_ids": e_ids,
                    "link_ceiling_GBps": link_ceiling(),
                    "counts_u8": {"api": "the same with SQ_TILE_COUNTS_U8: per-row counts as one byte each (4-byte counts for a tile "
                                         "in which some row has more than 255 hits)",
                                  "value": e_rows_total / (eu_ms_max * 1e-3), "unit": "probe intervals/s", "ms_per_step": eu_ms_max,
                                  "d2h_bytes_per_step": 4 * e_pairs + e2e_rows + 16 * n_tiles,
                                  "d2h_bytes_moved_per_step": eu_stats["d2h_bytes"] // max(args.e2e_steps, 1),
                                  "pairs_digest_equals_device": True},
                    "count_only": {"value": e_rows_total / (ec_ms_max * 1e-3), "unit": "probe intervals/s", "ms_per_step": ec_ms_max,
                                   "api": "same pipeline with SQ_TILE_COUNT_ONLY: count(1) of the join, host columns in, one number per tile out",
                                   "h2d_bytes_per_step": 16 * e2e_rows, "d2h_bytes_per_step": 16 * n_tiles,
                                   "link_GBps": {"h2d": 16 * e2e_rows / (ec_ms * 1e-3) / 1e9}}},
            "gpu_launches": int(launches), "clocks": clocks,
            "digest": {"pairs": dcount, "sum": dsum},
            "oracle_check": check, "ranks": per_rank,
        }
        if world > 1 or args.parallelism == "contig":
            pr = [r["probe_rows"] for r in per_rank]
            result["balance"] = {"probe_rows_max_over_mean": max(pr) / (sum(pr) / len(pr)),
                                 "build_rows_per_rank": [r["build_rows"] for r in per_rank],
                                 "build_ms_per_rank": [r["build_ms"] for r in per_rank],
                                 "index_bytes_per_rank": [r["index_bytes"] for r in per_rank]}
        for k, v in (("locality", locality), ("materialise", materialise), ("count_only", count_only), ("full_config", full),
                     ("sub_configs", sub), ("exec_node", exec_line)):
            if v:
                result[k] = v
        if not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)
            if args.workload == "cfg5_shard":
                bh, ph, what = cpu_sample_cfg5(args)
            else:
                b_, p_ = to_host(build), to_host(probe)
                bh, ph, what = b_, p_, "whole build side"
            result["cpu_baseline"] = cpu_baseline(bh, ph, what, threads=os.cpu_count() or 1, min_seconds=1.5)
            result["cpu_baseline"].pop("_index", None)
        print(json.dumps(result))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
