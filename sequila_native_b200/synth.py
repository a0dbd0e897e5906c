"""Seeded synthetic interval tables for BASELINE.json's configs 2-5.

Every generator returns ``(build, probe)``; each side is a dict of numpy arrays
``contig`` (int32 dictionary id), ``key`` (uint64, stands in for DataFusion's
``create_hashes(on)`` — any injective map of the key reproduces the reference's grouping,
interval_join.rs:1042-1048), ``start``/``end`` (int32, closed interval) — i.e. the columns
`contig, pos_start, pos_end` of the reference's tables (queries/q1-coitrees.sql:6-14).

``scale`` shrinks row counts *and* contig lengths together so that the hit fan-out per probe
stays that of the full-size config (parity tests run the same shapes at oracle-friendly sizes).
"""
from __future__ import annotations

import numpy as np

# hg38 primary contig lengths chr1..22, X, Y (external constants)
HG38 = np.array([248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973,
                 145138636, 138394717, 133797422, 135086622, 133275309, 114364328, 107043718,
                 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468,
                 156040895, 57227415], dtype=np.int64)
CONTIG_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def key_hash(contig_id) -> np.ndarray:
    """Injective 64-bit stand-in for create_hashes(on=[contig]) (splitmix64 finalizer)."""
    with np.errstate(over="ignore"):
        x = np.asarray(contig_id).astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(30)
        x *= _M1
        x ^= x >> np.uint64(27)
        x *= _M2
        x ^= x >> np.uint64(31)
    return x


def _table(contig, start, end):
    contig = contig.astype(np.int32)
    return {"contig": contig, "key": key_hash(contig), "start": start.astype(np.int32), "end": end.astype(np.int32)}


def _uniform_side(rng, n, lengths, weights, wlo, whi):
    contig = rng.choice(len(lengths), size=n, p=weights) if weights is not None else rng.integers(0, len(lengths), n)
    L = lengths[contig]
    w = rng.integers(wlo, whi + 1, n)
    w = np.minimum(w, L)
    start = (rng.random(n) * (L - w + 1)).astype(np.int64)
    return _table(contig, start, start + w - 1)


def _scaled_lengths(lengths, scale):
    return np.maximum((lengths * scale).astype(np.int64), 1000)


def cfg2(scale: float = 1.0, nb: int = 1_000_000, np_: int = 1_000_000):
    """1M x 1M, 24 contigs x 10 Mbp, uniform starts, widths U{50..150}; seeds 1001/1002."""
    L = _scaled_lengths(np.full(24, 10_000_000, dtype=np.int64), scale)
    nb, np_ = max(int(nb * scale), 1), max(int(np_ * scale), 1)
    build = _uniform_side(np.random.default_rng(1001), nb, L, None, 50, 150)
    probe = _uniform_side(np.random.default_rng(1002), np_, L, None, 50, 150)
    return build, probe


def cfg4(scale: float = 1.0, nb: int = 1_000_000, np_: int = 1_000_000):
    """High fan-out: build widths U{100k..500k}, probe widths U{50..150}, hg38-weighted contigs."""
    L = _scaled_lengths(HG38, scale)
    w = HG38 / HG38.sum()
    nb, np_ = max(int(nb * scale), 1), max(int(np_ * scale), 1)
    build = _uniform_side(np.random.default_rng(20261018), nb, L, w, 100_000, 500_000)
    probe = _uniform_side(np.random.default_rng(20261019), np_, L, w, 50, 150)
    return build, probe


def cfg5(scale: float = 1.0, nb: int = 100_000_000, np_: int = 100_000_000):
    """100M x 100M, hg38-weighted contigs, widths U{50..150} both sides; seeds 5001/5002."""
    L = _scaled_lengths(HG38, scale)
    w = HG38 / HG38.sum()
    nb, np_ = max(int(nb * scale), 1), max(int(np_ * scale), 1)
    build = _uniform_side(np.random.default_rng(5001), nb, L, w, 50, 150)
    probe = _uniform_side(np.random.default_rng(5002), np_, L, w, 50, 150)
    return build, probe


# ---- counter-based generator: row i of a side depends on (seed, i) only --------------
# so any slice of a side can be produced anywhere — on the device with torch for the CUDA arm of bench.py, on
# the host with numpy for the reference arm and the tests — and the two are bit-identical.  Integer arithmetic
# only: three splitmix64 outputs per row give the contig (32-bit value against cumulative hg38 thresholds), the
# width and the start (32 x 32 -> 64-bit multiply-shift).
_GAMMA = 0x9E3779B97F4A7C15
_MASK64 = (1 << 64) - 1


def _contig_thresholds(lengths):
    cum = np.cumsum(lengths.astype(np.float64)) / float(lengths.sum())
    t = np.minimum(np.floor(cum * 4294967296.0), 4294967295.0).astype(np.int64)
    t[-1] = 1 << 32  # every 32-bit value falls below the last threshold
    return t


def _mix64_np(x):
    with np.errstate(over="ignore"):
        x = x ^ (x >> np.uint64(30))
        x = x * _M1
        x = x ^ (x >> np.uint64(27))
        x = x * _M2
        x = x ^ (x >> np.uint64(31))
    return x


def counter_side(n: int, seed: int, first_row: int = 0, wlo: int = 50, whi: int = 150, lengths=HG38, chunk: int = 1 << 24,
                 keep_contigs=None):
    """Rows [first_row, first_row + n) of the side `seed` (numpy, host).  keep_contigs: only rows of those
    contigs are kept (the bounded CPU-baseline sample of bench.py), generated chunk by chunk."""
    thr = _contig_thresholds(lengths)
    out = []
    keep = None if keep_contigs is None else np.asarray(sorted(int(c) for c in keep_contigs), dtype=np.int64)
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        i = np.arange(first_row + lo, first_row + lo + m, dtype=np.uint64)
        with np.errstate(over="ignore"):
            base = np.uint64((seed * _GAMMA) & _MASK64) + (i * np.uint64(3) + np.uint64(1)) * np.uint64(_GAMMA)
            u1 = _mix64_np(base) >> np.uint64(32)
            u2 = _mix64_np(base + np.uint64(_GAMMA)) >> np.uint64(32)
            u3 = _mix64_np(base + np.uint64((2 * _GAMMA) & _MASK64)) >> np.uint64(32)
        contig = np.searchsorted(thr, u1.astype(np.int64), side="right").astype(np.int64)
        L = lengths[contig]
        w = np.minimum(wlo + (u2.astype(np.int64) % (whi - wlo + 1)), L)
        start = (u3.astype(np.int64) * (L - w + 1)) >> 32
        if keep is not None:
            sel = np.isin(contig, keep)
            contig, start, w = contig[sel], start[sel], w[sel]
        out.append((contig.astype(np.int32), start.astype(np.int32), (start + w - 1).astype(np.int32)))
    contig = np.concatenate([o[0] for o in out]) if out else np.zeros(0, np.int32)
    start = np.concatenate([o[1] for o in out]) if out else np.zeros(0, np.int32)
    end = np.concatenate([o[2] for o in out]) if out else np.zeros(0, np.int32)
    return {"contig": contig, "key": key_hash(contig), "start": start, "end": end}


def counter_side_torch(n: int, seed: int, device, first_row: int = 0, wlo: int = 50, whi: int = 150, lengths=HG38,
                       chunk: int = 1 << 25):
    """The same rows as counter_side, generated on `device` (torch int64 arithmetic wraps like uint64; logical
    right shifts are arithmetic shifts with the sign bits masked off).  'key' is int64 (the bits of the u64 hash)."""
    import torch

    def s64(v):
        v &= _MASK64
        return v - (1 << 64) if v >= (1 << 63) else v

    def mix(x):
        x = x ^ ((x >> 30) & ((1 << 34) - 1))
        x = x * s64(int(_M1))
        x = x ^ ((x >> 27) & ((1 << 37) - 1))
        x = x * s64(int(_M2))
        x = x ^ ((x >> 31) & ((1 << 33) - 1))
        return x

    thr = torch.from_numpy(_contig_thresholds(lengths)).to(device)
    Lt = torch.from_numpy(np.asarray(lengths, dtype=np.int64)).to(device)
    keys = torch.from_numpy(key_hash(np.arange(len(lengths))).view(np.int64)).to(device)
    contig_o = torch.empty(n, dtype=torch.int32, device=device)
    key_o = torch.empty(n, dtype=torch.int64, device=device)
    start_o = torch.empty(n, dtype=torch.int32, device=device)
    end_o = torch.empty(n, dtype=torch.int32, device=device)
    g = s64(_GAMMA)
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        i = torch.arange(first_row + lo, first_row + lo + m, dtype=torch.int64, device=device)
        base = (i * 3 + 1) * g + s64(seed * _GAMMA)
        u1 = (mix(base) >> 32) & 0xFFFFFFFF
        u2 = (mix(base + g) >> 32) & 0xFFFFFFFF
        u3 = (mix(base + s64(2 * _GAMMA)) >> 32) & 0xFFFFFFFF
        contig = torch.searchsorted(thr, u1, right=True)
        L = Lt[contig]
        w = torch.minimum(wlo + (u2 % (whi - wlo + 1)), L)
        start = (u3 * (L - w + 1)) >> 32
        contig_o[lo:lo + m] = contig.int()
        key_o[lo:lo + m] = keys[contig]
        start_o[lo:lo + m] = start.int()
        end_o[lo:lo + m] = (start + w - 1).int()
    return {"contig": contig_o, "key": key_o, "start": start_o, "end": end_o}


def cfg5_probe_shard(rank: int, world: int, rows_per_rank: int = 12_500_000, seed: int = 5002):
    """Probe shard of config 5 for one GPU (weak scaling: rows_per_rank fixed, world grows)."""
    L = HG38
    w = HG38 / HG38.sum()
    rng = np.random.default_rng([seed, rank, world])
    return _uniform_side(rng, rows_per_rank, L, w, 50, 150)


def cfg3(scale: float = 1.0, nb: int = 1_200_000, np_: int = 10_000_000, seed: int = 7):
    """'databio-shaped' (ex-anno x ex-rna like): exon-copy model with skewed lengths.

    6000 loci x 10 exons; exon width ~ lognormal(ln 150, 0.8) clipped [20, 20000]; a build row is a
    random (locus, exon) copy, 30 % of them jittered by +-30 bp (=> ~20 near-duplicates per exon);
    a probe is a 76 bp read on an exon of a locus drawn with lognormal(0, 1) expression weights,
    15 % 'spliced' with an extra span ~ lognormal(ln 2000, 1.2) clipped to 500 kbp.
    """
    rng = np.random.default_rng(seed)
    G, E = max(int(6000 * scale), 8), 10
    nb, np_ = max(int(nb * scale), 1), max(int(np_ * scale), 1)
    L = HG38
    locus_contig = rng.choice(24, size=G, p=HG38 / HG38.sum())
    locus_pos = (rng.random(G) * (L[locus_contig] - 2_000_000)).astype(np.int64) + 500_000
    exon_w = np.clip(np.exp(rng.normal(np.log(150.0), 0.8, (G, E))), 20, 20000).astype(np.int64)
    intron = np.clip(np.exp(rng.normal(np.log(3000.0), 1.0, (G, E))), 100, 200_000).astype(np.int64)
    exon_off = np.cumsum(exon_w + intron, axis=1) - (exon_w + intron)
    exon_start = locus_pos[:, None] + exon_off

    g = rng.integers(0, G, nb)
    e = rng.integers(0, E, nb)
    jit = np.where(rng.random(nb) < 0.3, rng.integers(-30, 31, nb), 0)
    bs = exon_start[g, e] + jit
    be = bs + exon_w[g, e] - 1 + np.where(rng.random(nb) < 0.3, rng.integers(-30, 31, nb), 0)
    be = np.maximum(be, bs)
    build = _table(locus_contig[g], bs, be)

    expr = np.exp(rng.normal(0.0, 1.0, G))
    g = rng.choice(G, size=np_, p=expr / expr.sum())
    e = rng.integers(0, E, np_)
    ps = exon_start[g, e] + (rng.random(np_) * np.maximum(exon_w[g, e] - 1, 1)).astype(np.int64) - 38
    extra = np.where(rng.random(np_) < 0.15,
                     np.clip(np.exp(rng.normal(np.log(2000.0), 1.2, np_)), 0, 500_000).astype(np.int64), 0)
    pe = ps + 75 + extra
    probe = _table(locus_contig[g], np.maximum(ps, 0), np.maximum(pe, 0))
    return build, probe


def fixtures_reads_targets():
    """BASELINE config 1: the reference's only shipped fixtures, testing/data/interval/reads.csv
    (build, 12 rows) x targets.csv (probe, 10 rows), inlined so that GPU-box tests do not need
    /root/reference.  contig ids: chr1 -> 0, chr2 -> 1."""
    reads = [(0, 150, 250), (0, 190, 300), (0, 300, 501), (0, 500, 700), (0, 22000, 22300), (0, 15000, 15000),
             (1, 150, 250), (1, 190, 300), (1, 300, 500), (1, 500, 700), (1, 22000, 22300), (1, 15000, 15000)]
    targets = [(0, 100, 190), (0, 200, 290), (0, 400, 600), (0, 10000, 20000), (0, 22100, 22100),
               (1, 100, 190), (1, 200, 290), (1, 400, 600), (1, 10000, 20000), (1, 22100, 22100)]
    r = np.array(reads, dtype=np.int64)
    t = np.array(targets, dtype=np.int64)
    return _table(r[:, 0], r[:, 1], r[:, 2]), _table(t[:, 0], t[:, 1], t[:, 2])


CONFIGS = {"cfg2": cfg2, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5}
