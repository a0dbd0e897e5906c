"""GPU: the two sort-key widths of the build (include/sequila_cuda.h, sq_index_sort_key_bits).  32-bit keys — the keys'
start ranges laid end to end — are picked when they fit, 64-bit (key id : start) otherwise or with
`SET sequila.cuda_build_sort TO wide`; the index, and so every pair in its order of emission, must be the same, and equal
to the oracle's (build = interval_join.rs:662-683, coitrees sort key CT/nosimd.rs:869-947)."""
import numpy as np
import pytest

import sequila_native_b200 as sn

pytestmark = pytest.mark.gpu

EMPTY_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)  # the hash table's own empty marker: a legal key hash all the same


def join(ctx, b, p, sort):
    ctx.set_option("sequila.cuda_build_sort", sort)
    try:
        idx = sn.CudaIndex.build(ctx, b["key"], b["start"], b["end"])
    finally:
        ctx.set_option("sequila.cuda_build_sort", "auto")
    st = sn.CudaStream(ctx)
    n = st.probe_count(idx, p["key"], p["start"], p["end"])
    l, r, c = st.emit_pairs()
    assert n == len(l)
    near = sn.CudaStream(ctx).probe_nearest(idx, p["key"], p["start"], p["end"])
    return idx, l, r, c, near


def check(ctx, oracle, b, p, want_bits):
    b = {k: np.ascontiguousarray(v) for k, v in b.items()}
    p = {k: np.ascontiguousarray(v) for k, v in p.items()}
    ia, la, ra, ca, na = join(ctx, b, p, "auto")
    iw, lw, rw, cw, nw = join(ctx, b, p, "wide")
    assert iw.sort_key_bits == 64
    assert ia.sort_key_bits == want_bits
    assert ia.keys == iw.keys and ia.uses_packed == iw.uses_packed and ia.uses_rank == iw.uses_rank
    assert np.array_equal(la, lw) and np.array_equal(ra, rw) and np.array_equal(ca, cw) and np.array_equal(na, nw)
    ol, orr, oc = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    assert np.array_equal(ca, oc)
    assert np.array_equal(oracle.sorted_pairs(la, ra), oracle.sorted_pairs(ol, orr))


# the domain on which the reference is defined: its AVX coitrees nodes store first - 1 and last + 1 (CT/avx.rs:410-427),
# which wrap at the int32 extremes (the oracle restates that; tests/test_gpu_adversarial.py keeps the same bounds)
LO, HI = -2**31 + 1, 2**31 - 2


def side(rng, n, n_keys, lo, hi, wmax, key_base=1):
    key = rng.integers(key_base, key_base + n_keys, n, dtype=np.uint64)
    start = rng.integers(lo, hi, n, dtype=np.int64)
    end = np.minimum(start + rng.integers(0, wmax, n, dtype=np.int64), np.int64(HI))
    return {"key": key, "start": start.astype(np.int32), "end": end.astype(np.int32)}


@pytest.mark.parametrize("layout", ["auto", "packed", "soa"])
@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
def test_benchmark_shapes(cuda_ctx, oracle, name, layout):
    b, p = {"cfg2": lambda: sn.synth.cfg2(scale=0.05), "cfg3": lambda: sn.synth.cfg3(scale=0.004),
            "cfg4": lambda: sn.synth.cfg4(scale=0.02)}[name]()
    cuda_ctx.set_option("sequila.cuda_probe_layout", layout)
    try:
        check(cuda_ctx, oracle, b, p, 32)
    finally:
        cuda_ctx.set_option("sequila.cuda_probe_layout", "auto")


def test_negative_and_extreme_starts_in_one_key(cuda_ctx, oracle):
    rng = np.random.default_rng(11)
    b = side(rng, 30_000, 1, LO, HI, 5_000_000)
    b["start"][:4] = [LO, HI, LO, HI]  # the span of the one key is 2^32 - 2: all 32 key bits in use
    b["end"][:4] = [LO, HI, HI, HI]
    p = side(rng, 20_000, 1, LO, HI, 80_000_000)
    p["start"][:2] = [LO, HI]
    p["end"][:2] = [LO, HI]
    check(cuda_ctx, oracle, b, p, 32)


def test_two_keys_spanning_everything_fall_back_to_wide_keys(cuda_ctx, oracle):
    rng = np.random.default_rng(12)
    b = side(rng, 20_000, 2, LO, HI, 3_000_000)
    b["start"][:4] = [LO, HI, LO, HI]
    b["end"][:4] = [0, HI, 5, HI]
    b["key"][:4] = [1, 1, 2, 2]
    p = side(rng, 20_000, 3, LO, HI, 50_000_000)
    check(cuda_ctx, oracle, b, p, 64)


@pytest.mark.parametrize("n_keys,bits", [(4096, 32), (4097, 64), (20_000, 64)])
def test_key_count_limit(cuda_ctx, oracle, n_keys, bits):
    rng = np.random.default_rng(n_keys)
    b = side(rng, 60_000, n_keys, 0, 100_000, 300)
    b["key"][:n_keys] = np.arange(1, n_keys + 1, dtype=np.uint64)  # every key present
    p = side(rng, 40_000, n_keys + 50, 0, 100_000, 300)
    check(cuda_ctx, oracle, b, p, bits)


def test_sentinel_key_single_rows_and_equal_starts(cuda_ctx, oracle):
    rng = np.random.default_rng(14)
    b = side(rng, 5_000, 6, 1000, 1200, 40)
    b["key"][::7] = EMPTY_KEY
    b["start"][100:400] = 1100  # a long run of equal starts in whatever keys they fall
    b["end"][100:400] = 1100 + (np.arange(300) % 9)
    p = side(rng, 5_000, 8, 900, 1300, 60)
    p["key"][::5] = EMPTY_KEY
    check(cuda_ctx, oracle, b, p, 32)
    one = {k: v[:1] for k, v in b.items()}
    check(cuda_ctx, oracle, one, p, 32)
    same_key = {k: v.copy() for k, v in b.items()}
    same_key["key"][:] = 7
    check(cuda_ctx, oracle, same_key, p, 32)
    only_sentinel = {k: v.copy() for k, v in b.items()}
    only_sentinel["key"][:] = EMPTY_KEY
    check(cuda_ctx, oracle, only_sentinel, p, 32)


def test_inverted_rows_and_position_ids(cuda_ctx, oracle):
    """end < start rows (never hit by the closed predicate unless the probe is inverted too) go through both sorts alike;
    position ids ride on either"""
    rng = np.random.default_rng(15)
    b = side(rng, 20_000, 5, -50_000, 50_000, 2_000)
    b["end"][::9] = b["start"][::9] - 3
    p = side(rng, 20_000, 5, -50_000, 50_000, 2_000)
    check(cuda_ctx, oracle, b, p, 32)
    cuda_ctx.set_option("sequila.cuda_build_ids", "positions")
    try:
        ia, la, ra, ca, _ = join(cuda_ctx, b, p, "auto")
        iw, lw, rw, cw, _ = join(cuda_ctx, b, p, "wide")
    finally:
        cuda_ctx.set_option("sequila.cuda_build_ids", "rows")
    assert ia.uses_positions and iw.uses_positions and ia.sort_key_bits == 32 and iw.sort_key_bits == 64
    # positions belong to one index (key ids, and so the order of the key segments, are handed out in hash-table order
    # by whichever thread gets there first); the build rows they name must agree
    assert np.array_equal(ia.position_rows()[la], iw.position_rows()[lw]) and np.array_equal(ra, rw)
    ol, orr, oc = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    assert np.array_equal(oracle.sorted_pairs(ia.position_rows()[la], ra), oracle.sorted_pairs(ol, orr))
