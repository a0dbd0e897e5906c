"""CPU: the host-side expansion of right_idx from the per-row counts (the wire format of the host entry points;
the reference's own loop, interval_join.rs:1611-1618) — every SIMD variant of sq_rle.cpp, through the library's
exported hook, against numpy.repeat; nothing may be written behind the last pair."""
import ctypes as C

import numpy as np
import pytest


def expand(lib, variant, counts, pad=16):
    want = np.repeat(np.arange(len(counts), dtype=np.uint32), counts)
    out = np.full(len(want) + pad, 0xDEADBEEF, np.uint32)
    ok = lib.sq_rle_expand_variant(variant, C.c_void_p(counts.ctypes.data), len(counts), C.c_void_p(out.ctypes.data), len(want))
    return ok, out, want


@pytest.mark.parametrize("variant", [-1, 0, 1, 2, 3])
def test_expand_counts_variants(variant):
    from sequila_native_b200 import _native
    lib = _native.lib()
    rng = np.random.default_rng(7 + variant)
    ran = 0
    for trial in range(200):
        n = int(rng.integers(0, 600))
        kind = trial % 5
        if kind == 0:
            counts = rng.poisson(6.4, n)
        elif kind == 1:
            counts = rng.integers(0, 3, n)
        elif kind == 2:
            counts = rng.integers(0, 70, n)          # runs longer than one 16-value block
        elif kind == 3:
            counts = (rng.random(n) < 0.05) * rng.integers(0, 900, n)  # mostly empty rows, a few long runs
        else:
            counts = np.zeros(n)
        counts = counts.astype(np.uint32)
        ok, out, want = expand(lib, variant, counts)
        if not ok:
            assert variant in (2, 3)  # the CPU lacks AVX2 / AVX-512: nothing to check
            continue
        ran += 1
        assert np.array_equal(out[:len(want)], want), (trial, variant)
        assert (out[len(want):] == 0xDEADBEEF).all(), "wrote behind the last pair"
    assert ran or variant in (2, 3)


@pytest.mark.parametrize("variant", [-1, 0, 1, 2, 3])
def test_counts_that_sum_past_n_pairs_never_write_past_it(variant):
    """counts and n_pairs that disagree (a caller bug, or counts of another tile): the expansion stops at n_pairs"""
    from sequila_native_b200 import _native
    lib = _native.lib()
    rng = np.random.default_rng(70 + variant)
    for trial in range(100):
        n = int(rng.integers(1, 400))
        counts = rng.integers(0, 90 if trial % 2 else 12, n).astype(np.uint32)
        full = np.repeat(np.arange(n, dtype=np.uint32), counts)
        if len(full) < 2:
            continue
        cut = int(rng.integers(1, len(full)))  # n_pairs < sum(counts)
        out = np.full(cut + 64, 0xDEADBEEF, np.uint32)
        ok = lib.sq_rle_expand_variant(variant, C.c_void_p(counts.ctypes.data), n, C.c_void_p(out.ctypes.data), cut)
        if not ok:
            assert variant in (2, 3)
            continue
        assert np.array_equal(out[:cut], full[:cut]), (trial, variant)
        assert (out[cut:] == 0xDEADBEEF).all(), "wrote past n_pairs"
