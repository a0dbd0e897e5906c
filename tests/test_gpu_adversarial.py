"""GPU: adversarial shapes for both index layouts (packed 128-byte lines / SoA arrays) against the oracle:
line-boundary sizes, equal starts, >32 hits per row, the narrow-encoding limits (16-bit width and in-line
offset), negative and extreme coordinates, thousands of tiny key segments, plus a hypothesis-driven sweep."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import sequila_native_b200 as sn
from helpers import canon

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["packed", "packed2", "soa", "staged", "soa_walk"])
def probe_layout(request, cuda_ctx):
    # "staged" = packed lines served by the shared-memory (TMA) staged kernel whatever the probe order: unsorted tiles
    # then exercise its global-memory path, sorted ones the staged path
    # "packed2" = the packed-line kernel with two tiles per CTA (option cuda_probe_tiles)
    # "soa" = the rank-difference kernel where the build side allows it; "soa_walk" = count / scan / write always
    cuda_ctx.set_option("sequila.cuda_probe_layout", {"staged": "packed", "packed2": "packed", "soa_walk": "soa"}.get(request.param, request.param))
    cuda_ctx.set_option("sequila.cuda_rank_count", "off" if request.param == "soa_walk" else "force")
    cuda_ctx.set_option("sequila.cuda_staged_probe", "on" if request.param == "staged" else "off")
    cuda_ctx.set_option("sequila.cuda_probe_tiles", 2 if request.param == "packed2" else 1)
    yield request.param
    cuda_ctx.set_option("sequila.cuda_probe_tiles", 1)
    cuda_ctx.set_option("sequila.cuda_probe_layout", "auto")
    cuda_ctx.set_option("sequila.cuda_staged_probe", "auto")
    cuda_ctx.set_option("sequila.cuda_rank_count", "on")


def check(oracle, ctx, b, p):
    b = {k: np.ascontiguousarray(v) for k, v in b.items()}
    p = {k: np.ascontiguousarray(v) for k, v in p.items()}
    idx = sn.CudaIndex.build(ctx, b["key"], b["start"], b["end"])
    s = sn.CudaStream(ctx)
    n = s.probe_count(idx, p["key"], p["start"], p["end"])
    l, r, c = s.emit_pairs()
    ol, orr, oc = oracle.join(b["key"], b["start"], b["end"], p["key"], p["start"], p["end"])
    assert n == len(ol) and np.array_equal(c, oc)
    assert np.all(np.diff(r.astype(np.int64)) >= 0)
    assert np.array_equal(canon(l, r), canon(ol, orr))
    # the fused one-call path must agree with count -> emit
    out = (np.empty(max(n, 1), np.uint32), np.empty(max(n, 1), np.uint32), np.empty(len(p["key"]), np.uint32))
    assert s.probe_join(idx, p["key"], p["start"], p["end"], out) == n
    assert np.array_equal(canon(out[0][:n], out[1][:n]), canon(ol, orr)) and np.array_equal(out[2], oc)


def side(key, start, width):
    start = np.asarray(start, dtype=np.int64)
    return {"key": np.asarray(key, dtype=np.uint64), "start": start.astype(np.int32),
            "end": (start + np.asarray(width, dtype=np.int64)).astype(np.int32)}


@pytest.mark.parametrize("n", [1, 14, 15, 16, 29, 30, 31, 45, 46, 255, 256, 257])
def test_segment_sizes_around_line_boundaries(cuda_ctx, oracle, n):
    rng = np.random.default_rng(n)
    b = side(np.zeros(n), np.sort(rng.integers(0, 40 * n, n)), rng.integers(0, 30, n))
    q = side(np.zeros(400), rng.integers(-20, 40 * n + 20, 400), rng.integers(0, 60, 400))
    check(oracle, cuda_ctx, b, q)


def test_all_starts_equal_and_duplicates(cuda_ctx, oracle):
    rng = np.random.default_rng(1)
    b = side(np.zeros(500), np.full(500, 1000), rng.integers(0, 50, 500))        # one bin, shift >= 32
    q = side(np.zeros(300), rng.integers(900, 1100, 300), rng.integers(0, 50, 300))
    check(oracle, cuda_ctx, b, q)
    b = side(np.zeros(2000), np.repeat(np.arange(100) * 7, 20), np.tile(np.arange(20), 100))  # 20 copies per start
    check(oracle, cuda_ctx, b, q)


def test_rows_with_more_hits_than_the_stash_holds(cuda_ctx, oracle):
    rng = np.random.default_rng(2)
    b = side(np.zeros(6000), rng.integers(0, 3000, 6000), rng.integers(0, 200, 6000))       # ~300 hits per probe row
    q = side(np.zeros(700), rng.integers(0, 3000, 700), rng.integers(0, 100, 700))
    check(oracle, cuda_ctx, b, q)
    q = side(np.zeros(64), np.zeros(64), np.full(64, 5000))                                  # every row hits everything
    check(oracle, cuda_ctx, b, q)


@pytest.mark.parametrize("width", [65534, 65535, 65536, 70000])
def test_narrow_encoding_limit_on_the_width(cuda_ctx, oracle, width):
    rng = np.random.default_rng(width)
    w = rng.integers(0, 100, 3000)
    w[::97] = width
    b = side(np.zeros(3000), rng.integers(0, 200000, 3000), w)
    q = side(np.zeros(2000), rng.integers(0, 270000, 2000), rng.integers(0, 100, 2000))
    check(oracle, cuda_ctx, b, q)


@pytest.mark.parametrize("gap", [65535, 65536, 1 << 20])
def test_narrow_encoding_limit_on_the_in_line_offset(cuda_ctx, oracle, gap):
    start = np.arange(400, dtype=np.int64) * 3
    start[200:] += gap                                      # one line spans the gap
    b = side(np.zeros(400), start, np.full(400, 10))
    q = side(np.zeros(1000), np.random.default_rng(5).integers(0, gap + 2000, 1000), np.full(1000, 50))
    check(oracle, cuda_ctx, b, q)


def test_negative_and_extreme_coordinates(cuda_ctx, oracle):
    rng = np.random.default_rng(3)
    lo, hi = -2 ** 31 + 1, 2 ** 31 - 2
    s = rng.integers(lo, hi - 1000, 4000)
    b = side(rng.integers(0, 3, 4000), s, rng.integers(0, 1000, 4000))
    qs = np.concatenate([s[:1500] + rng.integers(-500, 500, 1500), rng.integers(lo, hi - 1000, 1500)]).clip(lo, hi - 1000)
    q = side(rng.integers(0, 4, 3000), qs, rng.integers(0, 1000, 3000))
    check(oracle, cuda_ctx, b, q)
    # probes below / above everything, inverted probes, a probe spanning the whole domain
    q2 = {"key": np.zeros(4, np.uint64), "start": np.array([lo, hi - 5, 100, lo], np.int32), "end": np.array([lo + 3, hi, 50, hi], np.int32)}
    check(oracle, cuda_ctx, b, q2)


def test_thousands_of_tiny_key_segments(cuda_ctx, oracle):
    rng = np.random.default_rng(4)
    nk = 5000
    b = side(rng.integers(0, nk, 20000) * 2654435761 + 17, rng.integers(0, 500, 20000), rng.integers(0, 40, 20000))
    q = side(rng.integers(0, nk + 50, 15000) * 2654435761 + 17, rng.integers(0, 500, 15000), rng.integers(0, 40, 15000))
    check(oracle, cuda_ctx, b, q)


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
@given(nb=st.integers(0, 600), nq=st.integers(0, 400), nk=st.integers(1, 5), span=st.sampled_from([1, 7, 300, 70000, 5_000_000]),
       wmax=st.sampled_from([1, 3, 40, 900, 66000]), inverted=st.booleans(), seed=st.integers(0, 2 ** 31))
def test_hypothesis_sweep(cuda_ctx, oracle, nb, nq, nk, span, wmax, inverted, seed):
    rng = np.random.default_rng(seed)
    b = side(rng.integers(0, nk, nb), rng.integers(-span, span + 1, nb), rng.integers(0, wmax, nb))
    q = side(rng.integers(0, nk + 1, nq), rng.integers(-span, span + 1, nq), rng.integers(0, wmax, nq))
    if inverted and nb:
        m = rng.random(nb) < 0.2
        b["end"][m] = b["start"][m] - rng.integers(1, 9, int(m.sum())).astype(np.int32)
    check(oracle, cuda_ctx, b, q)
