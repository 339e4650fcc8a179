"""The product's coder / model arithmetic (constriction_b200/csrc/*_math.cuh, compiled for the host)
against the oracle and against native u64 division.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

import host_harness

u32p, i32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)


def P(a, t):
    return a.ctypes.data_as(t)


@pytest.fixture(scope="module")
def H():
    return host_harness.load()


def test_division_by_reciprocal_is_exact(H):
    rng = np.random.default_rng(0)
    n = 2_000_000
    # divisors: all magnitudes up to 2^24 incl. 1, powers of two and 2^24-1
    d = (rng.integers(1, 1 << 24, size=n, dtype=np.uint64) >> rng.integers(0, 24, size=n, dtype=np.uint64)).astype(np.uint32)
    d = np.maximum(d, 1)
    d[:64] = [1, 2, 3, (1 << 24) - 1, 1 << 23, (1 << 23) + 1, 1 << 12, 255] * 8
    # dividends: the encoder's range n < d * 2^40, plus arbitrary u64 and extremes
    nn = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * 2 + rng.integers(0, 2, size=n, dtype=np.uint64)
    lim = d.astype(np.uint64) << np.uint64(40)
    nn[: n // 2] = nn[: n // 2] % lim[: n // 2]
    nn[n // 2: n // 2 + 64] = np.uint64(2**64 - 1)
    nn[n // 2 + 64: n // 2 + 128] = lim[n // 2 + 64: n // 2 + 128] - np.uint64(1)
    nn = np.ascontiguousarray(nn)
    assert H.h_divmod_check(P(nn, u64p), P(d, u32p), n) == 0
    # the FP64-pipe estimate is only claimed for the encoder's operand range n < d * 2^40
    m = np.ascontiguousarray(nn % lim)
    m[:64] = lim[:64] - np.uint64(1)
    m[64:128] = (lim[64:128] // np.uint64(2)) | np.uint64(1)
    assert H.h_f64div_check(P(m, u64p), P(d, u32p), n) == 0


def test_range_quantile_division_is_exact(H):
    rng = np.random.default_rng(1)
    n = 2_000_000
    rg = rng.integers(1 << 32, 1 << 63, size=n, dtype=np.uint64) * 2 + 1
    rg[: n // 4] = rng.integers(1 << 32, 1 << 40, size=n // 4, dtype=np.uint64)
    rg[0] = np.uint64(2**64 - 1)
    rg[1] = np.uint64(1 << 32)
    frac = rng.random(n)
    diff = (rg.astype(np.float64) * frac).astype(np.uint64)
    diff = np.minimum(diff, rg - 1)
    # land exactly on and next to multiples of scale (worst case for the +-1 correction)
    scale = rg >> np.uint64(24)
    q = rng.integers(0, 1 << 24, size=n, dtype=np.uint64)
    diff[n // 2:] = (q * scale)[n // 2:] + rng.integers(0, 3, size=n - n // 2, dtype=np.uint64) - 1
    diff[n // 2] = 0
    # some invalid ones (quantile >= 2^24)
    diff[1::1000] = (scale[1::1000] << np.uint64(24)) + rng.integers(0, 1000, size=diff[1::1000].size, dtype=np.uint64)
    diff = np.ascontiguousarray(diff)
    rg = np.ascontiguousarray(rg)
    assert H.h_range_quantile_check(P(diff, u64p), P(rg, u64p), n) == 0


MODELS = [(-50, 50, 3.2, 9.6), (-100, 100, 12.6, 7.3), (-127, 127, 3.2, 5.1), (0, 1, 0.4, 0.2), (-2000, 2000, 10.0, 600.0),
          (-20, 20, 0.0, 1e-3), (-64, 64, -1.7, 0.3)]


@pytest.mark.parametrize("spec", MODELS)
def test_gaussian_table_matches_oracle(H, oracle, spec):
    lo, hi, mean, std = spec
    want = oracle.qgauss_cdf(lo, hi, mean, std)
    got = np.empty_like(want)
    assert H.h_qgauss_cdf(lo, hi, mean, std, P(got, u32p)) == 0
    assert np.array_equal(got, want)


def test_erf_exp_match_oracle(H, oracle):
    rng = np.random.default_rng(2)
    xs = np.concatenate([rng.normal(0, 3, 20000), rng.uniform(-7, 7, 20000), [0.0, -0.0, 1e-300, 0.84375, 1.25, 2.857, 6.0, 28.0]])
    for x in xs:
        assert H.h_erf(float(x)) == oracle.erf(float(x))
    for x in np.concatenate([rng.uniform(-750, 710, 20000), rng.normal(0, 1, 20000)]):
        assert H.h_exp(float(x)) == oracle.exp(float(x))


def _sample(rng, cdf, lo, n):
    pmf = np.diff(cdf.astype(np.int64)) / float(1 << 24)
    return (rng.choice(len(pmf), size=n, p=pmf) + lo).astype(np.int32)


@pytest.mark.parametrize("spec", MODELS[:4] + [(-20, 20, 0.0, 1e-3)])
@pytest.mark.parametrize("n", [0, 1, 2, 17, 1000, 20000])
def test_ans_stream_matches_oracle(H, oracle, spec, n):
    lo, hi, mean, std = spec
    cdf = oracle.qgauss_cdf(lo, hi, mean, std)
    rng = np.random.default_rng(n + 7)
    syms = _sample(rng, cdf, lo, n)
    if n >= 17:  # also hit the tails (probability 1..few), which the sampler never reaches
        syms[::5] = rng.integers(lo, hi + 1, size=syms[::5].size)
    want = oracle.ans_encode_iid(syms, cdf, lo)
    out = np.empty(n + 2, dtype=np.uint32)
    st = C.c_uint64()
    for f64 in (0, 1):
        m = H.h_ans_encode(P(syms, i32p), n, P(cdf, u32p), lo, 0, P(out, u32p), C.byref(st), f64)
        assert np.array_equal(out[:m], want), f64
    dec = np.empty(n, dtype=np.int32)
    H.h_ans_decode(P(want, u32p), want.size, P(dec, i32p), n, P(cdf, u32p), cdf.size - 1, lo)
    assert np.array_equal(dec, syms)


@pytest.mark.parametrize("spec", MODELS[:4] + [(-20, 20, 0.0, 1e-3)])
@pytest.mark.parametrize("n", [0, 1, 2, 17, 1000, 20000])
def test_range_stream_matches_oracle(H, oracle, spec, n):
    lo, hi, mean, std = spec
    cdf = oracle.qgauss_cdf(lo, hi, mean, std)
    rng = np.random.default_rng(n + 11)
    syms = _sample(rng, cdf, lo, n)
    if n >= 17:
        syms[::5] = rng.integers(lo, hi + 1, size=syms[::5].size)
    want = oracle.range_encode_iid(syms, cdf, lo)
    out = np.empty(n + 8, dtype=np.uint32)
    m = H.h_range_encode(P(syms, i32p), n, P(cdf, u32p), lo, P(out, u32p))
    assert np.array_equal(out[:m], want)
    dec = np.empty(n, dtype=np.int32)
    assert H.h_range_decode(P(want, u32p), want.size, P(dec, i32p), n, P(cdf, u32p), cdf.size - 1, lo) == 0
    assert np.array_equal(dec, syms)


@pytest.mark.parametrize("spec", MODELS[:2] + [(-20, 20, 0.0, 1e-3)])
def test_range_encoder_suspend_resume_at_every_symbol(H, oracle, spec):
    """The eager-word formulation recovers the reference's (num_inverted, first_inverted) from the words
    written so far; suspending and resuming at any symbol must not change the stream."""
    lo, hi, mean, std = spec
    cdf = oracle.qgauss_cdf(lo, hi, mean, std)
    rng = np.random.default_rng(77)
    n = 1500
    syms = _sample(rng, cdf, lo, n)
    syms[::7] = rng.integers(lo, hi + 1, size=syms[::7].size)
    want = oracle.range_encode_iid(syms, cdf, lo)
    out = np.empty(n + 8, dtype=np.uint32)
    for split in range(n + 1):
        m = H.h_range_encode_split(P(syms, i32p), n, split, P(cdf, u32p), lo, P(out, u32p))
        assert np.array_equal(out[:m], want), split


def test_oracle_erf_exp_within_one_ulp_of_libm(oracle):
    """The restated msun erf / exp (what Rust's `libm` crate computes) against this machine's C library on random
    and special arguments: at most 1 ulp apart (both are < 1 ulp accurate), exact at the special values."""
    import math
    rng = np.random.default_rng(99)
    xs = np.concatenate([rng.normal(0, 1.5, 40_000), rng.uniform(-6.5, 6.5, 40_000), rng.uniform(-1e-3, 1e-3, 5_000),
                         [0.0, -0.0, 0.84375, 1.25, 1 / 0.35, 6.0, -6.0, 27.0, -27.0, 1e-300, 5e-324]])
    worst = 0.0
    for x in xs:
        a, b = oracle.erf(float(x)), math.erf(float(x))
        if a != b:
            worst = max(worst, abs(a - b) / math.ulp(b))
    assert worst <= 1.0, worst
    es = np.concatenate([rng.uniform(-40, 3, 40_000), rng.uniform(-0.4, 0.4, 10_000), [0.0, -745.2, 709.0, -1e-20]])
    worst = 0.0
    for x in es:
        a, b = oracle.exp(float(x)), math.exp(float(x))
        if a != b:
            worst = max(worst, abs(a - b) / max(math.ulp(b), 5e-324))
    assert worst <= 1.0, worst
    assert oracle.erf(float("inf")) == 1.0 and oracle.erf(float("-inf")) == -1.0 and math.isnan(oracle.erf(float("nan")))
