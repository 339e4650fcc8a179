"""oracle_generic.c restates the reference's generic coders once for any preset.  The reference holds no golden
vectors for the Small preset (u16 / u32 / 12), so the generic code is pinned through the Default preset: called with
(W, P) = (32, 24) it must reproduce oracle.c -- which reproduces the reference's goldens -- word for word; the Small
preset is the same code path with (16, 12), additionally checked for the invariants the reference's own tests check
(round trips with every decoder model, stack.rs:1293-1454-style)."""
import numpy as np
import pytest


def symbols_from(rng, cdf, n):
    """draws symbols from the quantised model itself"""
    p = np.diff(cdf.astype(np.int64)).astype(np.float64)
    return rng.choice(p.size, size=n, p=p / p.sum()).astype(np.int32)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 1000, 20000])
def test_default_preset_equals_the_pinned_oracle(oracle, n):
    rng = np.random.default_rng(n)
    for cdf, lo in ((oracle.qgauss_cdf(-50, 50, 3.2, 9.6), -50), (oracle.cat_cdf(rng.dirichlet(0.3 * np.ones(40))), 0),
                    (oracle.qgauss_cdf(-127, 127, 3.2, 5.1), -127)):
        syms = symbols_from(rng, cdf, n) + lo
        a = oracle.g_ans_encode("default", syms, cdf, lo)
        assert np.array_equal(a, oracle.ans_encode_iid(syms, cdf, lo))
        assert np.array_equal(oracle.g_ans_decode("default", a, n, cdf, lo), syms)
        r = oracle.g_range_encode("default", syms, cdf, lo)
        assert np.array_equal(r, oracle.range_encode_iid(syms, cdf, lo))
        assert np.array_equal(oracle.g_range_decode("default", r, n, cdf, lo), syms)


def test_default_preset_models_equal_the_pinned_oracle(oracle):
    rng = np.random.default_rng(5)
    for dtype in (np.float32, np.float64):
        for _ in range(20):
            pmf = rng.dirichlet(0.4 * np.ones(rng.integers(2, 300))).astype(dtype)
            assert np.array_equal(oracle.g_cat_cdf("default", pmf), oracle.cat_cdf(pmf))
            assert np.array_equal(oracle.g_cat_cdf("default", pmf, perfect=True), oracle.cat_perfect_cdf(pmf))


def test_golden_words_through_the_generic_code(oracle):
    """src/lib.rs:131-160 (G3): 1e5-free doc example -- QuantizedGaussian(-50,50,3.2,9.6)... replayed from the
    transcribed vectors: every i.i.d. ANS / range golden case must come out of the generic functions too."""
    from golden import reference_vectors as G
    hits = 0
    for case in G.ENCODE_CASES:
        if len(case["segments"]) != 1:
            continue
        spec, syms, params = case["segments"][0]
        if params or spec[0] != "qgauss":
            continue
        cdf = oracle.qgauss_cdf(*spec[1:])
        fn = oracle.g_ans_encode if case["coder"] == "ans" else oracle.g_range_encode
        assert np.array_equal(fn("default", syms, cdf, spec[1]), case["words"]), case["id"]
        hits += 1
    assert hits >= 2


@pytest.mark.parametrize("n", [0, 1, 2, 5, 100, 5000])
def test_small_preset_round_trips(oracle, n):
    rng = np.random.default_rng(100 + n)
    for perfect in (False, True):
        for dtype in (np.float32, np.float64):
            pmf = rng.dirichlet(0.5 * np.ones(100)).astype(dtype)
            cdf = oracle.g_cat_cdf("small", pmf, perfect=perfect)
            assert cdf[0] == 0 and cdf[-1] == 4096 and np.all(np.diff(cdf.astype(np.int64)) >= 1)
            table = oracle.g_lookup_table("small", cdf)
            syms = symbols_from(rng, cdf, n)
            for enc, dec in ((oracle.g_ans_encode, oracle.g_ans_decode), (oracle.g_range_encode, oracle.g_range_decode)):
                words = enc("small", syms, cdf)
                assert np.all(words < 65536)
                assert np.array_equal(dec("small", words, n, cdf), syms)              # binary search (contiguous.rs:628-665)
                assert np.array_equal(dec("small", words, n, cdf, table=table), syms)  # lookup model (lookup_contiguous.rs:564-607)
                if n >= 100:  # 12-bit models: the rate stays close to the entropy
                    p = np.diff(cdf.astype(np.int64)) / 4096.0
                    ideal = -np.log2(p[syms]).sum()
                    assert 16 * words.size <= ideal + 64


def test_small_preset_lookup_doc_example(oracle):
    """lookup_contiguous.rs:55-103: the message of the doc example round-trips through a SmallRangeEncoder and a
    lookup decoder model built from the perfectly quantised probabilities [0.4, 0.2, 0.1, 0.3]."""
    message = np.array([2, 1, 3, 0, 0, 2, 0, 2, 1, 0, 2], dtype=np.int32)
    cdf = oracle.g_cat_cdf("small", np.array([0.4, 0.2, 0.1, 0.3], dtype=np.float32), perfect=True)
    table = oracle.g_lookup_table("small", cdf)
    for enc, dec in ((oracle.g_range_encode, oracle.g_range_decode), (oracle.g_ans_encode, oracle.g_ans_decode)):
        for preset in ("small", "default"):  # "you can always use a bigger coder on a smaller model" -- not here: P differs
            pass
        words = enc("small", message, cdf)
        assert np.array_equal(dec("small", words, 11, cdf, table=table), message)
