"""The remaining model classes of the reference's Python API on the CUDA path: the default
(perfect=True) Categorical / Bernoulli, QuantizedLaplace, QuantizedCauchy, Binomial, CustomModel, ScipyModel.
Every table equals the oracle's, and streams coded through the mirror equal streams coded through the oracle's
restatement of the Python API (which passes the reference's own pytest files, tests/test_reference_suite_cpu.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(oracle):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import constriction_b200.stream as S
    from constriction_b200 import batch as B
    return dict(S=S, B=B, O=oracle)


def both_ways(env, make_model, symbols, params=()):
    """Encode with the mirror and with the oracle API (ANS and range): same words; decode with the mirror."""
    S, O = env["S"], env["O"]
    for api in (S, O):
        pass
    ms, mo = make_model(S), make_model(O)
    a, b = S.AnsCoder(), O.AnsCoder()
    a.encode_reverse(symbols, ms, *params)
    b.encode_reverse(symbols, mo, *params)
    assert np.array_equal(a.get_compressed(), b.get_compressed())
    got = a.decode(ms, *params) if params else a.decode(ms, len(symbols))
    assert np.array_equal(got, symbols) and a.is_empty()
    e, f = S.RangeEncoder(), O.RangeEncoder()
    e.encode(symbols, ms, *params)
    f.encode(symbols, mo, *params)
    assert np.array_equal(e.get_compressed(), f.get_compressed())
    d = S.RangeDecoder(e.get_compressed())
    got = d.decode(ms, *params) if params else d.decode(ms, len(symbols))
    assert np.array_equal(got, symbols)


def test_perfect_categorical_tables(env):
    B, O = env["B"], env["O"]
    rng = np.random.default_rng(41)
    for dtype in (np.float32, np.float64):
        pmf = rng.dirichlet(0.3 * np.ones(37), size=40).astype(dtype)
        pmf[0, :] = 1.0                      # all equal: ties everywhere
        pmf[1, 5:] = 0.0                     # zero-probability symbols keep weight 1
        pmf[2] = np.array([0.15, 0.69, 0.15] + [0.0] * 34, dtype=dtype)  # contiguous.rs:836-870 regression example
        got = B.ModelTable.categorical_perfect(pmf).cdf()
        for m in range(pmf.shape[0]):
            assert np.array_equal(got[m], O.cat_perfect_cdf(pmf[m])), (dtype, m)
    wide = rng.dirichlet(0.05 * np.ones(3000), size=2)
    got = B.ModelTable.categorical_perfect(wide).cdf()
    for m in range(2):
        assert np.array_equal(got[m], O.cat_perfect_cdf(wide[m]))
    with pytest.raises(ValueError):
        B.ModelTable.categorical_perfect(np.array([0.5, -0.1, 0.6]))
    with pytest.raises(ValueError):
        B.ModelTable.categorical_perfect(np.array([0.0, 0.0]))


def test_other_quantised_tables(env):
    B, O = env["B"], env["O"]
    rng = np.random.default_rng(42)
    locs = np.concatenate([[0.0, 3.2, -49.7, 1e3], rng.normal(0, 20, 60)])
    scales = np.concatenate([[1.0, 9.6, 1e-3, 50.0], np.exp(rng.uniform(np.log(0.01), np.log(200), 60))])
    for kind in ("laplace", "cauchy", "gaussian"):
        got = B.ModelTable.quantized(kind, -50, 50, locs, scales).cdf()
        for m in range(locs.size):
            assert np.array_equal(got[m], O.qdist_cdf(kind, -50, 50, locs[m], scales[m])), (kind, m)
    ns = np.array([1, 2, 10, 40, 100, 33], dtype=np.int32)
    ps = np.array([0.5, 0.0, 0.3, 0.5, 0.99, 1.0])
    got = B.ModelTable.binomial(ns, ps).cdf()
    for m in range(ns.size):
        want = O.binomial_cdf(int(ns[m]), ps[m])
        assert np.array_equal(got[m, : ns[m] + 2], want), m
        assert np.all(got[m, ns[m] + 1:] == 1 << 24)
    with pytest.raises(ValueError):
        B.ModelTable.quantized("laplace", -5, 5, [0.0], [0.0])


def test_default_categorical_and_bernoulli(env):
    rng = np.random.default_rng(43)
    probs = np.array([0.2, 0.4, 0.1, 0.3])
    syms = rng.choice(4, size=300, p=probs).astype(np.int32)
    both_ways(env, lambda api: api.Categorical(probs), syms)                # perfect=True by default
    both_ways(env, lambda api: api.Categorical(probs, perfect=True), syms)
    both_ways(env, lambda api: api.Categorical(probs.astype(np.float32)), syms)
    mat = rng.dirichlet(np.ones(5), size=40)
    s2 = np.array([rng.choice(5, p=row) for row in mat], dtype=np.int32)
    both_ways(env, lambda api: api.Categorical(), s2, (mat,))
    both_ways(env, lambda api: api.Categorical(perfect=True), s2, (mat.astype(np.float32),))
    bits = (rng.random(200) < 0.3).astype(np.int32)
    both_ways(env, lambda api: api.Bernoulli(0.3), bits)                     # perfect=True by default
    both_ways(env, lambda api: api.Bernoulli(0.3, perfect=False), bits)
    ps = rng.uniform(0.05, 0.95, size=200)
    both_ways(env, lambda api: api.Bernoulli(), bits, (ps,))


def test_laplace_cauchy_binomial_mirror(env):
    rng = np.random.default_rng(44)
    syms = np.clip(np.rint(rng.laplace(2.0, 6.0, size=200)), -100, 100).astype(np.int32)
    both_ways(env, lambda api: api.QuantizedLaplace(-100, 100, 2.0, 6.0), syms)
    both_ways(env, lambda api: api.QuantizedCauchy(-100, 100, 2.0, 6.0), syms)
    means = rng.normal(0, 10, size=200)
    scales = np.exp(rng.uniform(-1, 3, size=200))
    both_ways(env, lambda api: api.QuantizedLaplace(-100, 100), syms, (means, scales))
    both_ways(env, lambda api: api.QuantizedCauchy(-100, 100, None, 3.5), syms, (means.astype(np.float32),))
    ns = rng.integers(5, 60, size=100).astype(np.int32)
    ps = rng.uniform(0.1, 0.9, size=100)
    ks = rng.binomial(ns, ps).astype(np.int32)
    both_ways(env, lambda api: api.Binomial(), ks, (ns, ps))
    both_ways(env, lambda api: api.Binomial(60), np.minimum(ks, 60), (ps,))
    both_ways(env, lambda api: api.Binomial(40, 0.5), rng.binomial(40, 0.5, size=100).astype(np.int32))


def test_custom_and_scipy_models(env):
    scipy_stats = pytest.importorskip("scipy.stats")
    S = env["S"]
    # tests/python/test_docexamples_f32.py:855-868 (ScipyModel family over scipy.stats.cauchy, f32 parameters)
    fam = S.ScipyModel(scipy_stats.cauchy, -100, 100)
    symbols = np.array([22, 14, 5, -3, 19, 7], dtype=np.int32)
    locs = np.array([26.2, 10.9, 8.7, -6.3, 25.1, 8.9], dtype=np.float32)
    scales = np.array([4.3, 7.4, 2.9, 4.1, 9.7, 3.4], dtype=np.float32)
    coder = S.AnsCoder()
    coder.encode_reverse(symbols, fam, locs, scales)
    assert np.array_equal(coder.get_compressed(), np.array([3611353862, 17526], dtype=np.uint32))
    assert np.array_equal(coder.decode(fam, locs, scales), symbols)
    rng = np.random.default_rng(45)
    syms = np.clip(np.rint(rng.normal(1.2, 4.9, size=150)), -100, 100).astype(np.int32)
    frozen = scipy_stats.norm(1.2, 4.9)
    both_ways(env, lambda api: api.CustomModel(frozen.cdf, frozen.ppf, -100, 100), syms)
    both_ways(env, lambda api: api.ScipyModel(frozen, -100, 100), syms)
    locs = rng.normal(0, 5, size=150)
    both_ways(env, lambda api: api.CustomModel(lambda x, loc, scale: scipy_stats.norm.cdf(x, loc, scale),
                                               lambda q, loc, scale: scipy_stats.norm.ppf(q, loc, scale), -100, 100),
              syms, (locs, np.full(150, 4.9)))


def test_impossible_symbol_keeps_the_symbols_coded_before_it(env):
    """Reference semantics after KeyError (stream/mod.rs:592-607: symbol-by-symbol loops): the coder holds the symbols
    that were coded before the impossible one -- the LAST ones for AnsCoder.encode_reverse, the FIRST ones for
    RangeEncoder.encode.  Same state in the mirror and in the oracle's restatement."""
    S, O = env["S"], env["O"]
    good = np.array([3, -7, 12, 0, 5, -1], dtype=np.int32)
    bad = good.copy()
    bad[2] = 99
    means = np.linspace(-3, 3, 6)
    stds = np.linspace(2, 8, 6)
    probs = np.array([0.25, 0.0, 0.5, 0.25])
    cases = [(lambda api: api.QuantizedGaussian(-50, 50, 1.0, 6.0), bad, ()),
             (lambda api: api.QuantizedGaussian(-50, 50), bad, (means, stds)),
             (lambda api: api.Categorical(probs, perfect=False), np.array([0, 2, 3, 9, 2, 0], dtype=np.int32), ())]
    for make, syms, params in cases:
        states = []
        for api in (S, O):
            a, e = api.AnsCoder(), api.RangeEncoder()
            with pytest.raises(KeyError):
                a.encode_reverse(syms, make(api), *params)
            with pytest.raises(KeyError):
                e.encode(syms, make(api), *params)
            states.append((a.get_compressed(), e.get_compressed(), a.pos(), e.pos()))
        assert np.array_equal(states[0][0], states[1][0]) and np.array_equal(states[0][1], states[1][1])
        assert states[0][2] == states[1][2] and states[0][3] == states[1][3]
        assert states[0][0].size > 0  # something was kept
