"""Per-symbol QuantizedGaussian parameters evaluated on the device (ctr_*_gaussian):
the table-free kernels must produce, word for word, what the oracle produces when it builds one
LeakilyQuantizedDistribution per symbol (reference: pybindings/stream/model/internals.rs:188-249,
quantize.rs:525-568,580-779), and what the tabulated path (one CDF row per symbol) produces."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LO, HI = -40, 60


@pytest.fixture(scope="module")
def env(oracle):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from constriction_b200 import batch as B
    return dict(torch=torch, B=B, bc=B.BatchCoder(), O=oracle)


def make_inputs(seed, n, lo=LO, hi=HI):
    rng = np.random.default_rng(seed)
    means = rng.normal(5.0, 25.0, size=n)                             # some far outside the support
    stds = np.exp(rng.uniform(np.log(1e-3), np.log(200.0), size=n))   # needle-sharp to nearly flat
    means[:4] = [lo - 0.5, hi + 0.5, 0.0, 1e9]
    stds[:4] = [1e-30, 1e-30, 1e30, 1.0]
    syms = np.clip(np.rint(rng.normal(means, stds)), lo, hi).astype(np.int32)
    # a few symbols in the far tails of their model (leaky probability 1 / 2^24)
    syms[4:12] = [lo, hi, lo + 1, hi - 1, lo, hi, lo + 7, hi - 7]
    return syms, means, stds


def dev(env, a):
    return env["torch"].from_numpy(np.ascontiguousarray(a)).cuda()


def oracle_cdfs(O, means, stds, lo=LO, hi=HI):
    return np.stack([O.qgauss_cdf(lo, hi, float(m), float(s)) for m, s in zip(means, stds)])


@pytest.mark.parametrize("coder", ["ans", "range"])
@pytest.mark.parametrize("k,contig", [(1, False), (33, False), (256, False), (7, True), (300, True)])
def test_lazy_equals_tables(env, coder, k, contig):
    """GaussianParams (no tables) == ModelTable with one row per symbol + model index, in every layout."""
    B, bc, torch = env["B"], env["bc"], env["torch"]
    n = 30_000
    syms, means, stds = make_inputs(11 + k, n)
    d_syms = dev(env, syms)
    lazy = B.GaussianParams(LO, HI, means, stds)
    table = B.ModelTable.quantized_gaussian(LO, HI, means, stds)
    idx = torch.arange(n, dtype=torch.int32, device="cuda")
    sym_off = None
    if contig:
        rng = np.random.default_rng(k)
        cuts = np.sort(rng.integers(0, n + 1, size=k - 1))
        sym_off = dev(env, np.concatenate([[0], cuts, [n]]).astype(np.int64))  # ragged, some streams empty
    enc = bc.ans_encode if coder == "ans" else bc.range_encode
    dec = bc.ans_decode if coder == "ans" else bc.range_decode
    c_lazy = enc(d_syms, lazy, n_streams=k, sym_offsets=sym_off)
    c_tab = enc(d_syms, table, n_streams=k, sym_offsets=sym_off, model_index=idx)
    bc.check()
    w1, o1 = c_lazy.to_host()
    w2, o2 = c_tab.to_host()
    assert np.array_equal(o1, o2) and np.array_equal(w1, w2)
    out = dec(c_lazy, lazy)
    bc.check()
    assert np.array_equal(out.cpu().numpy(), syms)
    out2 = dec(c_lazy, table, model_index=idx)
    bc.check()
    assert np.array_equal(out2.cpu().numpy(), syms)


def test_lazy_equals_oracle_ans(env):
    """One stream: the words equal the oracle's coder fed one quantised Gaussian per symbol."""
    B, bc, O = env["B"], env["bc"], env["O"]
    n = 6000
    syms, means, stds = make_inputs(3, n)
    cdfs = oracle_cdfs(O, means, stds)
    want = O.ans_encode_indexed(syms, np.arange(n), cdfs, LO)
    lazy = B.GaussianParams(LO, HI, means, stds)
    comp = bc.ans_encode(dev(env, syms), lazy, n_streams=1)
    bc.check()
    words, off = comp.to_host()
    assert np.array_equal(words, want)
    assert np.array_equal(bc.ans_decode(comp, lazy).cpu().numpy(), O.ans_decode_indexed(want, np.arange(n), cdfs, LO))


def test_lazy_equals_oracle_api_both_coders(env, monkeypatch):
    """Through the mirror of the reference's Python API, against the oracle's restatement of that API."""
    import constriction_b200.stream as S
    O = env["O"]
    n = 1500
    syms, means, stds = make_inputs(4, n)
    for dtype in (np.float64, np.float32):
        m, s = means.astype(dtype), np.maximum(stds, 1e-20).astype(dtype)
        fam, ofam = S.model.QuantizedGaussian(LO, HI), O.QuantizedGaussian(LO, HI)
        a, oa = S.stack.AnsCoder(), O.AnsCoder()
        a.encode_reverse(syms, fam, m, s)
        oa.encode_reverse(syms, ofam, m, s)
        assert np.array_equal(a.get_compressed(), oa.get_compressed())
        assert np.array_equal(a.decode(fam, m, s), syms)
        r, orr = S.queue.RangeEncoder(), O.RangeEncoder()
        r.encode(syms, fam, m, s)
        orr.encode(syms, ofam, m, s)
        assert np.array_equal(r.get_compressed(), orr.get_compressed())
        assert np.array_equal(S.queue.RangeDecoder(r.get_compressed()).decode(fam, m, s), syms)
        # mean fixed in the family, std per symbol (pybindings/stream/model.rs:682-700)
        fam2, ofam2 = S.model.QuantizedGaussian(LO, HI, mean=1.5), O.QuantizedGaussian(LO, HI, mean=1.5)
        a2, oa2 = S.stack.AnsCoder(), O.AnsCoder()
        sy2 = np.clip(syms, LO, HI)
        a2.encode_reverse(sy2, fam2, s)
        oa2.encode_reverse(sy2, ofam2, s)
        assert np.array_equal(a2.get_compressed(), oa2.get_compressed())
    # the tabulated path through the same API gives the same words
    monkeypatch.setenv("CTR_GAUSS_TABLES", "1")
    b = S.stack.AnsCoder()
    b.encode_reverse(syms, S.model.QuantizedGaussian(LO, HI), means, np.maximum(stds, 1e-20))
    c = S.stack.AnsCoder()
    monkeypatch.delenv("CTR_GAUSS_TABLES")
    c.encode_reverse(syms, S.model.QuantizedGaussian(LO, HI), means, np.maximum(stds, 1e-20))
    assert np.array_equal(b.get_compressed(), c.get_compressed())


def test_wide_support_and_errors(env):
    B, bc = env["B"], env["bc"]
    rng = np.random.default_rng(9)
    n, lo, hi = 20_000, -30_000, 30_000  # 60001 symbols: far too wide for a table per symbol (14 GB)
    means = rng.normal(0, 8000, size=n)
    stds = np.exp(rng.uniform(np.log(0.05), np.log(5000.0), size=n))
    syms = np.clip(np.rint(rng.normal(means, stds)), lo, hi).astype(np.int32)
    syms[:6] = [lo, hi, lo + 1, hi - 1, 0, 12345]  # leaky tails: the search gallops far from its starting point
    lazy = B.GaussianParams(lo, hi, means, stds)
    for enc, dec in ((bc.ans_encode, bc.ans_decode), (bc.range_encode, bc.range_decode)):
        comp = enc(dev(env, syms), lazy, n_streams=64)
        out = dec(comp, lazy)
        bc.check()
        assert np.array_equal(out.cpu().numpy(), syms)
    # impossible symbol -> KeyError, std <= 0 -> ValueError (pybindings/stream/model.rs:654-657)
    bad = syms.copy()
    bad[100] = hi + 1
    bc.ans_encode(dev(env, bad), lazy, n_streams=64)
    with pytest.raises(KeyError):
        bc.check()
    s0 = stds.copy()
    s0[7] = 0.0
    bc.ans_encode(dev(env, syms), B.GaussianParams(lo, hi, means, s0), n_streams=64)
    with pytest.raises(ValueError):
        bc.check()
    comp = bc.ans_encode(dev(env, syms), lazy, n_streams=64)
    bc.ans_decode(comp, B.GaussianParams(lo, hi, means, s0))
    with pytest.raises(ValueError):
        bc.check()
    with pytest.raises(ValueError):
        bc.ans_encode(dev(env, syms[:10]), lazy, n_streams=2)  # parameter count != symbol count


def test_learned_compression_shape(env):
    """1.5e6 latents, every one with its own (mean, std) from a hyperprior: round trip, both coders."""
    B, bc, torch = env["B"], env["bc"], env["torch"]
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    n, k = 1_572_864, 8 * 192
    means = torch.randn(n, device="cuda", generator=g, dtype=torch.float64) * 2.0
    stds = torch.exp(torch.rand(n, device="cuda", generator=g, dtype=torch.float64) * 3.7 - 1.2)
    syms = torch.clamp(torch.round(means + stds * torch.randn(n, device="cuda", generator=g, dtype=torch.float64)), -64, 64).to(torch.int32)
    lazy = B.GaussianParams(-64, 64, means, stds)
    for enc, dec in ((bc.ans_encode, bc.ans_decode), (bc.range_encode, bc.range_decode)):
        comp = enc(syms, lazy, n_streams=k)
        out = dec(comp, lazy)
        bc.check()
        assert torch.equal(out, syms)


def test_bernoulli_mirror(env):
    """Bernoulli(perfect=False) = two-symbol categorical [1 - p, p] (pybindings/stream/model.rs:985-1060)."""
    import constriction_b200.stream as S
    O = env["O"]
    rng = np.random.default_rng(12)
    n = 4000
    ps = rng.uniform(0.0, 1.0, size=n)
    ps[:3] = [0.0, 1.0, 0.5]
    bits = (rng.uniform(size=n) < ps).astype(np.int32)
    a = S.stack.AnsCoder()
    a.encode_reverse(bits, S.model.Bernoulli(perfect=False), ps)
    oa = O.AnsCoder()
    oa.encode_reverse(bits, O.Categorical(perfect=False), np.stack([1.0 - ps, ps], axis=1))
    assert np.array_equal(a.get_compressed(), oa.get_compressed())
    assert np.array_equal(a.decode(S.model.Bernoulli(perfect=False), ps), bits)
    m = S.model.Bernoulli(0.2, perfect=False)
    r = S.queue.RangeEncoder()
    r.encode(bits, m)
    orr = O.RangeEncoder()
    orr.encode(bits, O.Categorical(np.array([0.8, 0.2]), perfect=False))
    assert np.array_equal(r.get_compressed(), orr.get_compressed())
    assert np.array_equal(S.queue.RangeDecoder(r.get_compressed()).decode(m, n), bits)
    # like the reference's fast quantiser, p outside [0, 1] is only rejected when it breaks the normalisation
    # (categorical.rs:38-41); [-0.5, 1.5] sums to one and quantises to a valid model
    S.model.Bernoulli(1.5, perfect=False)
    with pytest.raises(ValueError):
        S.model.Bernoulli(float("nan"), perfect=False)


def test_uniform_family_mirror(env):
    """Uniform() with one int32 `size` per symbol (pybindings/stream/model.rs:570-600, uniform.rs:44-146)."""
    import constriction_b200.stream as S
    O = env["O"]
    rng = np.random.default_rng(13)
    n = 3000
    sizes = rng.integers(2, 300, size=n).astype(np.int32)
    sizes[:3] = [2, 299, 7]
    syms = (rng.uniform(size=n) * sizes).astype(np.int32)
    syms[:3] = [1, 298, 0]
    a, oa = S.stack.AnsCoder(), O.AnsCoder()
    a.encode_reverse(syms, S.model.Uniform(), sizes)
    oa.encode_reverse(syms, O.Uniform(), sizes)
    assert np.array_equal(a.get_compressed(), oa.get_compressed())
    assert np.array_equal(a.decode(S.model.Uniform(), sizes), syms)
    r, orr = S.queue.RangeEncoder(), O.RangeEncoder()
    r.encode(syms, S.model.Uniform(), sizes)
    orr.encode(syms, O.Uniform(), sizes)
    assert np.array_equal(r.get_compressed(), orr.get_compressed())
    assert np.array_equal(S.queue.RangeDecoder(r.get_compressed()).decode(S.model.Uniform(), sizes), syms)
    bad = syms.copy()
    bad[5] = sizes[5]  # one past the symbol's alphabet
    with pytest.raises(KeyError):
        S.stack.AnsCoder().encode_reverse(bad, S.model.Uniform(), sizes)
