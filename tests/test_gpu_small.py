"""The Small preset (u16 / u32 / 12) with lookup decoder models on the CUDA path: every
stream of a batch equals what the oracle's generic-preset restatement produces for one SmallAnsCoder /
SmallRangeEncoder (tests/test_oracle_generic.py pins that restatement through the Default preset)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(oracle):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from constriction_b200 import small as S
    return dict(torch=torch, S=S, O=oracle, bc=S.SmallBatchCoder())


def draw(rng, cdf, n):
    p = np.diff(cdf.astype(np.int64)).astype(np.float64)
    return rng.choice(p.size, size=n, p=p / p.sum()).astype(np.int32)


@pytest.mark.parametrize("perfect", [False, True])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_small_models_equal_the_oracle(env, perfect, dtype):
    S, O = env["S"], env["O"]
    rng = np.random.default_rng(7)
    pmf = rng.dirichlet(0.4 * np.ones(100), size=30).astype(dtype)
    pmf[0, :] = 1.0
    pmf[1, 3:] = 0.0
    got = S.SmallModel.categorical(pmf, perfect=perfect).cdf()
    for m in range(pmf.shape[0]):
        assert np.array_equal(got[m], O.g_cat_cdf("small", pmf[m], perfect=perfect)), m
    with pytest.raises(ValueError):
        S.SmallModel.from_cdf(np.array([0, 10, 5, 4096], dtype=np.uint16))       # decreasing
    with pytest.raises(ValueError):
        S.SmallModel.from_cdf(np.array([0, 10, 4095], dtype=np.uint16))           # does not end at 2^12


@pytest.mark.parametrize("coder", ["ans", "range"])
@pytest.mark.parametrize("n,k", [(10_000, 1), (10_000, 100), (100_003, 384), (50, 200), (0, 3), (300_000, 2048)])
def test_small_interleaved_matches_the_oracle(env, coder, n, k):
    S, O, bc, torch = env["S"], env["O"], env["bc"], env["torch"]
    rng = np.random.default_rng(n + k)
    pmf = rng.dirichlet(0.5 * np.ones(100)).astype(np.float32)   # benches/lookup.rs: a 100-symbol categorical model
    model = S.SmallModel.categorical(pmf, perfect=True)
    cdf = model.cdf()[0]
    syms = draw(rng, cdf, n)
    d = torch.from_numpy(syms).cuda()
    comp = (bc.ans_encode if coder == "ans" else bc.range_encode)(d, model, n_streams=k)
    out = (bc.ans_decode if coder == "ans" else bc.range_decode)(comp, model)
    bc.check()
    assert np.array_equal(out.cpu().numpy(), syms)
    words, off = comp.to_host()
    enc1 = O.g_ans_encode if coder == "ans" else O.g_range_encode
    for s in sorted(set([0, 1, k // 2, k - 1])):
        if s < k:
            want = enc1("small", syms[s::k], cdf)
            assert np.array_equal(words[int(off[s]):int(off[s + 1])], want.astype(np.uint16)), s


@pytest.mark.parametrize("coder", ["ans", "range"])
def test_small_contiguous_per_stream_models_and_noncontiguous_alphabet(env, coder):
    S, O, bc, torch = env["S"], env["O"], env["bc"], env["torch"]
    rng = np.random.default_rng(11)
    k = 300
    pmf = rng.dirichlet(0.6 * np.ones(37), size=5)
    model = S.SmallModel.categorical(pmf)              # fast quantisation, f64
    cdfs = model.cdf()
    lens = rng.integers(0, 400, size=k)
    lens[[0, 9]] = 0
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = rng.integers(0, 5, size=k).astype(np.int32)
    syms = np.concatenate([draw(rng, cdfs[idx[s]], lens[s]) for s in range(k)] + [np.empty(0, np.int32)]).astype(np.int32)
    comp = (bc.ans_encode if coder == "ans" else bc.range_encode)(torch.from_numpy(syms).cuda(), model, sym_offsets=torch.from_numpy(off).cuda(),
                                                                  model_index=torch.from_numpy(idx).cuda())
    out = (bc.ans_decode if coder == "ans" else bc.range_decode)(comp, model, model_index=torch.from_numpy(idx).cuda())
    bc.check()
    assert np.array_equal(out.cpu().numpy(), syms)
    enc1 = O.g_ans_encode if coder == "ans" else O.g_range_encode
    for s in range(0, k, 13):
        want = enc1("small", syms[off[s]:off[s + 1]], cdfs[idx[s]])
        assert np.array_equal(comp.stream_words(s), want.astype(np.uint16)), s
    # non-contiguous alphabet (lookup_noncontiguous.rs): index i stands for an arbitrary symbol
    alphabet = np.array([-7, 1000, 3, 42, -100000], dtype=np.int32)
    cdf = O.g_cat_cdf("small", np.array([0.3, 0.1, 0.2, 0.25, 0.15]), perfect=True)
    nc = S.SmallModel.from_cdf(cdf.astype(np.uint16), symbols=alphabet)
    ids = draw(rng, cdf, 5000)
    msg = alphabet[ids]
    comp = (bc.ans_encode if coder == "ans" else bc.range_encode)(torch.from_numpy(msg).cuda(), nc, n_streams=16)
    out = (bc.ans_decode if coder == "ans" else bc.range_decode)(comp, nc)
    bc.check()
    assert np.array_equal(out.cpu().numpy(), msg)
    assert np.array_equal(comp.stream_words(3), enc1("small", ids[3::16], cdf).astype(np.uint16))
    bad = msg.copy()
    bad[100] = 5  # not in the alphabet
    (bc.ans_encode if coder == "ans" else bc.range_encode)(torch.from_numpy(bad).cuda(), nc, n_streams=16)
    with pytest.raises(KeyError):
        bc.check()


def test_small_lookup_doc_example(env):
    """lookup_contiguous.rs:55-103: message, probabilities and the perfect quantisation of the doc example."""
    S, O, bc, torch = env["S"], env["O"], env["bc"], env["torch"]
    message = np.array([2, 1, 3, 0, 0, 2, 0, 2, 1, 0, 2], dtype=np.int32)
    model = S.SmallModel.categorical(np.array([0.4, 0.2, 0.1, 0.3], dtype=np.float32), perfect=True)
    cdf = model.cdf()[0]
    comp = bc.range_encode(torch.from_numpy(message).cuda(), model, n_streams=1)
    assert np.array_equal(comp.stream_words(0), O.g_range_encode("small", message, cdf).astype(np.uint16))
    # "fixed_point_probabilities" -> lookup decoder model, as in the example
    lookup = S.SmallModel.from_cdf(np.concatenate([[0], np.cumsum(np.diff(cdf.astype(np.int64)))]).astype(np.uint16))
    assert np.array_equal(bc.range_decode(comp, lookup).cpu().numpy(), message)
    bc.check()
