"""Replays tests/golden/reference_vectors.py against an implementation of the
reference's Python API (`api` = oracle.refapi or constriction_b200.stream facade)."""
import numpy as np

from golden import reference_vectors as G


def make_model(api, spec):
    if spec[0] == "qgauss":
        _, lo, hi, mean, std = spec
        return api.QuantizedGaussian(lo, hi, mean, std)
    if spec[0] == "cat":
        _, probs, kw = spec
        return api.Categorical(probs, **kw) if probs is not None else api.Categorical(**kw)
    raise ValueError(spec)


def run_encode_case(api, case):
    segs = case["segments"]
    if case["coder"] == "ans":
        coder = api.AnsCoder()
        for spec, syms, params in reversed(segs):
            coder.encode_reverse(syms, make_model(api, spec), *params)
        words = coder.get_compressed()
        assert words.dtype == np.uint32
        assert np.array_equal(words, case["words"]), (case["id"], words, case["words"])
        assert coder.num_words() == len(case["words"])
        # decode from a fresh coder and from the encoder itself
        for dec in (api.AnsCoder(words), coder):
            for spec, syms, params in segs:
                m = make_model(api, spec)
                got = dec.decode(m, *params) if params else dec.decode(m, len(syms))
                assert got.dtype == np.int32
                assert np.array_equal(got, syms), (case["id"], got, syms)
            assert dec.is_empty()
    else:
        enc = api.RangeEncoder()
        for spec, syms, params in segs:
            enc.encode(syms, make_model(api, spec), *params)
        words = enc.get_compressed()
        assert words.dtype == np.uint32
        assert np.array_equal(words, case["words"]), (case["id"], words, case["words"])
        assert enc.num_words() == len(case["words"])
        for dec in (api.RangeDecoder(words), enc.get_decoder()):
            for spec, syms, params in segs:
                m = make_model(api, spec)
                got = dec.decode(m, *params) if params else dec.decode(m, len(syms))
                assert np.array_equal(got, syms), (case["id"], got, syms)
            assert dec.maybe_exhausted()


def run_decode_case(api, case):
    cid, src, coder, spec, arg, compressed, expected = case
    compressed = np.array(compressed, dtype=np.uint32)
    dec = api.AnsCoder(compressed) if coder == "ans" else api.RangeDecoder(compressed)
    m = make_model(api, spec)
    if arg is None:
        got = dec.decode(m)
        assert got == expected, (cid, got)
    elif isinstance(arg, int):
        got = dec.decode(m, arg)
        assert np.array_equal(got, np.array(expected, dtype=np.int32)), (cid, got)
    else:
        got = dec.decode(m, *arg)
        assert np.array_equal(got, np.array(expected, dtype=np.int32)), (cid, got)


def run_seal_case(api):
    """G17: tests/python/test_constriction.py:102-117."""
    model = api.Categorical(perfect=False)
    for row0, expected in G.SEAL_EXPECT:
        probs = G.SEAL_PROBS.copy()
        probs[0, :] = row0
        coder = api.AnsCoder(G.SEAL_DATA, True)
        assert np.array_equal(coder.decode(model, probs), expected)
    # seal -> decode -> re-encode -> unseal restores the data (stack.rs:341-360, 944-955)
    probs = G.SEAL_PROBS.copy()
    coder = api.AnsCoder(G.SEAL_DATA, seal=True)
    syms = coder.decode(model, probs)
    coder.encode_reverse(syms, model, probs)
    assert np.array_equal(coder.get_compressed(unseal=True), G.SEAL_DATA)


def run_length_cases(api):
    """G18: src/stream/stack.rs:1250-1291, src/stream/queue.rs:1133-1173."""
    model = api.QuantizedGaussian(-127, 127, 3.2, 5.1)
    for syms, nwords in G.LENGTH_CASES:
        syms = np.array(syms, dtype=np.int32)
        ans = api.AnsCoder()
        ans.encode_reverse(syms[::-1].copy(), model)  # reference encodes forward, decodes in reverse
        assert ans.num_words() == nwords
        words = ans.get_compressed()
        assert len(words) == nwords
        dec = api.AnsCoder(words)
        assert np.array_equal(dec.decode(model, len(syms)), syms[::-1])
        assert dec.is_empty()
        enc = api.RangeEncoder()
        enc.encode(syms, model)
        words = enc.get_compressed()
        assert len(words) == nwords and enc.num_words() == nwords
        dec = api.RangeDecoder(words)
        assert np.array_equal(dec.decode(model, len(syms)), syms)
        assert dec.maybe_exhausted()
    # compress_none (stack.rs:1239-1248, queue.rs:1122-1131)
    ans = api.AnsCoder()
    assert ans.is_empty() and len(ans.get_compressed()) == 0 and ans.num_words() == 0
    assert api.AnsCoder(ans.get_compressed()).is_empty()
    enc = api.RangeEncoder()
    assert enc.is_empty() and len(enc.get_compressed()) == 0
    assert api.RangeDecoder(enc.get_compressed()).maybe_exhausted()


def run_misc_cases(api):
    # num_bits / num_valid_bits: tests/python/test_constriction.py:41-42
    case = next(c for c in G.ENCODE_CASES if c["id"] == "G1_ans_gauss_params_f64")
    spec, syms, params = case["segments"][0]
    ans = api.AnsCoder()
    ans.encode_reverse(syms, make_model(api, spec), *params)
    assert ans.num_bits() == 64 and ans.num_valid_bits() == 51

    # ANS seek: tests/python/test_docexamples.py:403-425
    model = api.Categorical(np.array([0.2, 0.4, 0.1, 0.3], dtype=np.float64), perfect=False)
    part1 = np.array([1, 2, 0, 3, 2, 3, 0], dtype=np.int32)
    part2 = np.array([2, 2, 0, 1, 3], dtype=np.int32)
    coder = api.AnsCoder()
    coder.encode_reverse(part2, model)
    position, state = coder.pos()
    coder.encode_reverse(part1, model)
    assert coder.decode(model) == 1
    coder.seek(position, state)
    assert np.array_equal(coder.decode(model, 5), part2)

    # range seek: tests/python/test_docexamples.py:619-640
    enc = api.RangeEncoder()
    enc.encode(part1, model)
    position, state = enc.pos()
    enc.encode(part2, model)
    dec = api.RangeDecoder(enc.get_compressed())
    assert dec.decode(model) == 1
    dec.seek(position, state)
    assert np.array_equal(dec.decode(model, 5), part2)

    # scalar symbol API: tests/python/test_docexamples.py:345-353, 533-542
    m3 = api.Categorical(np.array([0.1, 0.6, 0.3], dtype=np.float64), perfect=False)
    ans = api.AnsCoder()
    ans.encode_reverse(2, m3)
    assert ans.decode(m3) == 2
    enc = api.RangeEncoder()
    enc.encode(2, m3)
    assert api.RangeDecoder(enc.get_compressed()).decode(m3) == 2

    # error mapping (pybindings/stream/mod.rs:83-91, stack.rs:230-235)
    import pytest
    with pytest.raises(KeyError):
        api.AnsCoder().encode_reverse(np.array([3], dtype=np.int32), m3)
    with pytest.raises(KeyError):
        api.RangeEncoder().encode(np.array([-1], dtype=np.int32), m3)
    with pytest.raises(KeyError):
        api.AnsCoder().encode_reverse(np.array([51], dtype=np.int32), api.QuantizedGaussian(-50, 50, 0.0, 1.0))
    with pytest.raises(ValueError):
        api.AnsCoder(np.array([1, 0], dtype=np.uint32))
    with pytest.raises(ValueError):
        api.AnsCoder(None, True)
    with pytest.raises(ValueError):
        api.AnsCoder().encode_reverse(np.array([1, 2], dtype=np.int32), api.QuantizedGaussian(-5, 5),
                                      np.array([0.0]), np.array([1.0]))

    # mixed ANS stream: tests/python/test_docexamples.py:204-246 (round trip only, no golden there)
    fam = api.QuantizedGaussian(-100, 100)
    msg2 = np.array([6, 10, -4, 2], dtype=np.int32)
    means2 = np.array([2.5, 13.1, -1.1, -3.0])
    stds2 = np.array([4.1, 8.7, 6.2, 5.4])
    coder = api.AnsCoder()
    coder.encode_reverse(msg2, fam, means2, stds2)
    coder.encode_reverse(part1, model)
    dec = api.AnsCoder(coder.get_compressed())
    assert np.array_equal(dec.decode(model, 7), part1)
    assert np.array_equal(dec.decode(fam, means2, stds2), msg2)
    assert dec.is_empty()

    # decoding past the end of an ANS stream is allowed and deterministic (stack.rs:1062-1065)
    empty = api.AnsCoder()
    assert np.array_equal(empty.decode(api.QuantizedGaussian(-5, 5, 0.0, 2.0), 3), [-5, -5, -5])
