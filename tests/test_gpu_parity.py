"""Parity of the batched CUDA kernels (through the C ABI) with the CPU oracle: word-for-word equal
compressed streams and symbol-for-symbol equal decodes, on seeded inputs the oracle finishes in
seconds, plus size-independent properties at BASELINE.json's full size."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BASE = (-50, 50, 3.2, 9.6)  # BASELINE.json's model: QuantizedGaussian(-50, 50, 3.2, 9.6)


@pytest.fixture(scope="module")
def env(oracle):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from constriction_b200 import batch as B
    return dict(torch=torch, B=B, bc=B.BatchCoder(), O=oracle)


def gauss_symbols(rng, n, mean=3.2, std=9.6, lo=-50, hi=50):
    return np.clip(np.rint(rng.normal(mean, std, size=n)), lo, hi).astype(np.int32)


def dev(env, a, dtype=None):
    t = env["torch"].from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


# ---------------------------------------------------------------------------------------------------
# model tabulation (K5)
# ---------------------------------------------------------------------------------------------------
def test_gaussian_tables(env):
    B, O = env["B"], env["O"]
    rng = np.random.default_rng(5)
    means = np.concatenate([[3.2, 0.0, -49.9, 50.0, 1e3], rng.normal(0, 20, 200)])
    stds = np.concatenate([[9.6, 1e-40, 0.3, 1e6, 5.0], np.exp(rng.uniform(np.log(0.01), np.log(100), 200))])
    t = B.ModelTable.quantized_gaussian(-50, 50, means, stds)
    got = t.cdf()
    for m in range(means.size):
        assert np.array_equal(got[m], O.qgauss_cdf(-50, 50, means[m], stds[m])), m
    with pytest.raises(ValueError):
        B.ModelTable.quantized_gaussian(-5, 5, [0.0], [0.0])
    with pytest.raises(ValueError):
        B.ModelTable.quantized_gaussian(5, 5, [0.0], [1.0])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_categorical_tables(env, dtype):
    B, O = env["B"], env["O"]
    rng = np.random.default_rng(6)
    pmf = rng.dirichlet(0.5 * np.ones(256), size=300).astype(dtype)
    pmf[0, :] = 1.0
    pmf[1, 5:] = 0.0  # zero-probability symbols still get the leaky minimum
    got = B.ModelTable.categorical(pmf).cdf()
    for m in range(pmf.shape[0]):
        assert np.array_equal(got[m], O.cat_cdf(pmf[m])), m
    # device-resident pmf
    got2 = B.ModelTable.categorical(env["torch"].from_numpy(pmf).cuda()).cdf()
    assert np.array_equal(got, got2)
    bad = pmf[:2].copy()
    bad[1, :] = 0.0
    with pytest.raises(ValueError):
        B.ModelTable.categorical(bad)
    nan = pmf[:2].copy()
    nan[0, 3] = np.nan
    with pytest.raises(ValueError):
        B.ModelTable.categorical(nan)


def test_uniform_and_cdf_tables(env):
    B = env["B"]
    u = B.ModelTable.uniform(10).cdf()[0]
    per = (1 << 24) // 10
    assert list(u[:10]) == [i * per for i in range(10)] and u[10] == 1 << 24
    cdf = np.array([[0, 100, 100, 1 << 23, 1 << 24]], dtype=np.uint32)  # one zero-probability symbol
    assert np.array_equal(B.ModelTable.from_cdf(cdf, min_symbol=-2).cdf(), cdf)
    with pytest.raises(ValueError):
        B.ModelTable.from_cdf(np.array([[0, 5, 3, 1 << 24]], dtype=np.uint32))
    with pytest.raises(ValueError):
        B.ModelTable.from_cdf(np.array([[1, 5, 1 << 24]], dtype=np.uint32))


# ---------------------------------------------------------------------------------------------------
# ANS / range, interleaved layout, one shared model (BASELINE configs 1, 2)
# ---------------------------------------------------------------------------------------------------
SHAPES = [(0, 1), (1, 1), (5, 1), (100_000, 1), (1000, 32), (1000, 33), (31, 64), (100_003, 4096), (50_000, 777),
          (200_000, 128 * 3 + 5),
          # CTA sizes: < 37,888 streams -> 64-thread CTAs; up to 151,552 -> 256; beyond -> 1024-thread decoders
          (400_000, 40_001), (1_000_000, 151_552 + 7),
          # TMA symbol boxes (K a multiple of 4): last warp partly outside the tensor, rows above the highest box,
          # exactly one box, one row short of a box (per-row path)
          (300_000, 100), (2_000_003, 148 * 256 + 4), (9 * 64, 64), (8 * 64 + 5, 64)]


@pytest.mark.parametrize("coder", ["ans", "range"])
@pytest.mark.parametrize("n,k", SHAPES)
def test_interleaved_iid_matches_oracle(env, coder, n, k):
    B, O, bc = env["B"], env["O"], env["bc"]
    rng = np.random.default_rng(n * 31 + k)
    syms = gauss_symbols(rng, n)
    if n > 50:
        syms[::97] = rng.integers(-50, 51, size=syms[::97].size)  # tails: probability 1..few
    model = B.ModelTable.quantized_gaussian(*BASE[:2], [BASE[2]], [BASE[3]])
    cdf = model.cdf()[0]
    enc_o, dec_o = (O.multi_ans_encode, O.multi_ans_decode) if coder == "ans" else (O.multi_range_encode, O.multi_range_decode)
    want_words, want_off = enc_o(syms, k, cdf, -50, threads=8)
    d_syms = dev(env, syms) if n else env["torch"].zeros(0, dtype=env["torch"].int32, device="cuda")
    comp = (bc.ans_encode if coder == "ans" else bc.range_encode)(d_syms, model, n_streams=k)
    words, off = comp.to_host()
    bc.check()
    assert np.array_equal(off, want_off)
    assert np.array_equal(words, want_words)
    out = (bc.ans_decode if coder == "ans" else bc.range_decode)(comp, model)
    bc.check()
    assert np.array_equal(out.cpu().numpy(), syms)
    assert np.array_equal(dec_o(want_words, want_off, n, k, cdf, -50, threads=8), syms)


@pytest.mark.parametrize("lo,hi,mean,std", [(-500, 500, 20.5, 100.0), (-2000, 2000, -3.3, 3.0), (0, 255, 100.0, 30.0)])
def test_large_batch_ans_decoder_quantile_index(env, lo, hi, mean, std):
    """>= 148 * 1024 streams of one shared model: the one-CTA-per-SM ANS decoder with the fine quantile index
    (device_utils.cuh: kBigLutBits), for alphabets above and below 256 symbols, with tail symbols that share an index
    bucket with many others (second probe and cold search).  Words and symbols equal the oracle's."""
    B, O, bc = env["B"], env["O"], env["bc"]
    n, k = 700_000, 148 * 1024 + 40
    rng = np.random.default_rng(hi * 7 + 1)
    syms = gauss_symbols(rng, n, mean, std, lo, hi)
    syms[::5] = rng.integers(lo, hi + 1, size=syms[::5].size)  # every symbol of the alphabet, however improbable
    model = B.ModelTable.quantized_gaussian(lo, hi, [mean], [std])
    cdf = model.cdf()[0]
    want_words, want_off = O.multi_ans_encode(syms, k, cdf, lo, threads=8)
    comp = bc.ans_encode(dev(env, syms), model, n_streams=k)
    words, off = comp.to_host()
    bc.check()
    assert np.array_equal(off, want_off)
    assert np.array_equal(words, want_words)
    out = bc.ans_decode(comp, model)
    bc.check()
    assert np.array_equal(out.cpu().numpy(), syms)


@pytest.mark.parametrize("coder", ["ans", "range"])
@pytest.mark.parametrize("n", [0, 1, 148 * 1024 + 39, 148 * 1024 + 41, 3 * (148 * 1024 + 40) + 17])
def test_large_batch_decoders_short_streams(env, coder, n):
    """The one-CTA-per-SM decoders with (almost) nothing to decode: no symbols at all, fewer symbols than streams (empty
    streams and one-symbol streams side by side), a ragged last row.  Words, offsets and symbols equal the oracle's."""
    B, O, bc = env["B"], env["O"], env["bc"]
    k = 148 * 1024 + 40
    rng = np.random.default_rng(n + 5)
    syms = gauss_symbols(rng, n)
    model = B.ModelTable.quantized_gaussian(*BASE[:2], [BASE[2]], [BASE[3]])
    cdf = model.cdf()[0]
    enc_o = O.multi_ans_encode if coder == "ans" else O.multi_range_encode
    want_words, want_off = enc_o(syms, k, cdf, -50, threads=8)
    d_syms = dev(env, syms) if n else env["torch"].zeros(0, dtype=env["torch"].int32, device="cuda")
    comp = (bc.ans_encode if coder == "ans" else bc.range_encode)(d_syms, model, n_streams=k)
    words, off = comp.to_host()
    bc.check()
    assert np.array_equal(off, want_off)
    assert np.array_equal(words, want_words)
    out = (bc.ans_decode if coder == "ans" else bc.range_decode)(comp, model)
    bc.check()
    assert np.array_equal(out.cpu().numpy(), syms)


def test_interleaved_unaligned_symbol_buffer(env):
    """A symbol array that does not start on a 16-byte boundary cannot be a TMA tensor: per-row path, same words."""
    B, bc, torch = env["B"], env["bc"], env["torch"]
    rng = np.random.default_rng(77)
    n, k = 200_000, 256
    syms = gauss_symbols(rng, n + 1)
    model = B.ModelTable.quantized_gaussian(*BASE[:2], [BASE[2]], [BASE[3]])
    whole = dev(env, syms)
    aligned = whole[1:].clone()
    a = bc.ans_encode(aligned, model, n_streams=k)
    b = bc.ans_encode(whole[1:], model, n_streams=k)  # data pointer = base + 4 bytes
    bc.check()
    assert whole[1:].data_ptr() % 16 != 0 and aligned.data_ptr() % 16 == 0
    assert np.array_equal(a.to_host()[0], b.to_host()[0]) and np.array_equal(a.to_host()[1], b.to_host()[1])
    out = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    bc.ans_decode(a, model, out=out[1:])
    assert torch.equal(out[1:], whole[1:])


# ---------------------------------------------------------------------------------------------------
# contiguous (ragged) layout: empty streams, short streams, one long stream (config 4 / 5 shape)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("coder", ["ans", "range"])
def test_contiguous_ragged_matches_oracle(env, coder):
    B, O, bc, torch = env["B"], env["O"], env["bc"], env["torch"]
    rng = np.random.default_rng(17)
    lens = np.concatenate([[0, 1, 2, 31, 32, 33, 0, 5000, 64, 0], rng.integers(0, 300, size=150), [12345]])
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    n, k = int(off[-1]), lens.size
    syms = gauss_symbols(rng, n)
    syms[::89] = rng.integers(-50, 51, size=syms[::89].size)
    model = B.ModelTable.quantized_gaussian(*BASE[:2], [BASE[2]], [BASE[3]])
    cdf = model.cdf()[0]
    enc_o = O.multi_ans_encode if coder == "ans" else O.multi_range_encode
    want_words, want_off = enc_o(syms, k, cdf, -50, sym_offsets=off, threads=8)
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    comp = (bc.ans_encode if coder == "ans" else bc.range_encode)(dev(env, syms), model, sym_offsets=d_off)
    words, o = comp.to_host()
    bc.check()
    assert np.array_equal(o, want_off)
    assert np.array_equal(words, want_words)
    out = (bc.ans_decode if coder == "ans" else bc.range_decode)(comp, model)
    bc.check()
    assert np.array_equal(out.cpu().numpy(), syms)


# ---------------------------------------------------------------------------------------------------
# per-symbol / per-stream model index (config 3 and config 5 shapes, scaled down)
# ---------------------------------------------------------------------------------------------------
def _indexed_oracle_encode(O, coder, syms, idx, cdfs, lo):
    if coder == "ans":
        return O.ans_encode_indexed(syms, idx, cdfs, lo)
    enc = O.RangeEncoder()
    L = O.lib()
    for s, m in zip(syms, idx):
        l, p = C.c_uint32(), C.c_uint32()
        O._raise(L.orc_cdf_left_prob(O._p(cdfs[m], O.u32p), cdfs.shape[1] - 1, int(s) - lo, C.byref(l), C.byref(p)))
        O._raise(L.orc_renc_encode(C.byref(enc._e), l.value, p.value))
    return enc.get_compressed()


@pytest.mark.parametrize("coder", ["ans", "range"])
def test_per_symbol_categorical_models(env, coder):
    """config 3 shape: every symbol has its own 256-bin categorical model out of a pool."""
    B, O, bc, torch = env["B"], env["O"], env["bc"], env["torch"]
    rng = np.random.default_rng(3)
    M, A, n, k = 500, 256, 20_000, 40
    pmf = rng.dirichlet(0.5 * np.ones(A), size=M).astype(np.float32)
    model = B.ModelTable.categorical(pmf)
    cdfs = model.cdf()
    idx = rng.integers(0, M, size=n).astype(np.uint32)
    u = rng.integers(0, 1 << 24, size=n)
    syms = np.array([np.searchsorted(cdfs[m], q, side="right") - 1 for m, q in zip(idx, u)], dtype=np.int32)
    d_idx = dev(env, idx.view(np.int32))
    comp = (bc.ans_encode if coder == "ans" else bc.range_encode)(dev(env, syms), model, n_streams=k, model_index=d_idx,
                                                                  index_mode=1)
    words, off = comp.to_host()
    bc.check()
    for s in range(k):
        want = _indexed_oracle_encode(O, coder, syms[s::k], idx[s::k], cdfs, 0)
        assert np.array_equal(words[int(off[s]):int(off[s + 1])], want), s
    out = (bc.ans_decode if coder == "ans" else bc.range_decode)(comp, model, model_index=d_idx, index_mode=1)
    bc.check()
    assert np.array_equal(out.cpu().numpy(), syms)


@pytest.mark.parametrize("coder", ["ans", "range"])
def test_per_stream_gaussian_models(env, coder):
    """config 5 shape: one stream per (image, channel), one QuantizedGaussian per channel."""
    B, O, bc, torch = env["B"], env["O"], env["bc"], env["torch"]
    rng = np.random.default_rng(5)
    images, channels, hw = 3, 24, 100
    mu = rng.normal(0, 2, channels)
    sigma = np.exp(rng.uniform(np.log(0.3), np.log(12), channels))
    model = B.ModelTable.quantized_gaussian(-64, 64, mu, sigma)
    cdfs = model.cdf()
    lat = np.clip(np.rint(rng.normal(mu[None, :, None], sigma[None, :, None], size=(images, channels, hw))), -64, 64)
    syms = lat.astype(np.int32).reshape(-1)
    k = images * channels
    off = (np.arange(k + 1) * hw).astype(np.int64)
    sidx = np.tile(np.arange(channels), images).astype(np.int32)
    d_off, d_idx = torch.from_numpy(off).cuda(), dev(env, sidx)
    comp = (bc.ans_encode if coder == "ans" else bc.range_encode)(dev(env, syms), model, sym_offsets=d_off,
                                                                  model_index=d_idx, index_mode=2)
    words, o = comp.to_host()
    bc.check()
    for s in range(k):
        seg = syms[s * hw:(s + 1) * hw]
        want = (O.ans_encode_iid if coder == "ans" else O.range_encode_iid)(seg, cdfs[sidx[s]], -64)
        assert np.array_equal(words[int(o[s]):int(o[s + 1])], want), s
    out = (bc.ans_decode if coder == "ans" else bc.range_decode)(comp, model, model_index=d_idx, index_mode=2)
    bc.check()
    assert np.array_equal(out.cpu().numpy(), syms)


@pytest.mark.parametrize("coder", ["ans", "range"])
def test_contiguous_many_streams_big_ctas(env, coder):
    """>= 37,888 streams: the 256-thread CTA variants of the contiguous kernels."""
    B, O, bc, torch = env["B"], env["O"], env["bc"], env["torch"]
    rng = np.random.default_rng(23)
    lens = rng.integers(0, 20, size=40_000)
    lens[::1000] = rng.integers(100, 400, size=lens[::1000].size)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    n, k = int(off[-1]), lens.size
    syms = gauss_symbols(rng, n)
    model = B.ModelTable.quantized_gaussian(*BASE[:2], [BASE[2]], [BASE[3]])
    cdf = model.cdf()[0]
    enc_o = O.multi_ans_encode if coder == "ans" else O.multi_range_encode
    want_words, want_off = enc_o(syms, k, cdf, -50, sym_offsets=off, threads=8)
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    comp = (bc.ans_encode if coder == "ans" else bc.range_encode)(dev(env, syms), model, sym_offsets=d_off)
    words, o = comp.to_host()
    bc.check()
    assert np.array_equal(o, want_off)
    assert np.array_equal(words, want_words)
    out = (bc.ans_decode if coder == "ans" else bc.range_decode)(comp, model)
    bc.check()
    assert np.array_equal(out.cpu().numpy(), syms)


@pytest.mark.parametrize("coder", ["ans", "range"])
@pytest.mark.parametrize("M,A,contig", [(40, 256, False), (40, 256, True), (7, 1000, True), (500, 256, True)])
def test_model_sets_in_shared_memory_and_global(env, coder, M, A, contig):
    """Per-symbol model index: model sets that fit shared memory are decoded by the pool kernels
    (M = 40, or 7 models with a wide coarse index), bigger ones through L1/L2 (M = 500); both layouts."""
    B, O, bc, torch = env["B"], env["O"], env["bc"], env["torch"]
    rng = np.random.default_rng(M * 7 + A)
    n, k = 30_000, 50
    pmf = rng.dirichlet(0.3 * np.ones(A), size=M).astype(np.float32)
    model = B.ModelTable.categorical(pmf)
    cdfs = model.cdf()
    idx = rng.integers(0, M, size=n).astype(np.uint32)
    u = rng.integers(0, 1 << 24, size=n)
    syms = np.array([np.searchsorted(cdfs[m], q, side="right") - 1 for m, q in zip(idx, u)], dtype=np.int32)
    d_idx = dev(env, idx.view(np.int32))
    if contig:
        lens = rng.multinomial(n, np.ones(k) / k)
        lens[3] += lens[4]
        lens[4] = 0
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        layout = dict(sym_offsets=torch.from_numpy(off).cuda())
        pieces = [slice(int(off[s]), int(off[s + 1])) for s in range(k)]
    else:
        layout = dict(n_streams=k)
        pieces = [slice(s, n, k) for s in range(k)]
    comp = (bc.ans_encode if coder == "ans" else bc.range_encode)(dev(env, syms), model, model_index=d_idx, index_mode=1,
                                                                  **layout)
    words, o = comp.to_host()
    bc.check()
    for s in (0, 3, 4, 17, k - 1):
        want = _indexed_oracle_encode(O, coder, syms[pieces[s]], idx[pieces[s]], cdfs, 0)
        assert np.array_equal(words[int(o[s]):int(o[s + 1])], want), s
    out = (bc.ans_decode if coder == "ans" else bc.range_decode)(comp, model, model_index=d_idx, index_mode=1)
    bc.check()
    assert np.array_equal(out.cpu().numpy(), syms)


# ---------------------------------------------------------------------------------------------------
# data errors
# ---------------------------------------------------------------------------------------------------
def test_offsets_outside_the_symbol_array_are_rejected(env):
    """sym_offsets that reach beyond the symbol tensor (or decrease) must not be followed: ValueError, no access."""
    B, bc, torch = env["B"], env["bc"], env["torch"]
    model = B.ModelTable.quantized_gaussian(*BASE[:2], [BASE[2]], [BASE[3]])
    syms = dev(env, gauss_symbols(np.random.default_rng(1), 10_000))
    good = torch.arange(0, 11, device="cuda", dtype=torch.int64) * 1000
    comp = bc.ans_encode(syms, model, sym_offsets=good)
    bc.check()
    for bad in (good * 200, torch.flip(good, dims=[0])):
        for enc in (bc.ans_encode, bc.range_encode):
            enc(syms, model, sym_offsets=bad, checkpoint_every=64)
            with pytest.raises(ValueError):
                bc.check()
        comp.sym_offsets = bad
        bc.ans_decode(comp, model)
        with pytest.raises(ValueError):
            bc.check()
    comp.sym_offsets = good
    assert torch.equal(bc.ans_decode(comp, model), syms)


def test_data_errors(env):
    B, bc, torch = env["B"], env["bc"], env["torch"]
    model = B.ModelTable.quantized_gaussian(*BASE[:2], [BASE[2]], [BASE[3]])
    syms = np.zeros(1000, dtype=np.int32)
    syms[123] = 51
    bc.ans_encode(dev(env, syms), model, n_streams=10)
    with pytest.raises(KeyError):
        bc.check()
    bc.range_encode(dev(env, syms), model, n_streams=10)
    with pytest.raises(KeyError):
        bc.check()
    # ANS stream ending in a zero word
    words = torch.tensor([5, 0], dtype=torch.int32, device="cuda")
    offs = torch.tensor([0, 2], dtype=torch.int64, device="cuda")
    bc.ans_decode(B.Compressed(words, offs, 1, 4, "ans"), model)
    with pytest.raises(ValueError):
        bc.check()
    # range decoder: a point outside the coded range is invalid data only if quantile >= 2^24; craft it
    # with a model-independent trick: range = u64::MAX means scale = 2^40 - 1, point = u64::MAX -> q = 2^24
    words = torch.tensor([-1, -1], dtype=torch.int32, device="cuda")
    bc.range_decode(B.Compressed(words, offs, 1, 1, "range"), model)
    with pytest.raises(AssertionError):
        bc.check()


# ---------------------------------------------------------------------------------------------------
# host-buffer C ABI (the reference-facing call)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("coder", ["ans", "range"])
def test_host_buffer_entry_points(env, coder):
    from constriction_b200 import _native as N
    B, O = env["B"], env["O"]
    lib = N.load()
    rng = np.random.default_rng(23)
    n, k = 300_000, 2048
    syms = gauss_symbols(rng, n)
    model = B.ModelTable.quantized_gaussian(*BASE[:2], [BASE[2]], [BASE[3]])
    cdf = model.cdf()[0]
    want_words, want_off = (O.multi_ans_encode if coder == "ans" else O.multi_range_encode)(syms, k, cdf, -50, threads=8)
    L = N.Layout()
    L.n_streams, L.n_symbols = k, n
    cap = lib.ctr_ans_max_compressed_words(C.byref(L))
    words_buf = np.empty(cap, dtype=np.uint32)
    off = np.empty(k + 1, dtype=np.uint64)
    status, bad = C.c_int(), C.c_uint64()
    enc = lib.ctr_ans_encode_reverse_host if coder == "ans" else lib.ctr_range_encode_host
    rc = enc(model.handle, syms.ctypes.data, n, k, None, None, 0, words_buf.ctypes.data, cap, off.ctypes.data,
             C.byref(status), C.byref(bad))
    assert rc == 0 and status.value == 0
    words = words_buf[: int(off[-1])].copy()
    assert np.array_equal(off, want_off) and np.array_equal(words, want_words)
    out = np.empty(n, dtype=np.int32)
    dec = lib.ctr_ans_decode_host if coder == "ans" else lib.ctr_range_decode_host
    rc = dec(model.handle, words.ctypes.data, off.ctypes.data, n, k, None, None, 0, out.ctypes.data, C.byref(status),
             C.byref(bad))
    assert rc == 0 and status.value == 0
    assert np.array_equal(out, syms)


# ---------------------------------------------------------------------------------------------------
# BASELINE size (1e8 symbols): size-independent properties + spot checks against the oracle
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("coder", ["ans", "range"])
def test_full_size_round_trip(env, coder):
    B, O, bc, torch = env["B"], env["O"], env["bc"], env["torch"]
    n, k = 100_000_000, 148 * 1024
    g = torch.Generator(device="cuda")
    g.manual_seed(2)
    syms = torch.clamp(torch.round(torch.randn(n, device="cuda", generator=g) * 9.6 + 3.2), -50, 50).to(torch.int32)
    model = B.ModelTable.quantized_gaussian(*BASE[:2], [BASE[2]], [BASE[3]])
    comp = (bc.ans_encode if coder == "ans" else bc.range_encode)(syms, model, n_streams=k)
    out = (bc.ans_decode if coder == "ans" else bc.range_decode)(comp, model)
    bc.check()
    assert torch.equal(out, syms)                      # encode -> decode is the identity
    total = comp.total_words()
    bits_per_symbol = 32.0 * total / n
    assert 5.31 < bits_per_symbol < 5.31 * 1.05        # entropy of the model is 5.3108 bits
    off = comp.offsets.cpu().numpy()
    assert np.all(np.diff(off) >= 0) and off[0] == 0
    # spot-check streams against the oracle, word for word
    cdf = model.cdf()[0]
    for s in [0, 1, 31, 32, k // 2 + 7, k - 1]:
        seg = syms[s::k].cpu().numpy()
        want = (O.ans_encode_iid if coder == "ans" else O.range_encode_iid)(seg, cdf, -50)
        assert np.array_equal(comp.stream_words(s), want), s


def test_north_star_size_round_trip(env):
    """BASELINE.json's target size: 1e9 i.i.d. symbols on one GPU (ANS).  Size-independent properties (decode(encode(x))
    == x, monotone offsets, bits per symbol next to the model's entropy) plus streams cut out of the container and
    compared with the oracle word for word."""
    B, O, bc, torch = env["B"], env["O"], env["bc"], env["torch"]
    n, k = 1_000_000_000, 148 * 1024
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * (1 << 30):
        pytest.skip("needs ~30 GB of device memory")
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    syms = torch.empty(n, dtype=torch.int32, device="cuda")
    for lo in range(0, n, 250_000_000):  # in pieces: the float temporaries of 1e9 elements would be 8 GB each
        piece = torch.randn(250_000_000, device="cuda", generator=g)
        syms[lo:lo + 250_000_000] = torch.clamp(torch.round(piece * 9.6 + 3.2), -50, 50).to(torch.int32)
        del piece
    model = B.ModelTable.quantized_gaussian(*BASE[:2], [BASE[2]], [BASE[3]])
    comp = bc.ans_encode(syms, model, n_streams=k)
    out = bc.ans_decode(comp, model)
    bc.check()
    assert torch.equal(out, syms)
    del out
    bits_per_symbol = 32.0 * comp.total_words() / n
    assert 5.31 < bits_per_symbol < 5.31 * 1.01
    off = comp.offsets.cpu().numpy()
    assert np.all(np.diff(off) >= 0) and off[0] == 0
    cdf = model.cdf()[0]
    for s in [0, 31, 4097, k // 2 + 7, k - 1]:
        want = O.ans_encode_iid(syms[s::k].cpu().numpy(), cdf, -50)
        assert np.array_equal(comp.stream_words(s), want), s

