"""Checkpoints (CTR_FLAG_CHECKPOINTS): the encoders record `pos()` every C symbols, the decoders
decode every chunk of every stream on its own lane.  The records must equal what the reference's
`AnsCoder.pos()` / `RangeEncoder.pos()` return at those moments (stack.rs:1107-1115, queue.rs:182-196), the words
must be unchanged, a stock coder must be able to `seek` to any record (stack.rs:1117-1139, queue.rs:911-928) and the
chunk-parallel decode must return the symbols."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LO, HI, MEAN, STD = -50, 50, 3.2, 9.6


@pytest.fixture(scope="module")
def env(oracle):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from constriction_b200 import batch as B
    return dict(torch=torch, B=B, bc=B.BatchCoder(), O=oracle)


def make_batch(env, seed, lengths):
    rng = np.random.default_rng(seed)
    off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    syms = np.clip(np.rint(rng.normal(MEAN, STD, size=int(off[-1]))), LO, HI).astype(np.int32)
    t = env["torch"]
    return syms, off, t.from_numpy(syms).cuda(), t.from_numpy(off).cuda()


LENGTHS = [1000, 0, 64, 65, 1, 31, 32, 33, 4096, 777, 128, 2048 + 17]


@pytest.mark.parametrize("every", [32, 64, 256])
def test_ans_checkpoints_equal_reference_pos(env, every):
    B, bc, O = env["B"], env["bc"], env["O"]
    syms, off, d_syms, d_off = make_batch(env, 1, LENGTHS)
    model = B.ModelTable.quantized_gaussian(LO, HI, [MEAN], [STD])
    omodel = O.QuantizedGaussian(LO, HI, MEAN, STD)
    plain = bc.ans_encode(d_syms, model, sym_offsets=d_off)
    comp = bc.ans_encode(d_syms, model, sym_offsets=d_off, checkpoint_every=every)
    bc.check()
    w0, o0 = plain.to_host()
    w1, o1 = comp.to_host()
    assert np.array_equal(w0, w1) and np.array_equal(o0, o1), "checkpoints must not change the words"
    ck_off = comp.ckpt_offsets.cpu().numpy()
    rec = comp.checkpoints.cpu().numpy().view(np.uint64).reshape(-1, 2)
    for k, n in enumerate(LENGTHS):
        J = -(-n // every)
        assert ck_off[k + 1] - ck_off[k] == J
        if n == 0:
            continue
        s = syms[off[k]:off[k + 1]]
        starts = [0] + [n - (J - j) * every for j in range(1, J)]
        # the reference coder, fed the chunks last to first, reports these positions
        coder = O.AnsCoder()
        want = {}
        for j in range(J - 1, -1, -1):
            hi = starts[j + 1] if j + 1 < J else n
            coder.encode_reverse(s[starts[j]:hi], omodel)
            want[j] = coder.pos()
        for j in range(J):
            got = (int(rec[ck_off[k] + j, 0]), int(rec[ck_off[k] + j, 1]))
            assert got == want[j], (k, j)
        # a stock decoder can seek to any record and decode that chunk
        words = comp.stream_words(k)
        for j in (0, J // 2, J - 1):
            dec = O.AnsCoder(words)
            dec.seek(*want[j])
            hi = starts[j + 1] if j + 1 < J else n
            assert np.array_equal(dec.decode(omodel, hi - starts[j]), s[starts[j]:hi])
    out = bc.ans_decode(comp, model)
    bc.check()
    assert np.array_equal(out.cpu().numpy(), syms)
    out2 = bc.ans_decode(comp, model, use_checkpoints=False)
    assert np.array_equal(out2.cpu().numpy(), syms)


@pytest.mark.parametrize("every", [32, 128])
def test_range_checkpoints_equal_reference_pos(env, every):
    B, bc, O = env["B"], env["bc"], env["O"]
    syms, off, d_syms, d_off = make_batch(env, 2, LENGTHS)
    model = B.ModelTable.quantized_gaussian(LO, HI, [MEAN], [STD])
    omodel = O.QuantizedGaussian(LO, HI, MEAN, STD)
    plain = bc.range_encode(d_syms, model, sym_offsets=d_off)
    comp = bc.range_encode(d_syms, model, sym_offsets=d_off, checkpoint_every=every)
    bc.check()
    w0, o0 = plain.to_host()
    w1, o1 = comp.to_host()
    assert np.array_equal(w0, w1) and np.array_equal(o0, o1)
    ck_off = comp.ckpt_offsets.cpu().numpy()
    rec = comp.checkpoints.cpu().numpy().view(np.uint64).reshape(-1, 4)
    for k, n in enumerate(LENGTHS):
        J = -(-n // every)
        assert ck_off[k + 1] - ck_off[k] == J
        if n == 0:
            continue
        s = syms[off[k]:off[k + 1]]
        enc = O.RangeEncoder()
        for j in range(J):
            pos, (lower, rng) = enc.pos()
            got = rec[ck_off[k] + j]
            assert (int(got[0]), int(got[1]), int(got[2])) == (pos, lower, rng), (k, j)
            enc.encode(s[j * every:(j + 1) * every], omodel)
        words = comp.stream_words(k)
        assert np.array_equal(words, enc.get_compressed())
        for j in (0, J // 2, J - 1):
            got = rec[ck_off[k] + j]
            dec = O.RangeDecoder(words)
            dec.seek(int(got[0]), (int(got[1]), int(got[2])))
            chunk = s[j * every:(j + 1) * every]
            assert np.array_equal(dec.decode(omodel, chunk.size), chunk)
    out = bc.range_decode(comp, model)
    bc.check()
    assert np.array_equal(out.cpu().numpy(), syms)


@pytest.mark.parametrize("coder", ["ans", "range"])
def test_chunked_decode_models_and_big_batch(env, coder):
    """Per-stream and per-symbol model indices, per-symbol Gaussian parameters, and a long-stream batch."""
    B, bc, torch = env["B"], env["bc"], env["torch"]
    enc = bc.ans_encode if coder == "ans" else bc.range_encode
    dec = bc.ans_decode if coder == "ans" else bc.range_decode
    rng = np.random.default_rng(7)
    lengths = list(rng.integers(0, 5000, size=40))
    syms, off, d_syms, d_off = make_batch(env, 3, lengths)
    n, k = syms.size, len(lengths)
    means, stds = rng.normal(0, 10, 16), np.exp(rng.uniform(0, 3, 16))
    pool = B.ModelTable.quantized_gaussian(LO, HI, means, stds)
    per_stream = torch.from_numpy(rng.integers(0, 16, size=k).astype(np.int32)).cuda()
    per_symbol = torch.from_numpy(rng.integers(0, 16, size=n).astype(np.int32)).cuda()
    for idx, mode in ((per_stream, 2), (per_symbol, 1)):
        comp = enc(d_syms, pool, sym_offsets=d_off, model_index=idx, index_mode=mode, checkpoint_every=96)
        out = dec(comp, pool, model_index=idx, index_mode=mode)
        bc.check()
        assert np.array_equal(out.cpu().numpy(), syms)
    lazy = B.GaussianParams(LO, HI, rng.normal(0, 10, n), np.exp(rng.uniform(-2, 3, n)))
    comp = enc(d_syms, lazy, sym_offsets=d_off, checkpoint_every=64)
    ref = enc(d_syms, lazy, sym_offsets=d_off)
    assert np.array_equal(comp.to_host()[0], ref.to_host()[0])
    out = dec(comp, lazy)
    bc.check()
    assert np.array_equal(out.cpu().numpy(), syms)
    # BASELINE config 4's shard shape: 1024 streams x 122,070 symbols
    g = torch.Generator(device="cuda")
    g.manual_seed(4)
    kk, per = 1024, 122_070
    big = torch.clamp(torch.round(torch.randn(kk * per, device="cuda", generator=g) * STD + MEAN), LO, HI).to(torch.int32)
    boff = torch.arange(kk + 1, device="cuda", dtype=torch.int64) * per
    model = B.ModelTable.quantized_gaussian(LO, HI, [MEAN], [STD])
    comp = enc(big, model, sym_offsets=boff, checkpoint_every=1024)
    plain = enc(big, model, sym_offsets=boff)
    assert torch.equal(comp.offsets, plain.offsets) and torch.equal(comp.words[:comp.total_words()], plain.words[:plain.total_words()])
    out = dec(comp, model)
    bc.check()
    assert torch.equal(out, big)


@pytest.mark.parametrize("coder", ["ans", "range"])
def test_wire_container_round_trip_through_the_device(env, coder):
    """encode on the GPU with records -> bytes (ctr_container_pack) -> device again -> chunk-parallel decode; a stream
    cut from the bytes decodes with the oracle's stock coder."""
    from constriction_b200 import container as Cn
    B, bc, O = env["B"], env["bc"], env["O"]
    syms, off, d_syms, d_off = make_batch(env, 9, LENGTHS)
    model = B.ModelTable.quantized_gaussian(LO, HI, [MEAN], [STD])
    enc, dec = (bc.ans_encode, bc.ans_decode) if coder == "ans" else (bc.range_encode, bc.range_decode)
    comp = enc(d_syms, model, sym_offsets=d_off, checkpoint_every=64)
    bc.check()
    data = Cn.pack(comp)
    back = Cn.unpack(data)
    assert back.coder == coder and back.checkpoint_every == 64
    for use in (True, False):
        out = dec(back, model, use_checkpoints=use)
        bc.check()
        assert np.array_equal(out.cpu().numpy(), syms)
    h = Cn.unpack_host(data)
    omodel = O.QuantizedGaussian(LO, HI, MEAN, STD)
    k = 8
    w = h["words"][int(h["offsets"][k]):int(h["offsets"][k + 1])]
    stock = O.AnsCoder(w) if coder == "ans" else O.RangeDecoder(w)
    assert np.array_equal(stock.decode(omodel, LENGTHS[k]), syms[off[k]:off[k + 1]])
    plain = Cn.unpack(Cn.pack(enc(d_syms, model, n_streams=7)))  # interleaved deal, no records
    assert np.array_equal(dec(plain, model).cpu().numpy(), syms)
