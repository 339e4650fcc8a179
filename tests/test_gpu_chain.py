"""The few-long-streams ("chain") kernels: contiguous layout, one coder warp per 32 streams fed by producer warps.
Every stream must equal the oracle's coder word for word (ragged lengths, shared and per-stream models, checkpoints)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BASE = (-50, 50, 3.2, 9.6)


@pytest.fixture(scope="module")
def env(oracle):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from constriction_b200 import batch as B
    return dict(torch=torch, B=B, bc=B.BatchCoder(), O=oracle)


def make_batch(rng, k, max_len, means, stds, idx):
    lens = rng.integers(max_len // 2, max_len, size=k)
    if k > 3:
        lens[[1, k - 2]] = [0, 1]
    lens[0] = max_len + 17  # keeps the batch in the long-stream regime and exercises a lone long tail
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    syms = np.empty(int(off[-1]), dtype=np.int32)
    for s in range(k):
        a, b = int(off[s]), int(off[s + 1])
        syms[a:b] = np.clip(np.rint(rng.normal(means[idx[s]], stds[idx[s]], size=b - a)), -50, 50)
    return lens, off, syms


@pytest.mark.parametrize("coder", ["ans", "range"])
@pytest.mark.parametrize("k,max_len", [(1, 5000), (33, 700), (1000, 300), (64, 4096)])
def test_chain_shared_model_matches_the_oracle(env, coder, k, max_len):
    B, O, bc, torch = env["B"], env["O"], env["bc"], env["torch"]
    rng = np.random.default_rng(k * 7 + max_len)
    lens, off, syms = make_batch(rng, k, max_len, [BASE[2]], [BASE[3]], np.zeros(k, dtype=np.int64))
    assert syms.size >= 64 * k
    model = B.ModelTable.quantized_gaussian(*BASE[:2], [BASE[2]], [BASE[3]])
    cdf = model.cdf()[0]
    enc = bc.ans_encode if coder == "ans" else bc.range_encode
    dec = bc.ans_decode if coder == "ans" else bc.range_decode
    comp = enc(torch.from_numpy(syms).cuda(), model, sym_offsets=torch.from_numpy(off).cuda())
    out = dec(comp, model)
    bc.check()
    assert np.array_equal(out.cpu().numpy(), syms)
    words, o = comp.to_host()
    want_words, want_off = (O.multi_ans_encode if coder == "ans" else O.multi_range_encode)(syms, k, cdf, -50, sym_offsets=off.astype(np.uint64), threads=8)
    assert np.array_equal(o, want_off) and np.array_equal(words, want_words)


@pytest.mark.parametrize("coder", ["ans", "range"])
def test_chain_per_stream_models_and_checkpoints(env, coder):
    B, O, bc, torch = env["B"], env["O"], env["bc"], env["torch"]
    rng = np.random.default_rng(99)
    k = 200
    means, stds = np.array([3.2, -10.0, 20.0, 0.0]), np.array([9.6, 2.0, 30.0, 0.4])
    idx = rng.integers(0, 4, size=k)
    lens, off, syms = make_batch(rng, k, 1500, means, stds, idx)
    model = B.ModelTable.quantized_gaussian(-50, 50, means, stds)
    cdfs = model.cdf()
    enc = bc.ans_encode if coder == "ans" else bc.range_encode
    dec = bc.ans_decode if coder == "ans" else bc.range_decode
    d_syms, d_off = torch.from_numpy(syms).cuda(), torch.from_numpy(off).cuda()
    d_idx = torch.from_numpy(idx.astype(np.int32)).cuda()
    plain = enc(d_syms, model, sym_offsets=d_off, model_index=d_idx, index_mode=2)
    ck = enc(d_syms, model, sym_offsets=d_off, model_index=d_idx, index_mode=2, checkpoint_every=128)
    bc.check()
    enc1 = O.ans_encode_iid if coder == "ans" else O.range_encode_iid
    for s in list(range(0, k, 17)) + [0, 1, k - 2, k - 1]:
        want = enc1(syms[off[s]:off[s + 1]], cdfs[idx[s]], -50)
        assert np.array_equal(plain.stream_words(s), want), s
    w0, o0 = plain.to_host()
    w1, o1 = ck.to_host()
    assert np.array_equal(o0, o1) and np.array_equal(w0, w1)          # records never change the words
    for use in (True, False):                                        # chunk-parallel and plain decode of the same container
        out = dec(ck, model, model_index=d_idx, index_mode=2, use_checkpoints=use)
        bc.check()
        assert np.array_equal(out.cpu().numpy(), syms)


def test_chain_impossible_symbol(env):
    B, bc, torch = env["B"], env["bc"], env["torch"]
    rng = np.random.default_rng(5)
    lens, off, syms = make_batch(rng, 40, 600, [BASE[2]], [BASE[3]], np.zeros(40, dtype=np.int64))
    syms[int(off[7]) + 100] = 99
    model = B.ModelTable.quantized_gaussian(*BASE[:2], [BASE[2]], [BASE[3]])
    for enc in (bc.ans_encode, bc.range_encode):
        enc(torch.from_numpy(syms).cuda(), model, sym_offsets=torch.from_numpy(off).cuda())
        with pytest.raises(KeyError, match="stream 7"):
            bc.check()
