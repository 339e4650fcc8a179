"""The host-buffer C ABI (ctr_*_host / ctr_*_host_async): batches big enough to be cut into several pipelined
chunks must give, word for word, what the device-pointer path and the oracle give."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BASE = (-50, 50, 3.2, 9.6)


@pytest.fixture(scope="module")
def env(oracle):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from constriction_b200 import _native as N
    from constriction_b200 import batch as B
    return dict(torch=torch, B=B, N=N, lib=N.load(), O=oracle, bc=B.BatchCoder())


def gauss_symbols(rng, n):
    return np.clip(np.rint(rng.normal(BASE[2], BASE[3], size=n)), BASE[0], BASE[1]).astype(np.int32)


def host_encode(env, coder, model, syms, k, sym_off=None, idx=None, mode=0, use_async=False, pinned=False):
    N, lib, torch = env["N"], env["lib"], env["torch"]
    L = N.Layout()
    L.n_streams, L.n_symbols = k, syms.size
    cap = int(lib.ctr_ans_max_compressed_words(C.byref(L)))
    if pinned:
        words_t = torch.empty(cap, dtype=torch.int32).pin_memory()
        off_t = torch.empty(k + 1, dtype=torch.int64).pin_memory()
        words_buf, off = words_t.numpy().view(np.uint32), off_t.numpy().view(np.uint64)
    else:
        words_buf, off = np.empty(cap, dtype=np.uint32), np.empty(k + 1, dtype=np.uint64)
    status, bad = C.c_int(), C.c_uint64()
    args = (model.handle, syms.ctypes.data, syms.size, k, None if sym_off is None else sym_off.ctypes.data,
            None if idx is None else idx.ctypes.data, mode, words_buf.ctypes.data, cap, off.ctypes.data, C.byref(status), C.byref(bad))
    if use_async:
        job = C.c_void_p()
        fn = lib.ctr_ans_encode_reverse_host_async if coder == "ans" else lib.ctr_range_encode_host_async
        assert fn(*args, C.byref(job)) == 0
        rc = lib.ctr_host_job_wait(job)
    else:
        rc = (lib.ctr_ans_encode_reverse_host if coder == "ans" else lib.ctr_range_encode_host)(*args)
    return rc, status.value, bad.value, words_buf, off


def host_decode(env, coder, model, words, off, n, k, sym_off=None, idx=None, mode=0, use_async=False):
    lib = env["lib"]
    out = np.full(n, -777, dtype=np.int32)
    status, bad = C.c_int(), C.c_uint64()
    args = (model.handle, words.ctypes.data, off.ctypes.data, n, k, None if sym_off is None else sym_off.ctypes.data,
            None if idx is None else idx.ctypes.data, mode, out.ctypes.data, C.byref(status), C.byref(bad))
    if use_async:
        job = C.c_void_p()
        fn = lib.ctr_ans_decode_host_async if coder == "ans" else lib.ctr_range_decode_host_async
        assert fn(*args, C.byref(job)) == 0
        rc = lib.ctr_host_job_wait(job)
    else:
        rc = (lib.ctr_ans_decode_host if coder == "ans" else lib.ctr_range_decode_host)(*args)
    return rc, status.value, out


@pytest.mark.parametrize("coder", ["ans", "range"])
@pytest.mark.parametrize("n,k", [(9_000_000 + 13, 4096 + 40), (8_388_608, 8192), (5_000_001, 2048), (70_000, 100_000)])
def test_interleaved_chunks_match_the_oracle(env, coder, n, k):
    B, O = env["B"], env["O"]
    rng = np.random.default_rng(n % 1000 + k)
    syms = gauss_symbols(rng, n)
    model = B.ModelTable.quantized_gaussian(*BASE[:2], [BASE[2]], [BASE[3]])
    cdf = model.cdf()[0]
    want_words, want_off = (O.multi_ans_encode if coder == "ans" else O.multi_range_encode)(syms, k, cdf, -50, threads=16)
    for use_async in (False, True):
        rc, st, _, words_buf, off = host_encode(env, coder, model, syms, k, use_async=use_async, pinned=use_async)
        assert rc == 0 and st == 0
        assert np.array_equal(off, want_off)
        words = words_buf[: int(off[-1])].copy()
        assert np.array_equal(words, want_words)
        rc, st, out = host_decode(env, coder, model, words, off, n, k, use_async=use_async)
        assert rc == 0 and st == 0 and np.array_equal(out, syms)


@pytest.mark.parametrize("coder", ["ans", "range"])
def test_contiguous_ragged_chunks_with_per_stream_models(env, coder):
    B, O = env["B"], env["O"]
    rng = np.random.default_rng(77)
    k = 3000
    lens = rng.integers(0, 6000, size=k)
    lens[[0, 17, k - 1]] = 0
    sym_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    n = int(sym_off[-1])
    means, stds = np.array([3.2, -10.0, 20.0]), np.array([9.6, 2.0, 30.0])
    model = B.ModelTable.quantized_gaussian(-50, 50, means, stds)
    cdfs = model.cdf()
    idx = rng.integers(0, 3, size=k).astype(np.uint32)
    syms = np.empty(n, dtype=np.int32)
    for s in range(k):
        a, b = int(sym_off[s]), int(sym_off[s + 1])
        syms[a:b] = np.clip(np.rint(rng.normal(means[idx[s]], stds[idx[s]], size=b - a)), -50, 50)
    rc, st, _, words_buf, off = host_encode(env, coder, model, syms, k, sym_off=sym_off, idx=idx, mode=2)
    assert rc == 0 and st == 0
    enc1 = O.ans_encode_iid if coder == "ans" else O.range_encode_iid
    for s in list(range(0, k, 97)) + [k - 1]:
        a, b = int(sym_off[s]), int(sym_off[s + 1])
        assert np.array_equal(words_buf[int(off[s]):int(off[s + 1])], enc1(syms[a:b], cdfs[idx[s]], -50)), s
    words = words_buf[: int(off[-1])].copy()
    rc, st, out = host_decode(env, coder, model, words, off, n, k, sym_off=sym_off, idx=idx, mode=2, use_async=True)
    assert rc == 0 and st == 0 and np.array_equal(out, syms)


def test_host_errors(env):
    B = env["B"]
    rng = np.random.default_rng(3)
    n, k = 6_000_000, 4096
    syms = gauss_symbols(rng, n)
    bad_at = 5_000_000 + 77  # a stream in a late chunk
    syms[bad_at] = 1000
    model = B.ModelTable.quantized_gaussian(*BASE[:2], [BASE[2]], [BASE[3]])
    rc, st, bad, _, _ = host_encode(env, "ans", model, syms, k)
    assert rc == 0 and st == env["N"].ERR_IMPOSSIBLE_SYMBOL and bad == bad_at % k
    # capacity too small
    syms[bad_at] = 0
    N, lib = env["N"], env["lib"]
    words = np.empty(1000, dtype=np.uint32)
    off = np.empty(k + 1, dtype=np.uint64)
    status, b = C.c_int(), C.c_uint64()
    rc = lib.ctr_ans_encode_reverse_host(model.handle, syms.ctypes.data, n, k, None, None, 0, words.ctypes.data, 1000,
                                         off.ctypes.data, C.byref(status), C.byref(b))
    assert rc == N.ERR_OUT_OF_SPACE
    # offsets that leave the symbol array
    so = np.array([0, 10, 5], dtype=np.uint64)
    rc = lib.ctr_ans_encode_reverse_host(model.handle, syms.ctypes.data, 10, 2, so.ctypes.data, None, 0, words.ctypes.data, 1000,
                                         off.ctypes.data, C.byref(status), C.byref(b))
    assert rc == N.ERR_BAD_ARGUMENT
