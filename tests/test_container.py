"""Wire container (ctr_container_*; host code, no device): the bytes round-trip, a stream cut from them is a stock
constriction stream (decoded here by the oracle's AnsCoder / RangeDecoder), and the records are the coder positions
(Pos::pos) a stock coder can `seek` to."""
import ctypes as C

import numpy as np
import pytest

LO, HI, MEAN, STD = -50, 50, 3.2, 9.6
EVERY = 64


def build_view(oracle, coder, with_records):
    from constriction_b200 import _native as N
    rng = np.random.default_rng(7)
    lens = [300, 0, 64, 1, 129, 1000]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    syms = np.clip(np.rint(rng.normal(MEAN, STD, size=int(off[-1]))), LO, HI).astype(np.int32)
    model = oracle.QuantizedGaussian(LO, HI, MEAN, STD)
    streams, records, ck_off = [], [], [0]
    for k, n in enumerate(lens):
        s = syms[int(off[k]):int(off[k + 1])]
        J = -(-n // EVERY)
        if coder == "ans":
            starts = [0] + [n - (J - j) * EVERY for j in range(1, J)]
            enc, rec = oracle.AnsCoder(), {}
            for j in range(J - 1, -1, -1):
                hi = starts[j + 1] if j + 1 < J else n
                enc.encode_reverse(s[starts[j]:hi], model)
                pos, state = enc.pos()
                rec[j] = [pos, state]
            records += [rec[j] for j in range(J)]
        else:
            enc = oracle.RangeEncoder()
            for j in range(J):
                pos, (lower, rng_) = enc.pos()
                records.append([pos, lower, rng_, 0])
                enc.encode(s[j * EVERY:(j + 1) * EVERY], model)
        streams.append(enc.get_compressed())
        ck_off.append(ck_off[-1] + J)
    words = np.concatenate(streams + [np.empty(0, np.uint32)]).astype(np.uint32)
    woff = np.concatenate([[0], np.cumsum([w.size for w in streams])]).astype(np.uint64)
    rec = np.array(records, dtype=np.uint64).reshape(-1)
    ck = np.array(ck_off, dtype=np.uint64)
    v = N.ContainerView()
    v.coder = 0 if coder == "ans" else 1
    v.word_bits, v.precision = 32, 24
    v.n_streams, v.n_symbols, v.total_words = len(lens), syms.size, words.size
    v.sym_offsets, v.offsets, v.words = off.ctypes.data, woff.ctypes.data, words.ctypes.data
    if with_records:
        v.checkpoint_every, v.n_records = EVERY, int(ck[-1])
        v.ckpt_offsets, v.records = ck.ctypes.data, rec.ctypes.data
    return v, dict(lens=lens, off=off, syms=syms, words=words, woff=woff, rec=rec, ck=ck, model=model)


@pytest.mark.parametrize("coder", ["ans", "range"])
@pytest.mark.parametrize("with_records", [False, True])
def test_pack_unpack_and_stock_decode(oracle, coder, with_records):
    from constriction_b200 import _native as N
    from constriction_b200 import container as Cn
    lib = N.load()
    v, d = build_view(oracle, coder, with_records)
    size = lib.ctr_container_size(C.byref(v))
    buf = np.zeros(size // 8, dtype=np.uint64)
    assert size % 8 == 0 and lib.ctr_container_pack(C.byref(v), buf.ctypes.data, size) == 0
    assert lib.ctr_container_pack(C.byref(v), buf.ctypes.data, size - 8) == N.ERR_OUT_OF_SPACE
    data = buf.tobytes()
    assert data[:8] == b"CTRB200\0"
    h = Cn.unpack_host(data)
    assert h["coder"] == coder and h["n_streams"] == len(d["lens"]) and h["n_symbols"] == d["syms"].size
    assert np.array_equal(h["offsets"], d["woff"]) and np.array_equal(h["words"], d["words"]) and np.array_equal(h["sym_offsets"], d["off"])
    # every stream cut from the bytes is a stock stream
    for k, n in enumerate(d["lens"]):
        w = h["words"][int(h["offsets"][k]):int(h["offsets"][k + 1])]
        s = d["syms"][int(d["off"][k]):int(d["off"][k + 1])]
        dec = oracle.AnsCoder(w) if coder == "ans" else oracle.RangeDecoder(w)
        assert np.array_equal(dec.decode(d["model"], n) if n else np.empty(0, np.int32), s)
    if with_records:
        assert h["checkpoint_every"] == EVERY and np.array_equal(h["ckpt_offsets"], d["ck"]) and np.array_equal(h["records"], d["rec"])
        k = 5  # seek a stock coder to the records of the longest stream and decode single chunks
        w = h["words"][int(h["offsets"][k]):int(h["offsets"][k + 1])]
        s = d["syms"][int(d["off"][k]):int(d["off"][k + 1])]
        n, J, c0 = d["lens"][k], -(-d["lens"][k] // EVERY), int(h["ckpt_offsets"][k])
        for j in (0, 3, J - 1):
            if coder == "ans":
                r = h["records"].reshape(-1, 2)[c0 + j]
                starts = [0] + [n - (J - jj) * EVERY for jj in range(1, J)]
                dec = oracle.AnsCoder(w)
                dec.seek(int(r[0]), int(r[1]))
                hi = starts[j + 1] if j + 1 < J else n
                assert np.array_equal(dec.decode(d["model"], hi - starts[j]), s[starts[j]:hi])
            else:
                r = h["records"].reshape(-1, 4)[c0 + j]
                dec = oracle.RangeDecoder(w)
                dec.seek(int(r[0]), (int(r[1]), int(r[2])))
                assert np.array_equal(dec.decode(d["model"], min(EVERY, n - j * EVERY)), s[j * EVERY:(j + 1) * EVERY])


def test_unpack_rejects_damaged_containers(oracle):
    from constriction_b200 import _native as N
    lib = N.load()
    v, _ = build_view(oracle, "ans", True)
    size = lib.ctr_container_size(C.byref(v))
    buf = np.zeros(size // 8, dtype=np.uint64)
    assert lib.ctr_container_pack(C.byref(v), buf.ctypes.data, size) == 0
    out = N.ContainerView()
    assert lib.ctr_container_unpack(buf.ctypes.data, size, C.byref(out)) == 0
    assert lib.ctr_container_unpack(buf.ctypes.data, size - 16, C.byref(out)) == N.ERR_INVALID_DATA      # truncated
    bad = buf.copy()
    bad.view(np.uint8)[0] = ord("X")
    assert lib.ctr_container_unpack(bad.ctypes.data, size, C.byref(out)) == N.ERR_INVALID_DATA           # magic
    bad = buf.copy()
    k1 = int(v.n_streams) + 1
    bad[8 + k1 + 2] = 10**9                                                                               # offsets not monotone
    assert lib.ctr_container_unpack(bad.ctypes.data, size, C.byref(out)) == N.ERR_INVALID_DATA
    assert lib.ctr_container_unpack(buf.ctypes.data + 4, size, C.byref(out)) == N.ERR_BAD_ARGUMENT       # misaligned
