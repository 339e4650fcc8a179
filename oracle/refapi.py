"""CPU oracle with the reference's Python API shape (TEST INFRASTRUCTURE ONLY).

Mirrors `constriction.stream.{stack,queue,model}` (reference:
src/pybindings/stream/stack.rs, queue.rs, model.rs, model/internals.rs) on top
of `liboracle.so` (oracle.c), evaluating the entropy model lazily *per symbol*
exactly like the reference does.  The golden vectors of the reference's own
pytest files are replayed against these classes (tests/test_oracle_golden.py),
which is what pins the oracle.

Nothing under `constriction_b200/` imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OK, ERR_IMPOSSIBLE, ERR_INVALID, ERR_TRAILING_ZERO, ERR_NOT_SEALED, ERR_BAD_MODEL, ERR_SEEK = range(7)

u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


class _Ans(C.Structure):
    _fields_ = [("bulk", u32p), ("len", C.c_size_t), ("cap", C.c_size_t), ("state", C.c_uint64)]


class _REnc(C.Structure):
    _fields_ = [("bulk", u32p), ("len", C.c_size_t), ("cap", C.c_size_t), ("lower", C.c_uint64),
                ("range", C.c_uint64), ("num_inverted", C.c_size_t), ("first_inverted", C.c_uint32)]


class _RDec(C.Structure):
    _fields_ = [("bulk", u32p), ("len", C.c_size_t), ("pos", C.c_size_t), ("lower", C.c_uint64),
                ("range", C.c_uint64), ("point", C.c_uint64)]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    hdr = os.path.join(_HERE, "oracle.h")
    stale = (not os.path.exists(so)) or any(
        os.path.getmtime(f) > os.path.getmtime(so) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_erf.restype = C.c_double
        L.orc_erf.argtypes = [C.c_double]
        L.orc_exp.restype = C.c_double
        L.orc_exp.argtypes = [C.c_double]
        L.orc_gaussian_cdf.restype = C.c_double
        L.orc_gaussian_cdf.argtypes = [C.c_double] * 3
        L.orc_qgauss_left_prob.argtypes = [C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, u32p, u32p]
        L.orc_qgauss_cdf.argtypes = [C.c_int32, C.c_int32, C.c_double, C.c_double, u32p]
        L.orc_qgauss_quantile.argtypes = [C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_uint32, i32p, u32p, u32p]
        for sfx, fp in (("f32", f32p), ("f64", f64p)):
            getattr(L, f"orc_cat_cdf_{sfx}").argtypes = [fp, C.c_size_t, u32p]
            getattr(L, f"orc_cat_lazy_left_prob_{sfx}").argtypes = [fp, C.c_size_t, C.c_int32, u32p, u32p]
            getattr(L, f"orc_cat_lazy_quantile_{sfx}").argtypes = [fp, C.c_size_t, C.c_uint32, i32p, u32p, u32p]
        L.orc_log1p.argtypes = [C.c_double]
        L.orc_log1p.restype = C.c_double
        L.orc_atan.argtypes = [C.c_double]
        L.orc_atan.restype = C.c_double
        L.orc_qdist_cdf.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_double, C.c_double, u32p]
        L.orc_binomial_cdf.argtypes = [C.c_int32, C.c_double, u32p]
        L.orc_cat_perfect_cdf_f32.argtypes = [f32p, C.c_size_t, u32p]
        L.orc_cat_perfect_cdf_f64.argtypes = [f64p, C.c_size_t, u32p]
        L.orc_cdf_left_prob.argtypes = [u32p, C.c_size_t, C.c_int64, u32p, u32p]
        L.orc_cdf_quantile.argtypes = [u32p, C.c_size_t, C.c_uint32, C.POINTER(C.c_size_t), u32p, u32p]
        L.orc_cdf_quantile.restype = None
        L.orc_uniform_left_prob.argtypes = [C.c_uint32, C.c_int32, u32p, u32p]
        L.orc_uniform_quantile.argtypes = [C.c_uint32, C.c_uint32, i32p, u32p, u32p]
        L.orc_uniform_quantile.restype = None
        ap = C.POINTER(_Ans)
        for name in ("orc_ans_init", "orc_ans_free", "orc_ans_clear"):
            getattr(L, name).argtypes = [ap]
            getattr(L, name).restype = None
        L.orc_ans_from_compressed.argtypes = [ap, u32p, C.c_size_t]
        L.orc_ans_from_binary.argtypes = [ap, u32p, C.c_size_t]
        L.orc_ans_from_binary.restype = None
        L.orc_ans_encode.argtypes = [ap, C.c_uint32, C.c_uint32]
        L.orc_ans_encode.restype = None
        L.orc_ans_peek_quantile.argtypes = [ap]
        L.orc_ans_peek_quantile.restype = C.c_uint32
        L.orc_ans_decode_advance.argtypes = [ap, C.c_uint32, C.c_uint32]
        L.orc_ans_decode_advance.restype = None
        L.orc_ans_num_words.argtypes = [ap]
        L.orc_ans_num_words.restype = C.c_size_t
        L.orc_ans_num_valid_bits.argtypes = [ap]
        L.orc_ans_num_valid_bits.restype = C.c_size_t
        L.orc_ans_is_empty.argtypes = [ap]
        L.orc_ans_get_compressed.argtypes = [ap, u32p]
        L.orc_ans_get_compressed.restype = C.c_size_t
        L.orc_ans_get_binary.argtypes = [ap, u32p, C.POINTER(C.c_size_t)]
        L.orc_ans_seek.argtypes = [ap, C.c_size_t, C.c_uint64]
        L.orc_ans_encode_iid_reverse.argtypes = [ap, i32p, C.c_size_t, u32p, C.c_int32, C.c_size_t]
        L.orc_ans_decode_iid.argtypes = [ap, i32p, C.c_size_t, u32p, C.c_int32, C.c_size_t]
        L.orc_ans_decode_iid.restype = None
        L.orc_ans_encode_indexed_reverse.argtypes = [ap, i32p, u32p, C.c_size_t, u32p, C.c_size_t, C.c_int32, C.c_size_t]
        L.orc_ans_decode_indexed.argtypes = [ap, i32p, u32p, C.c_size_t, u32p, C.c_size_t, C.c_int32, C.c_size_t]
        L.orc_ans_decode_indexed.restype = None
        L.orc_ans_encode_qgauss_lazy_reverse.argtypes = [ap, i32p, C.c_size_t, C.c_int32, C.c_int32, f64p, f64p, C.c_int]
        L.orc_ans_decode_qgauss_lazy.argtypes = [ap, i32p, C.c_size_t, C.c_int32, C.c_int32, f64p, f64p, C.c_int]
        L.orc_qgauss_quantile_guided.argtypes = [C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_uint32, i32p, u32p, u32p]
        ep = C.POINTER(_REnc)
        for name in ("orc_renc_init", "orc_renc_free", "orc_renc_clear"):
            getattr(L, name).argtypes = [ep]
            getattr(L, name).restype = None
        L.orc_renc_encode.argtypes = [ep, C.c_uint32, C.c_uint32]
        L.orc_renc_num_seal_words.argtypes = [ep]
        L.orc_renc_num_seal_words.restype = C.c_size_t
        L.orc_renc_num_words.argtypes = [ep]
        L.orc_renc_num_words.restype = C.c_size_t
        L.orc_renc_get_compressed.argtypes = [ep, u32p]
        L.orc_renc_get_compressed.restype = C.c_size_t
        L.orc_renc_encode_iid.argtypes = [ep, i32p, C.c_size_t, u32p, C.c_int32, C.c_size_t]
        dp = C.POINTER(_RDec)
        L.orc_rdec_init.argtypes = [dp, u32p, C.c_size_t]
        L.orc_rdec_init.restype = None
        L.orc_rdec_peek_quantile.argtypes = [dp, u32p]
        L.orc_rdec_advance.argtypes = [dp, C.c_uint32, C.c_uint32]
        L.orc_rdec_advance.restype = None
        L.orc_rdec_maybe_exhausted.argtypes = [dp]
        L.orc_rdec_seek.argtypes = [dp, C.c_size_t, C.c_uint64, C.c_uint64]
        L.orc_rdec_decode_iid.argtypes = [dp, i32p, C.c_size_t, u32p, C.c_int32, C.c_size_t]
        menc = [i32p, C.c_uint64, C.c_uint64, C.c_int, u64p, u32p, C.c_int32, C.c_size_t, C.POINTER(u32p), u64p, C.c_int]
        mdec = [u32p, u64p, C.c_uint64, C.c_uint64, C.c_int, u64p, u32p, C.c_int32, C.c_size_t, i32p, C.c_int]
        L.orc_multi_ans_encode.argtypes = menc
        L.orc_multi_range_encode.argtypes = menc
        L.orc_multi_ans_decode.argtypes = mdec
        L.orc_multi_range_decode.argtypes = mdec
        gen_enc = [C.c_uint, C.c_uint, i32p, C.c_size_t, u32p, C.c_int32, C.c_size_t, C.POINTER(u32p), C.POINTER(C.c_size_t)]
        gen_dec = [C.c_uint, C.c_uint, u32p, C.c_size_t, C.c_size_t, u32p, C.c_int32, C.c_size_t, u32p, i32p]
        L.orc_g_ans_encode_iid_reverse.argtypes = gen_enc
        L.orc_g_range_encode_iid.argtypes = gen_enc
        L.orc_g_ans_decode_iid.argtypes = gen_dec
        L.orc_g_range_decode_iid.argtypes = gen_dec
        L.orc_g_lookup_table.argtypes = [C.c_uint, u32p, C.c_size_t, u32p]
        L.orc_g_cat_cdf_f32.argtypes = [C.c_uint, C.c_uint, f32p, C.c_size_t, u32p]
        L.orc_g_cat_cdf_f64.argtypes = [C.c_uint, C.c_uint, f64p, C.c_size_t, u32p]
        L.orc_g_cat_perfect_cdf_f64.argtypes = [C.c_uint, C.c_uint, f64p, C.c_size_t, u32p]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_free.restype = None
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t)


def _raise(rc):
    if rc == OK:
        return
    if rc == ERR_IMPOSSIBLE:
        raise KeyError("Tried to encode symbol that has zero probability under the used entropy model.")
    if rc == ERR_INVALID:
        raise AssertionError("Tried to decode invalid compressed data.")
    if rc == ERR_TRAILING_ZERO:
        raise ValueError("Invalid compressed data: ANS compressed data never ends in a zero word.")
    if rc == ERR_NOT_SEALED:
        raise AssertionError("Cannot unseal compressed data because it doesn't fit into integer number of words.")
    if rc == ERR_BAD_MODEL:
        raise ValueError("Probability distribution not normalizable or invalid model parameter.")
    if rc == ERR_SEEK:
        raise ValueError("Tried to seek past end of stream.")
    raise RuntimeError(f"oracle error {rc}")


# ---------------------------------------------------------------------------
# bulk helpers (numpy in / numpy out)
# ---------------------------------------------------------------------------

def erf(x: float) -> float:
    return lib().orc_erf(float(x))


def exp(x: float) -> float:
    return lib().orc_exp(float(x))


def qgauss_cdf(min_sym: int, max_sym: int, mean: float, std: float) -> np.ndarray:
    n = max_sym - min_sym + 1
    cdf = np.empty(n + 1, dtype=np.uint32)
    _raise(lib().orc_qgauss_cdf(min_sym, max_sym, float(mean), float(std), _p(cdf, u32p)))
    return cdf


def cat_cdf(pmf: np.ndarray) -> np.ndarray:
    pmf = np.ascontiguousarray(pmf)
    cdf = np.empty(pmf.shape[0] + 1, dtype=np.uint32)
    if pmf.dtype == np.float32:
        _raise(lib().orc_cat_cdf_f32(_p(pmf, f32p), pmf.shape[0], _p(cdf, u32p)))
    elif pmf.dtype == np.float64:
        _raise(lib().orc_cat_cdf_f64(_p(pmf, f64p), pmf.shape[0], _p(cdf, u32p)))
    else:
        raise TypeError("pmf must be float32 or float64")
    return cdf


def log1p(x: float) -> float:
    return lib().orc_log1p(float(x))


def atan(x: float) -> float:
    return lib().orc_atan(float(x))


def cat_perfect_cdf(pmf: np.ndarray) -> np.ndarray:
    """categorical.rs:56-177 + contiguous.rs:301-312 (Categorical / Bernoulli with perfect=True)."""
    pmf = np.ascontiguousarray(pmf)
    cdf = np.empty(pmf.shape[0] + 1, dtype=np.uint32)
    if pmf.dtype == np.float32:
        _raise(lib().orc_cat_perfect_cdf_f32(_p(pmf, f32p), pmf.shape[0], _p(cdf, u32p)))
    elif pmf.dtype == np.float64:
        _raise(lib().orc_cat_perfect_cdf_f64(_p(pmf, f64p), pmf.shape[0], _p(cdf, u32p)))
    else:
        raise TypeError("pmf must be float32 or float64")
    return cdf


QDIST_KINDS = {"gaussian": 0, "laplace": 1, "cauchy": 2}


def qdist_cdf(kind: str, min_sym: int, max_sym: int, p0: float, p1: float) -> np.ndarray:
    cdf = np.empty(max_sym - min_sym + 2, dtype=np.uint32)
    _raise(lib().orc_qdist_cdf(QDIST_KINDS[kind], min_sym, max_sym, float(p0), float(p1), _p(cdf, u32p)))
    return cdf


def binomial_cdf(n: int, p: float) -> np.ndarray:
    cdf = np.empty(int(n) + 2, dtype=np.uint32)
    _raise(lib().orc_binomial_cdf(int(n), float(p), _p(cdf, u32p)))
    return cdf


def custom_cdf(cdf_fn, lo: int, hi: int, args=()) -> np.ndarray:
    """LeakyQuantizer over a Python CDF callback (internals.rs:255-420; quantize.rs:525-568)."""
    n = hi - lo + 1
    fw = float((1 << 24) - 1 - (hi - lo))
    row = np.empty(n + 1, dtype=np.uint32)
    row[0], row[n] = 0, 1 << 24
    for i in range(1, n):
        v = fw * float(cdf_fn(float(lo + i) - 0.5, *args))
        q = 0 if not v > 0.0 else (0xFFFFFFFF if v >= 4294967295.0 else int(v))
        row[i] = (q + i) & 0xFFFFFFFF
    return row


def ans_encode_iid(symbols: np.ndarray, cdf: np.ndarray, min_sym: int) -> np.ndarray:
    """One DefaultAnsCoder: encode_iid_symbols_reverse + into_compressed."""
    L = lib()
    c = _Ans()
    L.orc_ans_init(C.byref(c))
    symbols = np.ascontiguousarray(symbols, dtype=np.int32)
    cdf = np.ascontiguousarray(cdf, dtype=np.uint32)
    try:
        _raise(L.orc_ans_encode_iid_reverse(C.byref(c), _p(symbols, i32p), symbols.size, _p(cdf, u32p), min_sym, cdf.size - 1))
        out = np.empty(c.len + 2, dtype=np.uint32)
        n = L.orc_ans_get_compressed(C.byref(c), _p(out, u32p))
        return out[:n].copy()
    finally:
        L.orc_ans_free(C.byref(c))


def ans_encode_qgauss_lazy(symbols: np.ndarray, lo: int, hi: int, mean: float, std: float) -> np.ndarray:
    """One DefaultAnsCoder fed a lazily evaluated QuantizedGaussian: two Gaussian CDFs per symbol, exactly the work
    `encode_iid_symbols_reverse(symbols, &quantizer.quantize(Gaussian))` does in the reference."""
    L = lib()
    c = _Ans()
    L.orc_ans_init(C.byref(c))
    symbols = np.ascontiguousarray(symbols, dtype=np.int32)
    m, s = (C.c_double * 1)(mean), (C.c_double * 1)(std)
    try:
        _raise(L.orc_ans_encode_qgauss_lazy_reverse(C.byref(c), _p(symbols, i32p), symbols.size, lo, hi, m, s, 0))
        out = np.empty(c.len + 2, dtype=np.uint32)
        n = L.orc_ans_get_compressed(C.byref(c), _p(out, u32p))
        return out[:n].copy()
    finally:
        L.orc_ans_free(C.byref(c))


def ans_decode_qgauss_lazy(words: np.ndarray, n: int, lo: int, hi: int, mean: float, std: float) -> np.ndarray:
    """The matching decode: `quantile_function` of the lazily evaluated model per symbol (quantize.rs:580-779)."""
    L = lib()
    c = _Ans()
    L.orc_ans_init(C.byref(c))
    words = np.ascontiguousarray(words, dtype=np.uint32)
    m, s = (C.c_double * 1)(mean), (C.c_double * 1)(std)
    try:
        _raise(L.orc_ans_from_compressed(C.byref(c), _p(words, u32p), words.size))
        out = np.empty(n, dtype=np.int32)
        _raise(L.orc_ans_decode_qgauss_lazy(C.byref(c), _p(out, i32p), n, lo, hi, m, s, 0))
        return out
    finally:
        L.orc_ans_free(C.byref(c))


def qgauss_quantile_guided(lo: int, hi: int, mean: float, std: float, q: int):
    sym, left, prob = C.c_int32(), C.c_uint32(), C.c_uint32()
    _raise(lib().orc_qgauss_quantile_guided(lo, hi, mean, std, q, C.byref(sym), C.byref(left), C.byref(prob)))
    return sym.value, left.value, prob.value


def ans_decode_iid(words: np.ndarray, n: int, cdf: np.ndarray, min_sym: int) -> np.ndarray:
    L = lib()
    c = _Ans()
    L.orc_ans_init(C.byref(c))
    words = np.ascontiguousarray(words, dtype=np.uint32)
    cdf = np.ascontiguousarray(cdf, dtype=np.uint32)
    try:
        _raise(L.orc_ans_from_compressed(C.byref(c), _p(words, u32p), words.size))
        out = np.empty(n, dtype=np.int32)
        L.orc_ans_decode_iid(C.byref(c), _p(out, i32p), n, _p(cdf, u32p), min_sym, cdf.size - 1)
        return out
    finally:
        L.orc_ans_free(C.byref(c))


def ans_encode_indexed(symbols, model_idx, cdfs, min_sym) -> np.ndarray:
    L = lib()
    c = _Ans()
    L.orc_ans_init(C.byref(c))
    symbols = np.ascontiguousarray(symbols, dtype=np.int32)
    model_idx = np.ascontiguousarray(model_idx, dtype=np.uint32)
    cdfs = np.ascontiguousarray(cdfs, dtype=np.uint32)
    try:
        _raise(L.orc_ans_encode_indexed_reverse(C.byref(c), _p(symbols, i32p), _p(model_idx, u32p), symbols.size,
                                                _p(cdfs, u32p), cdfs.shape[1], min_sym, cdfs.shape[1] - 1))
        out = np.empty(c.len + 2, dtype=np.uint32)
        n = L.orc_ans_get_compressed(C.byref(c), _p(out, u32p))
        return out[:n].copy()
    finally:
        L.orc_ans_free(C.byref(c))


def ans_decode_indexed(words, model_idx, cdfs, min_sym) -> np.ndarray:
    L = lib()
    c = _Ans()
    L.orc_ans_init(C.byref(c))
    words = np.ascontiguousarray(words, dtype=np.uint32)
    model_idx = np.ascontiguousarray(model_idx, dtype=np.uint32)
    cdfs = np.ascontiguousarray(cdfs, dtype=np.uint32)
    try:
        _raise(L.orc_ans_from_compressed(C.byref(c), _p(words, u32p), words.size))
        out = np.empty(model_idx.size, dtype=np.int32)
        L.orc_ans_decode_indexed(C.byref(c), _p(out, i32p), _p(model_idx, u32p), model_idx.size, _p(cdfs, u32p),
                                 cdfs.shape[1], min_sym, cdfs.shape[1] - 1)
        return out
    finally:
        L.orc_ans_free(C.byref(c))


def range_encode_iid(symbols: np.ndarray, cdf: np.ndarray, min_sym: int) -> np.ndarray:
    L = lib()
    e = _REnc()
    L.orc_renc_init(C.byref(e))
    symbols = np.ascontiguousarray(symbols, dtype=np.int32)
    cdf = np.ascontiguousarray(cdf, dtype=np.uint32)
    try:
        _raise(L.orc_renc_encode_iid(C.byref(e), _p(symbols, i32p), symbols.size, _p(cdf, u32p), min_sym, cdf.size - 1))
        out = np.empty(L.orc_renc_num_words(C.byref(e)) + 1, dtype=np.uint32)
        n = L.orc_renc_get_compressed(C.byref(e), _p(out, u32p))
        return out[:n].copy()
    finally:
        L.orc_renc_free(C.byref(e))


def range_decode_iid(words: np.ndarray, n: int, cdf: np.ndarray, min_sym: int) -> np.ndarray:
    L = lib()
    d = _RDec()
    words = np.ascontiguousarray(words, dtype=np.uint32)
    cdf = np.ascontiguousarray(cdf, dtype=np.uint32)
    L.orc_rdec_init(C.byref(d), _p(words, u32p), words.size)
    out = np.empty(n, dtype=np.int32)
    _raise(L.orc_rdec_decode_iid(C.byref(d), _p(out, i32p), n, _p(cdf, u32p), min_sym, cdf.size - 1))
    return out


def _multi_encode(fn, symbols, K, cdf, min_sym, sym_offsets, threads):
    L = lib()
    symbols = np.ascontiguousarray(symbols, dtype=np.int32)
    cdf = np.ascontiguousarray(cdf, dtype=np.uint32)
    offsets = np.empty(K + 1, dtype=np.uint64)
    words = u32p()
    if sym_offsets is None:
        inter, so = 1, None
    else:
        sym_offsets = np.ascontiguousarray(sym_offsets, dtype=np.uint64)
        inter, so = 0, _p(sym_offsets, u64p)
    rc = fn(_p(symbols, i32p), symbols.size, K, inter, so, _p(cdf, u32p), min_sym, cdf.size - 1,
            C.byref(words), _p(offsets, u64p), threads)
    total = int(offsets[K])
    out = np.ctypeslib.as_array(words, shape=(max(total, 1),))[:total].copy()
    L.orc_free(words)
    _raise(rc)
    return out, offsets


def _multi_decode(fn, words, offsets, n_total, K, cdf, min_sym, sym_offsets, threads):
    words = np.ascontiguousarray(words, dtype=np.uint32)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    cdf = np.ascontiguousarray(cdf, dtype=np.uint32)
    out = np.empty(n_total, dtype=np.int32)
    if sym_offsets is None:
        inter, so = 1, None
    else:
        sym_offsets = np.ascontiguousarray(sym_offsets, dtype=np.uint64)
        inter, so = 0, _p(sym_offsets, u64p)
    _raise(fn(_p(words, u32p), _p(offsets, u64p), n_total, K, inter, so, _p(cdf, u32p), min_sym, cdf.size - 1,
              _p(out, i32p), threads))
    return out


def multi_ans_encode(symbols, K, cdf, min_sym, sym_offsets=None, threads=1):
    """K independent DefaultAnsCoders (interleaved deal if sym_offsets is None)."""
    return _multi_encode(lib().orc_multi_ans_encode, symbols, K, cdf, min_sym, sym_offsets, threads)


def multi_ans_decode(words, offsets, n_total, K, cdf, min_sym, sym_offsets=None, threads=1):
    return _multi_decode(lib().orc_multi_ans_decode, words, offsets, n_total, K, cdf, min_sym, sym_offsets, threads)


def multi_range_encode(symbols, K, cdf, min_sym, sym_offsets=None, threads=1):
    return _multi_encode(lib().orc_multi_range_encode, symbols, K, cdf, min_sym, sym_offsets, threads)


def multi_range_decode(words, offsets, n_total, K, cdf, min_sym, sym_offsets=None, threads=1):
    return _multi_decode(lib().orc_multi_range_decode, words, offsets, n_total, K, cdf, min_sym, sym_offsets, threads)


# ---------------------------------------------------------------------------
# any preset (oracle_generic.c): PRESETS[name] = (word bits, precision, probability bits)
# ---------------------------------------------------------------------------
PRESETS = {"default": (32, 24, 32), "small": (16, 12, 16)}


def g_cat_cdf(preset: str, pmf: np.ndarray, perfect: bool = False) -> np.ndarray:
    """Contiguous categorical model of a preset: `from_floating_point_probabilities_fast` / `_perfect`."""
    W, P, pb = PRESETS[preset]
    pmf = np.ascontiguousarray(pmf)
    cdf = np.empty(pmf.shape[0] + 1, dtype=np.uint32)
    if perfect:
        _raise(lib().orc_g_cat_perfect_cdf_f64(P, pb, _p(pmf.astype(np.float64), f64p), pmf.shape[0], _p(cdf, u32p)))
    elif pmf.dtype == np.float32:
        _raise(lib().orc_g_cat_cdf_f32(P, pb, _p(pmf, f32p), pmf.shape[0], _p(cdf, u32p)))
    else:
        _raise(lib().orc_g_cat_cdf_f64(P, pb, _p(pmf.astype(np.float64), f64p), pmf.shape[0], _p(cdf, u32p)))
    return cdf


def g_lookup_table(preset: str, cdf: np.ndarray) -> np.ndarray:
    W, P, _ = PRESETS[preset]
    cdf = np.ascontiguousarray(cdf, dtype=np.uint32)
    table = np.empty(1 << P, dtype=np.uint32)
    _raise(lib().orc_g_lookup_table(P, _p(cdf, u32p), cdf.size - 1, _p(table, u32p)))
    return table


def _g_encode(fn, preset, symbols, cdf, min_sym):
    W, P, _ = PRESETS[preset]
    symbols = np.ascontiguousarray(symbols, dtype=np.int32)
    cdf = np.ascontiguousarray(cdf, dtype=np.uint32)
    out, n = u32p(), C.c_size_t()
    _raise(fn(W, P, _p(symbols, i32p), symbols.size, _p(cdf, u32p), min_sym, cdf.size - 1, C.byref(out), C.byref(n)))
    words = np.ctypeslib.as_array(out, shape=(n.value,)).copy() if n.value else np.empty(0, dtype=np.uint32)
    lib().orc_free(out)
    return words


def _g_decode(fn, preset, words, n, cdf, min_sym, table):
    W, P, _ = PRESETS[preset]
    words = np.ascontiguousarray(words, dtype=np.uint32)
    cdf = np.ascontiguousarray(cdf, dtype=np.uint32)
    out = np.empty(n, dtype=np.int32)
    t = None if table is None else np.ascontiguousarray(table, dtype=np.uint32)
    _raise(fn(W, P, _p(words, u32p), words.size, n, _p(cdf, u32p), min_sym, cdf.size - 1, None if t is None else _p(t, u32p),
              _p(out, i32p)))
    return out


def g_ans_encode(preset, symbols, cdf, min_sym=0):
    """One AnsCoder of the preset: encode_iid_symbols_reverse + into_compressed (words as uint32 values)."""
    return _g_encode(lib().orc_g_ans_encode_iid_reverse, preset, symbols, cdf, min_sym)


def g_range_encode(preset, symbols, cdf, min_sym=0):
    return _g_encode(lib().orc_g_range_encode_iid, preset, symbols, cdf, min_sym)


def g_ans_decode(preset, words, n, cdf, min_sym=0, table=None):
    return _g_decode(lib().orc_g_ans_decode_iid, preset, words, n, cdf, min_sym, table)


def g_range_decode(preset, words, n, cdf, min_sym=0, table=None):
    return _g_decode(lib().orc_g_range_decode_iid, preset, words, n, cdf, min_sym, table)


# ---------------------------------------------------------------------------
# entropy models (reference: pybindings/stream/model.rs, model/internals.rs)
# ---------------------------------------------------------------------------

class _Concrete:
    """A fully parameterised model: left_prob(symbol) and quantile(q)."""

    def left_prob(self, symbol):
        raise NotImplementedError

    def quantile(self, q):
        raise NotImplementedError


class _QGauss(_Concrete):
    def __init__(self, lo, hi, mean, std):
        if not std > 0.0:
            raise ValueError("Invalid model parameter: `std` must be positive.")
        self.a = (int(lo), int(hi), float(mean), float(std))

    def left_prob(self, symbol):
        l, p = C.c_uint32(), C.c_uint32()
        _raise(lib().orc_qgauss_left_prob(*self.a, int(symbol), C.byref(l), C.byref(p)))
        return l.value, p.value

    def quantile(self, q):
        s, l, p = C.c_int32(), C.c_uint32(), C.c_uint32()
        _raise(lib().orc_qgauss_quantile(*self.a, int(q), C.byref(s), C.byref(l), C.byref(p)))
        return s.value, l.value, p.value


class _LazyCat(_Concrete):
    def __init__(self, pmf):
        self.pmf = np.ascontiguousarray(pmf)
        if self.pmf.dtype == np.float32:
            self.sfx, self.pt = "f32", f32p
        elif self.pmf.dtype == np.float64:
            self.sfx, self.pt = "f64", f64p
        else:
            raise TypeError("probabilities must be float32 or float64")
        # constructor-time validation (lazy_contiguous.rs:131-167)
        cat_cdf(self.pmf)

    def left_prob(self, symbol):
        l, p = C.c_uint32(), C.c_uint32()
        fn = getattr(lib(), f"orc_cat_lazy_left_prob_{self.sfx}")
        _raise(fn(_p(self.pmf, self.pt), self.pmf.size, int(symbol), C.byref(l), C.byref(p)))
        return l.value, p.value

    def quantile(self, q):
        s, l, p = C.c_int32(), C.c_uint32(), C.c_uint32()
        fn = getattr(lib(), f"orc_cat_lazy_quantile_{self.sfx}")
        _raise(fn(_p(self.pmf, self.pt), self.pmf.size, int(q), C.byref(s), C.byref(l), C.byref(p)))
        return s.value, l.value, p.value


class _TableCat(_Concrete):
    """ContiguousCategoricalEntropyModel::from_floating_point_probabilities_fast."""

    def __init__(self, pmf):
        self.cdf = cat_cdf(np.ascontiguousarray(pmf))

    def left_prob(self, symbol):
        l, p = C.c_uint32(), C.c_uint32()
        _raise(lib().orc_cdf_left_prob(_p(self.cdf, u32p), self.cdf.size - 1, int(symbol), C.byref(l), C.byref(p)))
        return l.value, p.value

    def quantile(self, q):
        i, l, p = C.c_size_t(), C.c_uint32(), C.c_uint32()
        lib().orc_cdf_quantile(_p(self.cdf, u32p), self.cdf.size - 1, int(q), C.byref(i), C.byref(l), C.byref(p))
        return int(i.value), l.value, p.value


class _Table(_Concrete):
    """Any model given as a CDF row over {lo .. lo + n - 1} (contiguous.rs:628-700)."""

    def __init__(self, cdf, lo=0):
        self.cdf = np.ascontiguousarray(cdf, dtype=np.uint32)
        self.lo = int(lo)

    def left_prob(self, symbol):
        l, p = C.c_uint32(), C.c_uint32()
        _raise(lib().orc_cdf_left_prob(_p(self.cdf, u32p), self.cdf.size - 1, int(symbol) - self.lo, C.byref(l), C.byref(p)))
        if p.value == 0:
            _raise(1)  # a padded / empty bin: ImpossibleSymbol
        return l.value, p.value

    def quantile(self, q):
        i, l, p = C.c_size_t(), C.c_uint32(), C.c_uint32()
        lib().orc_cdf_quantile(_p(self.cdf, u32p), self.cdf.size - 1, int(q), C.byref(i), C.byref(l), C.byref(p))
        return self.lo + int(i.value), l.value, p.value


class _Uniform(_Concrete):
    def __init__(self, size):
        self.size = int(size)

    def left_prob(self, symbol):
        l, p = C.c_uint32(), C.c_uint32()
        _raise(lib().orc_uniform_left_prob(self.size, int(symbol), C.byref(l), C.byref(p)))
        return l.value, p.value

    def quantile(self, q):
        s, l, p = C.c_int32(), C.c_uint32(), C.c_uint32()
        lib().orc_uniform_quantile(self.size, int(q), C.byref(s), C.byref(l), C.byref(p))
        return s.value, l.value, p.value


class Model:
    """Base: either concrete (`self._concrete`) or a family (`self._build`, `self._nparams`)."""
    _concrete = None
    _nparams = 0

    def as_parameterized(self):
        if self._concrete is None:
            raise ValueError("No model parameters specified.")
        return self._concrete

    def length(self, param0):
        if self._concrete is not None:
            raise ValueError("Model parameters were specified but the model is already fully parameterized.")
        return len(param0)

    def parameterize(self, params, reverse):
        if self._concrete is not None:
            raise ValueError("Model parameters were specified but the model is already fully parameterized.")
        if len(params) != self._nparams:
            raise ValueError(f"Wrong number of model parameters: expected {self._nparams}, got {len(params)}.")
        cols = [self._cast(p) for p in params]
        n = len(cols[0])
        if any(len(c) != n for c in cols):
            raise ValueError("Model parameters have unequal shape")
        order = range(n - 1, -1, -1) if reverse else range(n)
        for i in order:
            yield self._build(*[c[i] for c in cols])

    @staticmethod
    def _cast(p):
        p = np.asarray(p)
        if p.ndim != 1 or p.dtype not in (np.float32, np.float64):
            raise TypeError("model parameters must be rank-1 float32/float64 arrays")
        return p.astype(np.float64)  # internals.rs:169-174: f32 params are widened to f64


class QuantizedGaussian(Model):
    def __init__(self, min_symbol_inclusive, max_symbol_inclusive, mean=None, std=None):
        lo, hi = int(min_symbol_inclusive), int(max_symbol_inclusive)
        if mean is not None and std is not None:
            self._concrete = _QGauss(lo, hi, mean, std)
        elif mean is None and std is None:
            self._nparams, self._build = 2, lambda m, s: _QGauss(lo, hi, m, s)
        elif mean is None:
            self._nparams, self._build = 1, lambda m: _QGauss(lo, hi, m, std)
        else:
            self._nparams, self._build = 1, lambda s: _QGauss(lo, hi, mean, s)


class Uniform(Model):
    def __init__(self, size=None):
        if size is not None:
            self._concrete = _Uniform(size)
        else:
            self._nparams, self._build = 1, lambda s: _Uniform(s)

    @staticmethod
    def _cast(p):
        p = np.asarray(p)
        if p.ndim != 1 or p.dtype != np.int32:
            raise TypeError("size must be a rank-1 int32 array")
        return p


class Categorical(Model):
    def __init__(self, probabilities=None, lazy=None, perfect=None):
        if lazy is None and perfect is None:
            lazy, perfect = False, True  # model.rs:509-526
        elif lazy and perfect:
            raise ValueError("Both arguments `lazy` and `perfect` cannot be set to `True` at the same time.")
        else:
            lazy, perfect = bool(lazy), bool(perfect)
        self._perfect = perfect
        self._lazy = lazy
        if probabilities is not None:
            probabilities = np.asarray(probabilities)
            if probabilities.ndim != 1:
                raise TypeError("probabilities must be rank 1")
            if perfect:
                self._concrete = _Table(cat_perfect_cdf(probabilities))
            else:
                self._concrete = _LazyCat(probabilities) if lazy else _TableCat(probabilities)

    def length(self, param0):
        if self._concrete is not None:
            raise ValueError("Model parameters were specified but the model is already fully parameterized.")
        return np.asarray(param0).shape[0]

    def parameterize(self, params, reverse):
        if self._concrete is not None:
            raise ValueError("Model parameters were specified but the model is already fully parameterized.")
        if len(params) != 1:
            raise ValueError(f"Wrong number of model parameters: expected 1, got {len(params)}.")
        probs = np.ascontiguousarray(params[0])
        if probs.ndim != 2 or probs.dtype not in (np.float32, np.float64):
            raise TypeError("probabilities must be a rank-2 float32/float64 array")
        order = range(probs.shape[0] - 1, -1, -1) if reverse else range(probs.shape[0])
        for i in order:
            # internals.rs:431-460: perfectly quantised table if `perfect`, else always lazy
            yield _Table(cat_perfect_cdf(probs[i])) if self._perfect else _LazyCat(probs[i])


class Bernoulli(Model):
    """pybindings/stream/model.rs:985-1060"""

    def __init__(self, p=None, perfect=None):
        perfect = True if perfect is None else bool(perfect)
        make = (lambda q: _Table(cat_perfect_cdf(np.array([1.0 - q, q])))) if perfect else \
            (lambda q: _TableCat(np.array([1.0 - q, q])))
        if p is not None:
            self._concrete = make(float(p))
        else:
            self._nparams, self._build = 1, make


class _TwoParam(Model):
    _kind = None

    def __init__(self, lo, hi, a=None, b=None):
        lo, hi, kind = int(lo), int(hi), self._kind

        def make(x, y):
            if not y > 0.0:
                raise ValueError("Invalid model parameter: `scale` must be positive.")
            return _Table(qdist_cdf(kind, lo, hi, x, y), lo)

        if a is not None and b is not None:
            self._concrete = make(float(a), float(b))
        elif a is None and b is None:
            self._nparams, self._build = 2, make
        elif a is None:
            self._nparams, self._build = 1, lambda x: make(x, float(b))
        else:
            self._nparams, self._build = 1, lambda y: make(float(a), y)


class QuantizedLaplace(_TwoParam):
    _kind = "laplace"

    def __init__(self, min_symbol_inclusive, max_symbol_inclusive, mean=None, scale=None):
        super().__init__(min_symbol_inclusive, max_symbol_inclusive, mean, scale)


class QuantizedCauchy(_TwoParam):
    _kind = "cauchy"

    def __init__(self, min_symbol_inclusive, max_symbol_inclusive, loc=None, scale=None):
        super().__init__(min_symbol_inclusive, max_symbol_inclusive, loc, scale)


class Binomial(Model):
    def __init__(self, n=None, p=None):
        make = lambda nn, pp: _Table(binomial_cdf(int(nn), float(pp)))  # noqa: E731
        self._n, self._p = n, p
        if n is not None and p is not None:
            self._concrete = make(n, p)
        elif n is None and p is None:
            self._nparams, self._build = 2, make
        elif n is None:
            self._nparams, self._build = 1, lambda nn: make(nn, p)
        else:
            self._nparams, self._build = 1, lambda pp: make(n, pp)

    def parameterize(self, params, reverse):
        if self._concrete is not None:
            raise ValueError("Model parameters were specified but the model is already fully parameterized.")
        if len(params) != self._nparams:
            raise ValueError(f"Wrong number of model parameters: expected {self._nparams}, got {len(params)}.")
        cols = [np.asarray(p) for p in params]
        n = len(cols[0])
        if any(len(c) != n for c in cols):
            raise ValueError("Model parameters have unequal shape")
        order = range(n - 1, -1, -1) if reverse else range(n)
        for i in order:
            yield self._build(*[c[i] for c in cols])


class CustomModel(Model):
    def __init__(self, cdf, approximate_inverse_cdf, min_symbol_inclusive, max_symbol_inclusive):
        self._cdf, self._lo, self._hi = cdf, int(min_symbol_inclusive), int(max_symbol_inclusive)
        self._nparams = None

    def as_parameterized(self):
        return _Table(custom_cdf(self._cdf, self._lo, self._hi), self._lo)

    def length(self, param0):
        return len(param0)

    def parameterize(self, params, reverse):
        cols = [self._cast(p) for p in params]
        n = len(cols[0])
        if any(len(c) != n for c in cols):
            raise ValueError("Model parameters have unequal lengths.")
        order = range(n - 1, -1, -1) if reverse else range(n)
        for i in order:
            yield _Table(custom_cdf(self._cdf, self._lo, self._hi, tuple(float(c[i]) for c in cols)), self._lo)


class ScipyModel(CustomModel):
    def __init__(self, scipy_model, min_symbol_inclusive, max_symbol_inclusive):
        super().__init__(scipy_model.cdf, scipy_model.ppf, min_symbol_inclusive, max_symbol_inclusive)


# ---------------------------------------------------------------------------
# coders (reference: pybindings/stream/stack.rs, queue.rs)
# ---------------------------------------------------------------------------

def _is_scalar(x):
    return isinstance(x, (int, np.integer)) and not isinstance(x, bool)


def _symbols_array(symbols):
    a = np.asarray(symbols)
    if a.ndim != 1 or a.dtype != np.int32:
        raise TypeError("symbols must be a rank-1 numpy array with dtype=np.int32")
    return a


class AnsCoder:
    def __init__(self, compressed=None, seal=False):
        L = lib()
        self._c = _Ans()
        L.orc_ans_init(C.byref(self._c))
        if compressed is None:
            if seal:
                raise ValueError("Need compressed data to seal.")
            return
        w = np.ascontiguousarray(compressed)
        if w.dtype != np.uint32 or w.ndim != 1:
            raise TypeError("compressed must be a rank-1 uint32 array")
        if seal:
            L.orc_ans_from_binary(C.byref(self._c), _p(w, u32p), w.size)
        else:
            _raise(L.orc_ans_from_compressed(C.byref(self._c), _p(w, u32p), w.size))

    def __del__(self):
        try:
            lib().orc_ans_free(C.byref(self._c))
        except Exception:
            pass

    def pos(self):
        return (int(self._c.len), int(self._c.state))

    def seek(self, position, state):
        _raise(lib().orc_ans_seek(C.byref(self._c), int(position), int(state)))

    def clear(self):
        lib().orc_ans_clear(C.byref(self._c))

    def num_words(self):
        return int(lib().orc_ans_num_words(C.byref(self._c)))

    def num_bits(self):
        return 32 * self.num_words()

    def num_valid_bits(self):
        return int(lib().orc_ans_num_valid_bits(C.byref(self._c)))

    def is_empty(self):
        return bool(lib().orc_ans_is_empty(C.byref(self._c)))

    def get_compressed(self, unseal=False):
        out = np.empty(self._c.len + 2, dtype=np.uint32)
        if unseal:
            n = C.c_size_t()
            _raise(lib().orc_ans_get_binary(C.byref(self._c), _p(out, u32p), C.byref(n)))
            return out[:n.value].copy()
        n = lib().orc_ans_get_compressed(C.byref(self._c), _p(out, u32p))
        return out[:n].copy()

    def _enc(self, symbol, m):
        l, p = m.left_prob(symbol)
        lib().orc_ans_encode(C.byref(self._c), l, p)

    def _dec(self, m):
        s, l, p = m.quantile(lib().orc_ans_peek_quantile(C.byref(self._c)))
        lib().orc_ans_decode_advance(C.byref(self._c), l, p)
        return s

    def encode_reverse(self, symbols, model, *params):
        if _is_scalar(symbols):
            if params:
                raise ValueError("To encode a single symbol, use a concrete model.")
            self._enc(int(symbols), model.as_parameterized())
            return
        symbols = _symbols_array(symbols)
        if not params:
            m = model.as_parameterized()
            for s in symbols[::-1]:
                self._enc(int(s), m)
        else:
            if symbols.size != model.length(params[0]):
                raise ValueError("`symbols` argument has wrong length.")
            for s, m in zip(symbols[::-1], model.parameterize(params, True)):
                self._enc(int(s), m)

    def decode(self, model, *params):
        if len(params) == 0:
            return self._dec(model.as_parameterized())
        if len(params) == 1 and _is_scalar(params[0]):
            m = model.as_parameterized()
            return np.array([self._dec(m) for _ in range(int(params[0]))], dtype=np.int32)
        return np.array([self._dec(m) for m in model.parameterize(params, False)], dtype=np.int32)

    def clone(self):
        c = AnsCoder()
        w = self.get_compressed()
        if w.size:
            _raise(lib().orc_ans_from_compressed(C.byref(c._c), _p(w, u32p), w.size))
        return c


class RangeEncoder:
    def __init__(self):
        self._e = _REnc()
        lib().orc_renc_init(C.byref(self._e))

    def __del__(self):
        try:
            lib().orc_renc_free(C.byref(self._e))
        except Exception:
            pass

    def clear(self):
        lib().orc_renc_clear(C.byref(self._e))

    def pos(self):
        return (int(self._e.len + self._e.num_inverted), (int(self._e.lower), int(self._e.range)))

    def num_words(self):
        return int(lib().orc_renc_num_words(C.byref(self._e)))

    def num_bits(self):
        return 32 * self.num_words()

    def is_empty(self):
        return self._e.range == 2**64 - 1 and self._e.len == 0

    def get_compressed(self):
        out = np.empty(self.num_words() + 1, dtype=np.uint32)
        n = lib().orc_renc_get_compressed(C.byref(self._e), _p(out, u32p))
        return out[:n].copy()

    def get_decoder(self):
        return RangeDecoder(self.get_compressed())

    def _enc(self, symbol, m):
        l, p = m.left_prob(symbol)
        _raise(lib().orc_renc_encode(C.byref(self._e), l, p))

    def encode(self, symbols, model, *params):
        if _is_scalar(symbols):
            if params:
                raise ValueError("To encode a single symbol, use a concrete model.")
            self._enc(int(symbols), model.as_parameterized())
            return
        symbols = _symbols_array(symbols)
        if not params:
            m = model.as_parameterized()
            for s in symbols:
                self._enc(int(s), m)
        else:
            if symbols.size != model.length(params[0]):
                raise ValueError("`symbols` argument has wrong length.")
            for s, m in zip(symbols, model.parameterize(params, False)):
                self._enc(int(s), m)


class RangeDecoder:
    def __init__(self, compressed):
        w = np.ascontiguousarray(compressed)
        if w.dtype != np.uint32 or w.ndim != 1:
            raise TypeError("compressed must be a rank-1 uint32 array")
        self._w = w.copy()
        self._d = _RDec()
        lib().orc_rdec_init(C.byref(self._d), _p(self._w, u32p), self._w.size)

    def seek(self, position, state):
        lower, rng = state
        rc = lib().orc_rdec_seek(C.byref(self._d), int(position), int(lower), int(rng))
        if rc:
            raise ValueError("Invalid coder state or tried to seek past end of stream.")

    def maybe_exhausted(self):
        return bool(lib().orc_rdec_maybe_exhausted(C.byref(self._d)))

    def _dec(self, m):
        q = C.c_uint32()
        _raise(lib().orc_rdec_peek_quantile(C.byref(self._d), C.byref(q)))
        s, l, p = m.quantile(q.value)
        lib().orc_rdec_advance(C.byref(self._d), l, p)
        return s

    def decode(self, model, *params):
        if len(params) == 0:
            return self._dec(model.as_parameterized())
        if len(params) == 1 and _is_scalar(params[0]):
            m = model.as_parameterized()
            return np.array([self._dec(m) for _ in range(int(params[0]))], dtype=np.int32)
        return np.array([self._dec(m) for m in model.parameterize(params, False)], dtype=np.int32)
