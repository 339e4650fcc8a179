"""The C-ABI library loads without a GPU and exports every symbol include/constriction_b200.h declares
(no compute calls here)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "constriction_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ctr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from constriction_b200 import _native as N
    lib = N.load()
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(names) == set(N.SIGNATURES), set(names) ^ set(N.SIGNATURES)
    assert lib.ctr_abi_version() == 3
    assert lib.ctr_status_string(1).decode().startswith("Tried to encode symbol")


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every compute entry point fails loudly (CTR_ERR_CUDA), never computes."""
    import torch
    if torch.cuda.is_available():
        return
    from constriction_b200 import _native as N
    lib = N.load()
    assert lib.ctr_device_count() == 0
    out = C.c_void_p()
    means = (C.c_double * 1)(0.0)
    stds = (C.c_double * 1)(1.0)
    rc = lib.ctr_model_quantized_gaussian(-5, 5, means, stds, 1, None, C.byref(out))
    assert rc == N.ERR_CUDA
    L = N.Layout()
    L.n_streams, L.n_symbols = 1, 1
    off = (C.c_uint64 * 2)()
    rc = lib.ctr_ans_decode(C.c_void_p(1), None, off, C.byref(L), None, C.c_void_p(1), None, None, None, None)
    assert rc == N.ERR_CUDA


def test_product_never_imports_the_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may touch oracle/ (it is the checker, not the product)."""
    pkg = os.path.join(ROOT, "constriction_b200")
    pattern = re.compile(r"(import\s+oracle|from\s+oracle|liboracle|oracle[/.]refapi|#include\s+\"[^\"]*oracle)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pattern.search(text), (dirpath, f)
    header = open(os.path.join(ROOT, "include", "constriction_b200.h")).read()
    assert not pattern.search(header)


def test_call_level_argument_checks():
    """Call-level errors of the newer entry points are reported before any device work (so they can be checked
    here): checkpoints need the contiguous layout and a multiple of 32; the table-free Gaussian entry points take
    no model index, a non-empty support and fewer than 2^32 symbols."""
    from constriction_b200 import _native as N
    lib = N.load()
    one = C.c_void_p(16)  # a non-null dummy pointer; none of these calls may dereference it

    def layout(**kw):
        L = N.Layout()
        L.n_streams, L.n_symbols = 4, 4096
        for k, v in kw.items():
            setattr(L, k, v)
        return L

    enc_args = (None, one, 1 << 20, one, 1 << 20, one, None, None, None)
    # checkpoints: interleaved layout / C not a multiple of 32 / missing record arrays
    for L in (layout(flags=N.FLAG_CHECKPOINTS, checkpoint_every=64, ckpt_offsets_dev=16, checkpoints_dev=16),
              layout(flags=N.FLAG_CHECKPOINTS, sym_offsets_dev=16, checkpoint_every=48, ckpt_offsets_dev=16, checkpoints_dev=16),
              layout(flags=N.FLAG_CHECKPOINTS, sym_offsets_dev=16, checkpoint_every=64)):
        assert lib.ctr_ans_encode_reverse(one, one, C.byref(L), *enc_args) == N.ERR_BAD_ARGUMENT
        assert lib.ctr_range_decode(one, one, one, C.byref(L), None, one, None, None, None, None) == N.ERR_BAD_ARGUMENT
    assert lib.ctr_checkpoint_offsets(C.byref(layout(checkpoint_every=64)), one, None) == N.ERR_BAD_ARGUMENT  # no sym_offsets
    assert lib.ctr_checkpoint_offsets(C.byref(layout(sym_offsets_dev=16, checkpoint_every=33)), one, None) == N.ERR_BAD_ARGUMENT
    assert lib.ctr_checkpoint_max_records(C.byref(layout(checkpoint_every=64))) == 4096 // 64 + 4
    assert lib.ctr_checkpoint_max_records(C.byref(layout())) == 0
    # table-free Gaussian entry points
    g_enc = (one, one, one)
    L = layout(model_index_dev=16, model_index_mode=N.INDEX_PER_SYMBOL)
    assert lib.ctr_ans_encode_reverse_gaussian(-5, 5, *g_enc, C.byref(L), *enc_args) == N.ERR_BAD_ARGUMENT
    L = layout()
    assert lib.ctr_ans_encode_reverse_gaussian(5, 5, *g_enc, C.byref(L), *enc_args) == N.ERR_BAD_MODEL
    assert lib.ctr_range_decode_gaussian(7, -7, one, one, one, one, C.byref(L), None, one, None, None, None, None) == N.ERR_BAD_MODEL
    assert lib.ctr_ans_decode_gaussian(-5, 5, None, one, one, one, C.byref(L), None, one, None, None, None, None) == N.ERR_BAD_ARGUMENT
    big = layout(n_symbols=1 << 32)
    assert lib.ctr_range_encode_gaussian(-5, 5, *g_enc, C.byref(big), *enc_args) == N.ERR_BAD_ARGUMENT
    # stream memory operations: null address
    assert lib.ctr_stream_write_value32(None, 1, None) == N.ERR_BAD_ARGUMENT
    # gather: null / inconsistent arguments
    out = C.c_void_p()
    assert lib.ctr_gather_create(0, 0, 2, 64, 4, one, one, one, C.byref(out)) == N.ERR_BAD_ARGUMENT
    assert lib.ctr_gather_create(2, 2, 2, 64, 4, one, one, one, C.byref(out)) == N.ERR_BAD_ARGUMENT
    assert lib.ctr_gather_push(None, 1, 1, None) == N.ERR_BAD_ARGUMENT
    assert lib.ctr_gather_compressed_nccl(None, 2, 0, one, one, 1, 64, 4, one, one, one, one, None) == N.ERR_BAD_ARGUMENT
