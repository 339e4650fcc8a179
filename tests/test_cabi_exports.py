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
    assert lib.ctr_abi_version() == 2
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
