#!/usr/bin/env python
"""Times the host-buffer C ABI piece by piece on one GPU: encode alone, decode alone, both at once (async jobs),
and plain pinned copies in each direction / both directions.  python tools/e2e_diag.py [symbols]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from constriction_b200 import _native as N  # noqa: E402
from constriction_b200 import batch as B  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
k = 148 * 1024
lib = N.load()
g = torch.Generator(device="cuda")
g.manual_seed(2)
syms = torch.clamp(torch.round(torch.randn(n, device="cuda", generator=g) * 9.6 + 3.2), -50, 50).to(torch.int32)
model = B.ModelTable.quantized_gaussian(-50, 50, [3.2], [9.6])
h_syms = torch.empty(n, dtype=torch.int32).pin_memory()
h_syms.copy_(syms)
h_out = torch.empty(n, dtype=torch.int32).pin_memory()
cap = n // 4 + 4 * k
conts = [dict(words=torch.empty(cap, dtype=torch.int32).pin_memory(), off=torch.empty(k + 1, dtype=torch.int64).pin_memory(),
              st=C.c_int(), bad=C.c_uint64()) for _ in range(2)]
dst, dbad = C.c_int(), C.c_uint64()


def enc_args(ct):
    return (model.handle, h_syms.data_ptr(), n, k, None, None, 0, ct["words"].data_ptr(), cap, ct["off"].data_ptr(), C.byref(ct["st"]), C.byref(ct["bad"]))


def dec_args(ct):
    return (model.handle, ct["words"].data_ptr(), ct["off"].data_ptr(), n, k, None, None, 0, h_out.data_ptr(), C.byref(dst), C.byref(dbad))


def timeit(name, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name:48s} {dt * 1e3:8.2f} ms", flush=True)
    return dt


def both():
    j1, j2 = C.c_void_p(), C.c_void_p()
    if os.environ.get("DEC_FIRST"):
        assert lib.ctr_ans_decode_host_async(*dec_args(conts[0]), C.byref(j2)) == 0
        assert lib.ctr_ans_encode_reverse_host_async(*enc_args(conts[1]), C.byref(j1)) == 0
    else:
        assert lib.ctr_ans_encode_reverse_host_async(*enc_args(conts[1]), C.byref(j1)) == 0
        assert lib.ctr_ans_decode_host_async(*dec_args(conts[0]), C.byref(j2)) == 0
    assert lib.ctr_host_job_wait(j1) == 0 and lib.ctr_host_job_wait(j2) == 0


d_a = torch.empty(n, dtype=torch.int32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_syms, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(syms, non_blocking=True)


def duplex():
    h2d()
    d2h()


def h2d_2d(chunks=8):  # the strided copies the interleaved pipeline issues
    rows = n // k
    kc = k // chunks
    for c in range(chunks):
        torch.cuda.cudart().cudaMemcpy2DAsync if False else None
    return None


for chunks in os.environ.get("DIAG_CHUNKS", "8").split(","):
    os.environ["CTR_HOST_CHUNKS"] = chunks
    print(f"-- CTR_HOST_CHUNKS={chunks}")
    timeit("encode_host", lambda: lib.ctr_ans_encode_reverse_host(*enc_args(conts[0])))
    timeit("decode_host", lambda: lib.ctr_ans_decode_host(*dec_args(conts[0])))
    assert torch.equal(h_out, h_syms)
    timeit("encode_async + decode_async", both)
timeit("pinned H2D 4n bytes", h2d)
timeit("pinned D2H 4n bytes", d2h)
timeit("both directions at once", duplex)
