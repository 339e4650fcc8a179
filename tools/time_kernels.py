#!/usr/bin/env python
"""Kernel-only timing of the bench workload without result checks (for diagnostic builds that break parity on
purpose, e.g. -DCTR_DBG_NO_GATHER).  Prints one JSON line: average kernel times from the library's own CUDA events."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from constriction_b200 import _native as N  # noqa: E402
from constriction_b200 import batch as B  # noqa: E402

n = int(os.environ.get("N", 100_000_000))
k = int(os.environ.get("K", 148 * 1024))
reps = int(os.environ.get("REPS", 20))
what = os.environ.get("WHAT", "ans")
lib = N.load()
g = torch.Generator(device="cuda")
g.manual_seed(2)
syms = torch.clamp(torch.round(torch.randn(n, device="cuda", generator=g) * 9.6 + 3.2), -50, 50).to(torch.int32)
model = B.ModelTable.quantized_gaussian(-50, 50, [3.2], [9.6])
bc = B.BatchCoder()
enc, dec = (bc.ans_encode, bc.ans_decode) if what == "ans" else (bc.range_encode, bc.range_decode)
slots = (0, 1) if what == "ans" else (2, 3)
comp = enc(syms, model, n_streams=k)
out = torch.empty_like(syms)
for _ in range(3):
    comp = enc(syms, model, n_streams=k, out=comp)
    if "--encode-only" not in sys.argv:
        dec(comp, model, out=out)
torch.cuda.synchronize()
lib.ctr_profile_enable(1)
for s in slots:
    lib.ctr_profile_read(s, None, None)
for _ in range(reps):
    comp = enc(syms, model, n_streams=k, out=comp)
    if "--encode-only" not in sys.argv:
        dec(comp, model, out=out)
torch.cuda.synchronize()
res = {}
for name, s in zip(("encode_ms", "decode_ms"), slots):
    ms, cnt = C.c_double(), C.c_uint64()
    lib.ctr_profile_read(s, C.byref(ms), C.byref(cnt))
    res[name] = round(ms.value / max(cnt.value, 1), 4)
res["flags"] = os.environ.get("CTR_EXTRA_NVCC_FLAGS", "")
print(json.dumps(res))
