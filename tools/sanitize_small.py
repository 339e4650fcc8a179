"""Small invocations of every kernel path added in this round, for compute-sanitizer (memcheck / racecheck):
TMA symbol boxes (ANS + range encoders, opt-in TMA decoder), checkpoints + chunk-parallel decode, table-free Gaussian
encode / decode, bad offsets.   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from constriction_b200 import batch as B  # noqa: E402

bc = B.BatchCoder()
rng = np.random.default_rng(0)
model = B.ModelTable.quantized_gaussian(-50, 50, [3.2], [9.6])


def gauss(n):
    return torch.from_numpy(np.clip(np.rint(rng.normal(3.2, 9.6, n)), -50, 50).astype(np.int32)).cuda()


# interleaved, TMA boxes: K multiple of 4 but not of 32, rows above the highest box, ragged last row
for n, k in ((40_003, 100), (9 * 64, 64), (30_000, 256)):
    s = gauss(n)
    for enc, dec in ((bc.ans_encode, bc.ans_decode), (bc.range_encode, bc.range_decode)):
        c = enc(s, model, n_streams=k)
        assert torch.equal(dec(c, model), s)
# contiguous + checkpoints, per-stream model index
lengths = [1000, 0, 64, 65, 1, 31, 2048 + 17, 300]
off = torch.from_numpy(np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)).cuda()
s = gauss(int(off[-1]))
pool = B.ModelTable.quantized_gaussian(-50, 50, rng.normal(0, 10, 4), np.exp(rng.uniform(0, 3, 4)))
idx = torch.from_numpy(rng.integers(0, 4, size=len(lengths)).astype(np.int32)).cuda()
for enc, dec in ((bc.ans_encode, bc.ans_decode), (bc.range_encode, bc.range_decode)):
    c = enc(s, pool, sym_offsets=off, model_index=idx, index_mode=2, checkpoint_every=64)
    assert torch.equal(dec(c, pool, model_index=idx, index_mode=2), s)
# table-free Gaussian parameters, both layouts
n = 6000
lazy = B.GaussianParams(-50, 50, rng.normal(0, 20, n), np.exp(rng.uniform(-3, 4, n)))
s = gauss(n)
off2 = torch.tensor([0, 1000, 1000, 4500, n], dtype=torch.int64, device="cuda")
for enc, dec in ((bc.ans_encode, bc.ans_decode), (bc.range_encode, bc.range_decode)):
    for kw in (dict(n_streams=33), dict(sym_offsets=off2), dict(sym_offsets=off2, checkpoint_every=32)):
        c = enc(s, lazy, **kw)
        assert torch.equal(dec(c, lazy), s)
bc.check()
# offsets that leave the symbol array are rejected without being followed
bc.ans_encode(s, model, sym_offsets=off2 * 50)
try:
    bc.check()
    raise SystemExit("bad offsets were not flagged")
except ValueError:
    pass
torch.cuda.synchronize()
print("sanitize_small: all paths ok")
