"""One encode / decode of BASELINE configs[3]'s per-GPU shard (1024 range streams x 122,070 symbols) for ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from constriction_b200 import batch as B
k, per = 1024, int(os.environ.get("PER", "122070"))
coder = os.environ.get("CODER", "range")
g = torch.Generator(device="cuda"); g.manual_seed(4)
syms = torch.clamp(torch.round(torch.randn(k * per, device="cuda", generator=g) * 9.6 + 3.2), -50, 50).to(torch.int32)
off = torch.arange(k + 1, device="cuda", dtype=torch.int64) * per
model = B.ModelTable.quantized_gaussian(-50, 50, [3.2], [9.6])
bc = B.BatchCoder()
enc, dec = (bc.range_encode, bc.range_decode) if coder == "range" else (bc.ans_encode, bc.ans_decode)
for _ in range(2):
    comp = enc(syms, model, sym_offsets=off)
    out = dec(comp, model)
torch.cuda.synchronize()
bc.check()
assert torch.equal(out, syms)
