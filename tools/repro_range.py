import sys, os
sys.path.insert(0, os.getcwd())
import torch
from constriction_b200 import batch as B
bc = B.BatchCoder()
model = B.ModelTable.quantized_gaussian(-50, 50, [3.2], [9.6])
def run(name, fn):
    try:
        fn(); torch.cuda.synchronize(); bc.check(); print(name, "ok", flush=True)
    except Exception as e:
        print(name, "FAILED", repr(e)[:200], flush=True); raise SystemExit(1)
per, k = int(os.environ.get("PER", 122070)), int(os.environ.get("K", 1024))
g = torch.Generator(device="cuda"); g.manual_seed(0)
symbols = torch.randint(-50, 51, (k * per,), dtype=torch.int32, device="cuda", generator=g)
offsets = torch.arange(0, k + 1, device="cuda") * per
st = {}
run("range_encode plain", lambda: st.__setitem__("a", bc.range_encode(symbols, model, sym_offsets=offsets)))
run("range_decode plain", lambda: st.__setitem__("o", bc.range_decode(st["a"], model)))
assert torch.equal(st["o"], symbols)
run("range_encode ckpt", lambda: st.__setitem__("c", bc.range_encode(symbols, model, sym_offsets=offsets, checkpoint_every=1024)))
assert torch.equal(st["c"].words[:st["c"].total_words()], st["a"].words[:st["a"].total_words()])
run("range_decode ckpt", lambda: st.__setitem__("o2", bc.range_decode(st["c"], model)))
assert torch.equal(st["o2"], symbols)
run("ans_encode ckpt", lambda: st.__setitem__("d", bc.ans_encode(symbols, model, sym_offsets=offsets, checkpoint_every=1024)))
run("ans_decode ckpt", lambda: st.__setitem__("o3", bc.ans_decode(st["d"], model)))
assert torch.equal(st["o3"], symbols)
print("all ok")
