import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from constriction_b200 import batch as B
k, per = int(os.environ.get("K", "1024")), int(os.environ.get("PER", "122070"))
g = torch.Generator(device="cuda"); g.manual_seed(4)
syms = torch.clamp(torch.round(torch.randn(k * per, device="cuda", generator=g) * 9.6 + 3.2), -50, 50).to(torch.int32)
off = torch.arange(k + 1, device="cuda", dtype=torch.int64) * per
model = B.ModelTable.quantized_gaussian(-50, 50, [3.2], [9.6])
bc = B.BatchCoder()
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, enc, dec in (("range", bc.range_encode, bc.range_decode), ("ans", bc.ans_encode, bc.ans_decode)):
    st = {}
    def e(): st["c"] = enc(syms, model, sym_offsets=off, out=st.get("c"))
    ms_e = timed(e)
    out = torch.empty_like(syms)
    ms_d = timed(lambda: dec(st["c"], model, out=out))
    bc.check()
    print(f"CTR_CHAIN={os.environ.get('CTR_CHAIN')} {name}: K={k} per={per} encode {ms_e:.3f} ms ({ms_e*1e6/per:.1f} ns/sym/stream) decode {ms_d:.3f} ms ({ms_d*1e6/per:.1f} ns/sym/stream) ok={bool(torch.equal(out, syms))}", flush=True)
