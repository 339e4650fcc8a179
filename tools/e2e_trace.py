import ctypes as C, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from constriction_b200 import _native as N, batch as B
n, k = 100_000_000, 148 * 1024
lib = N.load()
g = torch.Generator(device="cuda"); g.manual_seed(2)
syms = torch.clamp(torch.round(torch.randn(n, device="cuda", generator=g) * 9.6 + 3.2), -50, 50).to(torch.int32)
model = B.ModelTable.quantized_gaussian(-50, 50, [3.2], [9.6])
h_syms = torch.empty(n, dtype=torch.int32).pin_memory(); h_syms.copy_(syms)
h_out = torch.empty(n, dtype=torch.int32).pin_memory()
cap = n // 4 + 4 * k
conts = [dict(words=torch.empty(cap, dtype=torch.int32).pin_memory(), off=torch.empty(k + 1, dtype=torch.int64).pin_memory(), st=C.c_int(), bad=C.c_uint64()) for _ in range(2)]
dst, dbad = C.c_int(), C.c_uint64()
enc_args = lambda ct: (model.handle, h_syms.data_ptr(), n, k, None, None, 0, ct["words"].data_ptr(), cap, ct["off"].data_ptr(), C.byref(ct["st"]), C.byref(ct["bad"]))
dec_args = lambda ct: (model.handle, ct["words"].data_ptr(), ct["off"].data_ptr(), n, k, None, None, 0, h_out.data_ptr(), C.byref(dst), C.byref(dbad))
for _ in range(2):
    lib.ctr_ans_encode_reverse_host(*enc_args(conts[0])); lib.ctr_ans_decode_host(*dec_args(conts[0]))
def both():
    j1, j2 = C.c_void_p(), C.c_void_p()
    lib.ctr_ans_encode_reverse_host_async(*enc_args(conts[1]), C.byref(j1)); lib.ctr_ans_decode_host_async(*dec_args(conts[0]), C.byref(j2))
    lib.ctr_host_job_wait(j1); lib.ctr_host_job_wait(j2)
both(); both()
sys.stderr.write("==== traced\n")
os.environ["X"] = "1"
both()
