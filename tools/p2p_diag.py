"""Diagnostic (2 GPUs, one process): is cudaMemcpy between the GPUs peer-to-peer, and how fast?"""
import time
import torch
print("can access peer 0->1:", torch.cuda.can_device_access_peer(0, 1), " 1->0:", torch.cuda.can_device_access_peer(1, 0))
n = 64 * 1024 * 1024 // 4
a = torch.randint(0, 100, (n,), dtype=torch.int32, device="cuda:0")
b = torch.empty(n, dtype=torch.int32, device="cuda:1")
for name, fn in (("b.copy_(a) [cuda:0 -> cuda:1], current device 0", lambda: b.copy_(a, non_blocking=True)),):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    dt = (time.perf_counter() - t0) / 10
    print(f"{name}: {dt*1e6:.1f} us  {n*4/dt/1e9:.1f} GB/s")
assert torch.equal(a.cpu(), b.cpu())
import subprocess
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
