#!/usr/bin/env python
"""Isolated timing of the exchange on N GPUs (no coder kernels running): NCCL all-gather vs the slotted peer pushes.
torchrun --nproc-per-node N tools/gather_diag.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from constriction_b200 import dist as D  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
words_n, k = 16_753_128, 151_552  # the container of the headline workload: 67 MB of words, 1.2 MB of offsets


def timed(name, fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{name:60s} {float(t.item()):8.3f} ms", flush=True)


src = torch.randint(0, 2**31 - 1, (words_n,), dtype=torch.int32, device="cuda")
dense = torch.empty(world * words_n, dtype=torch.int32, device="cuda")
timed("NCCL all_gather_into_tensor (67 MB per rank)", lambda: dist.all_gather_into_tensor(dense, src))

sg = D.SlotGather(words_n + 1024, k)
off = torch.zeros(k + 1, dtype=torch.int64, device="cuda")
off[-1] = words_n
pending = []


def turn_lag(lag):
    def f():
        t = sg.begin_turn(k, 0, "ans")
        t.out.words[:words_n].copy_(src)  # stands in for the encoder writing its slot
        t.out.offsets.copy_(off)
        sg.push(t, k)
        pending.append(t)
        if len(pending) > lag:
            p = pending.pop(0)
            sg.wait(p)
            sg.release(p)
    return f


timed(f"SlotGather, push streams {os.environ.get('CTR_PUSH_STREAMS', '1')}, consumed at once", turn_lag(0))
while pending:
    p = pending.pop(0); sg.wait(p); sg.release(p)
timed(f"SlotGather, push streams {os.environ.get('CTR_PUSH_STREAMS', '1')}, consumed one turn later", turn_lag(1))
while pending:
    p = pending.pop(0); sg.wait(p); sg.release(p)
sg.sync()
torch.cuda.synchronize()
dist.barrier()
sg.close()
dist.destroy_process_group()
