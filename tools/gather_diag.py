"""Diagnostic: where does the time of dist.all_gather_compressed go?  (torchrun, N >= 2)"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from constriction_b200 import dist as D

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
k, total = 151552, 16_753_000 + 1000 * rank
words = torch.randint(-2**31, 2**31 - 1, (total + 5_000_000,), dtype=torch.int32, device="cuda")
offsets = (torch.arange(k + 1, device="cuda", dtype=torch.int64) * total) // k


def timeit(name, fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    if rank == 0:
        print(f"{name:45s} {dt * 1e6:9.1f} us", flush=True)


meta = torch.zeros(2, dtype=torch.int64, device="cuda")
metas = torch.empty(world * 2, dtype=torch.int64, device="cuda")
timeit("meta all_gather_into_tensor (16 B)", lambda: dist.all_gather_into_tensor(metas, meta))
timeit("meta all_gather + .cpu()", lambda: (dist.all_gather_into_tensor(metas, meta), metas.cpu()))
eq = torch.empty(world * total, dtype=torch.int32, device="cuda")
timeit("words all_gather_into_tensor (equal, 67 MB)", lambda: dist.all_gather_into_tensor(eq, words[:total]))
lens = [16_753_000 + 1000 * r for r in range(world)]
base = [sum(lens[:r]) for r in range(world + 1)]
dense = torch.empty(base[-1], dtype=torch.int32, device="cuda")
views = [dense[base[r]:base[r + 1]] for r in range(world)]
timeit("words all_gather (uneven list)", lambda: dist.all_gather(views, words[:lens[rank]]))
off_all = torch.empty(world * (k + 1), dtype=torch.int64, device="cuda")
timeit("offsets all_gather_into_tensor (1.2 MB)", lambda: dist.all_gather_into_tensor(off_all, offsets))
timeit("full all_gather_compressed", lambda: D.all_gather_compressed(words, offsets, stream_counts=[k] * world))
pg = D.PeerGather(words.numel(), [k] * world)
peer = (rank + 1) % world
n = lens[rank]
timeit("raw push: peer_dense[peer][0][:n].copy_(words[:n])", lambda: pg.peer_dense[peer][0][:n].copy_(words[:n], non_blocking=True))
timeit("raw self copy: dense[0][:n].copy_(words[:n])", lambda: pg.dense[0][:n].copy_(words[:n], non_blocking=True))
timeit("PeerGather begin+end+finish", lambda: pg.finish(pg.gather_end(pg.gather_begin(words, offsets))))
gc = pg.finish(pg.gather_end(pg.gather_begin(words, offsets)))
ref = D.all_gather_compressed(words, offsets, stream_counts=[k] * world)
torch.cuda.synchronize()
assert torch.equal(gc.offsets, ref.offsets) and torch.equal(gc.words[:ref.words.numel()], ref.words), "PeerGather != NCCL gather"
if rank == 0:
    print("PeerGather result equals the NCCL all-gather", flush=True)
dist.destroy_process_group()
