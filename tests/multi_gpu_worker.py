"""Worker of tests/test_gpu_multi.py (one process per GPU, launched with torch.distributed.run).

Every rank encodes its own shard of streams (ANS on the interleaved layout, range coder on ragged contiguous
streams), the containers are exchanged three ways -- SlotGather (ctr_gather_*: copy-engine pushes into peer-mapped
slots), ctr_gather_compressed_nccl (raw communicator) and all_gather_compressed (torch.distributed) -- and EVERY rank
decodes EVERY OTHER rank's shard out of the gathered container.  Words are compared with the CPU oracle run on the
other rank's symbols (regenerated here from that rank's seed), decoded symbols with those symbols.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MODEL = (-50, 50, 3.2, 9.6)


def shard_data(rank, turn):
    """The shard rank `rank` encodes in turn `turn`: (ans symbols, ans streams, range symbols, range offsets)."""
    rng = np.random.default_rng(1000 * turn + rank)
    k_ans = 96 + 32 * rank  # ranks own different numbers of streams
    n_ans = k_ans * (40 + 7 * turn) + 13 * rank + 5  # ragged last row
    ans = np.clip(np.rint(rng.normal(MODEL[2], MODEL[3], size=n_ans)), MODEL[0], MODEL[1]).astype(np.int32)
    k_rng = 40 + 5 * rank
    lens = rng.integers(0, 300, size=k_rng)
    lens[rng.integers(0, k_rng)] = 0  # an empty stream somewhere
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    rsyms = np.clip(np.rint(rng.normal(MODEL[2], MODEL[3], size=int(off[-1]))), MODEL[0], MODEL[1]).astype(np.int32)
    return ans, k_ans, rsyms, off


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    from constriction_b200 import batch as B
    from constriction_b200 import dist as D
    from oracle import refapi as O

    model = B.ModelTable.quantized_gaussian(MODEL[0], MODEL[1], [MODEL[2]], [MODEL[3]])
    cdf = O.qgauss_cdf(*MODEL)
    bc = B.BatchCoder()
    turns = 5  # more than n_buffers: slots are reused, the release flags matter
    data = [[shard_data(r, t) for r in range(world)] for t in range(turns)]
    max_ans = max(d[1] for row in data for d in row)
    max_rng = max(len(d[3]) - 1 for row in data for d in row)
    slot_streams = max(max_ans, max_rng)
    slot_words = max(max(d[0].size, d[2].size) for row in data for d in row) + 4 * slot_streams + 64

    def check_shard(kind, comp, r, t, label):
        """`comp` = rank r's container of turn t as it arrived here: words == oracle, decode == symbols"""
        ans, k_ans, rsyms, off = data[t][r]
        if kind == "ans":
            want_words, want_off = O.multi_ans_encode(ans, k_ans, cdf, MODEL[0])
            out = bc.ans_decode(comp, model)
            syms = ans
        else:
            want_words, want_off = O.multi_range_encode(rsyms, len(off) - 1, cdf, MODEL[0], sym_offsets=off.astype(np.uint64))
            out = bc.range_decode(comp, model)
            syms = rsyms
        bc.check()
        got_off = comp.offsets.cpu().numpy().astype(np.uint64)
        got_words = comp.words[:int(got_off[-1])].cpu().numpy().view(np.uint32)
        assert np.array_equal(got_off, want_off), f"{label}: offsets of rank {r} turn {t} differ from the oracle"
        assert np.array_equal(got_words, want_words), f"{label}: words of rank {r} turn {t} differ from the oracle"
        assert np.array_equal(out.cpu().numpy(), syms), f"{label}: decode of rank {r} turn {t} differs"

    # ---- SlotGather: two independent gathers (ANS and range containers), pipelined over the turns ------------------
    for kind in ("ans", "range"):
        sg = D.SlotGather(slot_words, slot_streams)
        pending = []
        for t in range(turns):
            ans, k_ans, rsyms, off = data[t][rank]
            if kind == "ans":
                turn = sg.begin_turn(k_ans, ans.size, "ans")
                bc.ans_encode(torch.from_numpy(ans).cuda(), model, n_streams=k_ans, out=turn.out)
                sg.push(turn, k_ans)
            else:
                k = len(off) - 1
                turn = sg.begin_turn(k, rsyms.size, "range")
                bc.range_encode(torch.from_numpy(rsyms).cuda(), model, sym_offsets=torch.from_numpy(off).cuda(), out=turn.out)
                sg.push(turn, k)
            pending.append((t, turn))
            if len(pending) == 2:  # consume turn t-1 while turn t is in flight
                consume(sg, kind, pending.pop(0), data, world, rank, check_shard)
        while pending:
            consume(sg, kind, pending.pop(0), data, world, rank, check_shard)
        sg.sync()
        torch.cuda.synchronize()
        dist.barrier()
        sg.close()

    # ---- NCCL through the C ABI (raw communicator) and through torch.distributed -------------------------------------
    comm = D.NcclComm()
    for t in range(2):
        ans, k_ans, rsyms, off = data[t][rank]
        comp = bc.ans_encode(torch.from_numpy(ans).cuda(), model, n_streams=k_ans)
        w, o, meta = D.gather_compressed_nccl(comm, comp.words, comp.offsets, slot_words // 4 * 4, slot_streams)
        torch.cuda.synchronize()
        for r in range(world):
            kr, nr = data[t][r][1], data[t][r][0].size
            assert int(meta[r, 1]) == kr
            check_shard("ans", B.Compressed(w[r], o[r, :kr + 1].contiguous(), kr, nr, "ans"), r, t, "ctr_gather_compressed_nccl")
        g = D.all_gather_compressed(comp.words, comp.offsets)
        for r in range(world):
            kr, nr = data[t][r][1], data[t][r][0].size
            lo = g.stream_base[r]
            base = g.word_base[r]
            # a rank's shard inside the dense container: rebase its offsets to a 16-byte aligned copy of its words
            offs = (g.offsets[lo:lo + kr + 1] - base).contiguous()
            words = g.words[base:base + int(offs[-1])].clone()
            check_shard("ans", B.Compressed(words, offs, kr, nr, "ans"), r, t, "all_gather_compressed")
    comm.close()
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("MULTI_GPU_OK", flush=True)
    dist.destroy_process_group()


def consume(sg, kind, item, data, world, rank, check_shard):
    t, turn = item
    sg.wait(turn)
    for r in range(world):  # every rank's shard, the others' first
        src = (rank + 1 + r) % world
        if kind == "ans":
            k, n = data[t][src][1], data[t][src][0].size
            comp = sg.shard(turn, src, k, n, "ans")
        else:
            off = data[t][src][3]
            comp = sg.shard(turn, src, len(off) - 1, data[t][src][2].size, "range", sym_offsets=torch.from_numpy(off).cuda())
        check_shard(kind, comp, src, t, "SlotGather")
    sg.release(turn)


if __name__ == "__main__":
    main()
