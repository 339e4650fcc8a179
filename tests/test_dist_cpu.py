"""Host-side multi-GPU logic on CPU: world_size-2 `gloo` run of the shard partition and of the
container all-gather (constriction_b200/dist.py).  No compute kernels involved."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from constriction_b200 import dist as D
        rng = np.random.default_rng(100 + rank)
        k_local = 5 + 3 * rank                      # ragged: ranks own different numbers of streams
        lens = rng.integers(0, 50, size=k_local)
        offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64))
        total = int(offsets[-1])
        payload = torch.from_numpy(rng.integers(-2**31, 2**31 - 1, size=total + 17).astype(np.int32))  # + slack capacity
        g = D.all_gather_compressed(payload, offsets)
        # every rank must hold the same global container
        q.put((rank, g.words.numpy().copy(), g.offsets.numpy().copy(), g.stream_base, g.word_base,
               payload[:total].numpy().copy(), offsets.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_partition():
    from constriction_b200.dist import shard_bounds
    for n in (0, 1, 7, 8, 8192, 12_288):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_streams_offsets():
    from constriction_b200.dist import shard_streams
    off = torch.tensor([0, 4, 4, 10, 11, 20, 26, 30], dtype=torch.int64)
    seen = []
    for r in range(3):
        lo, hi, s_lo, s_hi, local = shard_streams(off, 3, r)
        assert local[0] == 0 and int(local[-1]) == s_hi - s_lo
        assert torch.equal(local, off[lo:hi + 1] - s_lo)
        seen.append((lo, hi))
    assert seen[0][0] == 0 and seen[-1][1] == 7


@pytest.mark.timeout(120)
def test_all_gather_compressed_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=100) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    words0, off0 = results[0][1], results[0][2]
    for r in results:
        assert np.array_equal(r[1][: off0[-1]], words0[: off0[-1]]) and np.array_equal(r[2], off0)
    # the global container is the concatenation of the per-rank containers
    want_words = np.concatenate([r[5] for r in results])
    assert np.array_equal(words0[: want_words.size], want_words)
    base = 0
    stream = 0
    for r in results:
        loc = r[6]
        k = loc.size - 1
        assert np.array_equal(off0[stream:stream + k + 1], loc + base)
        assert results[0][3][r[0]] == stream and results[0][4][r[0]] == base
        base += int(loc[-1])
        stream += k
