"""Multi-GPU: shard independent streams over the ranks of one node, gather the compressed words.

Streams are independent coders (the reference's own many-small-streams pattern, tests/issue52.rs:38-53),
so encode and decode need no communication at all: rank r owns a contiguous block of streams.  The only
exchange step is the concatenation of the per-rank compressed containers: the ranks exchange their word
counts (8 bytes each), then one all-gather with per-rank sizes writes every rank's words straight into
their place in the dense global buffer (NCCL over NVLink; no padding, no re-packing), and the offset
tables are all-gathered and rebased with one fused tensor expression.  `torch.distributed` is the plumbing
(backend "nccl" on GPUs, "gloo" in the CPU tests).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition: rank r owns items [lo, hi); sizes differ by at most one."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_streams(sym_offsets: torch.Tensor, world_size: int, rank: int):
    """Contiguous layout: returns (stream_lo, stream_hi, symbol_lo, symbol_hi, local_offsets)."""
    k = sym_offsets.numel() - 1
    lo, hi = shard_bounds(k, world_size, rank)
    s_lo, s_hi = int(sym_offsets[lo].item()), int(sym_offsets[hi].item())
    return lo, hi, s_lo, s_hi, sym_offsets[lo:hi + 1] - s_lo


@dataclass
class GatheredContainer:
    words: torch.Tensor        # dense concatenation of all ranks' words (int32 bit patterns)
    offsets: torch.Tensor      # int64[K_total + 1], global word offsets of every stream
    stream_base: List[int]     # first global stream index of each rank
    word_base: List[int]       # first global word index of each rank


@dataclass
class PendingGather:
    """Handle between `all_gather_compressed_begin` and `all_gather_compressed_end`."""
    words: torch.Tensor
    offsets: torch.Tensor
    group: Optional[dist.ProcessGroup]
    metas: torch.Tensor                 # int64[world * 2] on the device: (total words, streams) of every rank
    metas_host: Optional[torch.Tensor]  # pinned copy (CUDA) -- valid once `ready` has fired
    ready: Optional["torch.cuda.Event"]
    stream_counts: Optional[Sequence[int]]


def all_gather_compressed_begin(words: torch.Tensor, offsets: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                                stream_counts: Optional[Sequence[int]] = None) -> PendingGather:
    """First half of the exchange: enqueues the 16-byte size exchange and its copy to pinned host memory on the
    current stream and returns at once.  The caller can enqueue independent work (the decode of its own shard, the
    next encode) before it calls `all_gather_compressed_end`, whose only host wait is then for these 16 bytes."""
    world = dist.get_world_size(group)
    dev = words.device
    k_local = offsets.numel() - 1
    meta = torch.stack([offsets[-1], torch.tensor(k_local, dtype=torch.int64, device=dev)])
    metas = torch.empty(world * 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(metas, meta, group=group)
    metas_host, ready = None, None
    if dev.type == "cuda":
        metas_host = torch.empty(world * 2, dtype=torch.int64).pin_memory()
        metas_host.copy_(metas, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record()
    return PendingGather(words, offsets, group, metas, metas_host, ready, stream_counts)


def all_gather_compressed_end(p: PendingGather) -> GatheredContainer:
    """Second half: waits (host) for the sizes, then enqueues the all-gather of the words with per-rank sizes --
    every rank's slice lands at its final place in the dense buffer -- and of the rebased offset tables."""
    words, offsets, group = p.words, p.offsets, p.group
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = words.device
    k_local = offsets.numel() - 1
    if p.ready is not None:
        p.ready.synchronize()
        metas_host = p.metas_host.view(world, 2)
    else:
        metas_host = p.metas.view(world, 2).cpu()
    lens = [int(x) for x in metas_host[:, 0]]
    ks = [int(x) for x in metas_host[:, 1]]
    if p.stream_counts is not None and [int(x) for x in p.stream_counts] != ks:
        raise ValueError("stream_counts does not match the ranks' containers")
    word_base, stream_base = [0], [0]
    for r in range(world):
        word_base.append(word_base[-1] + lens[r])
        stream_base.append(stream_base[-1] + ks[r])

    # words: every rank's slice lands at its final place in the dense buffer
    dense = torch.empty(max(word_base[-1], 1), dtype=words.dtype, device=dev)
    views = [dense[word_base[r]:word_base[r + 1]] for r in range(world)]
    if all(n == lens[0] for n in lens):
        dist.all_gather_into_tensor(dense[:world * lens[0]], words[:lens[0]], group=group)
    elif dist.get_backend(group) == "nccl":
        dist.all_gather(views, words[:lens[rank]], group=group)  # per-rank sizes: grouped NCCL broadcasts
    else:
        # backends without uneven all-gather (gloo, CPU tests): pad to the longest, then slice
        max_len = max(lens)
        send = words[:max_len] if words.numel() >= max_len else torch.cat(
            [words, torch.zeros(max_len - words.numel(), dtype=words.dtype, device=dev)])
        padded = torch.empty(world * max_len, dtype=words.dtype, device=dev)
        dist.all_gather_into_tensor(padded, send.contiguous(), group=group)
        for r in range(world):
            views[r].copy_(padded[r * max_len:r * max_len + lens[r]])

    # offsets: gather the (equal-length padded) tables, rebase by each rank's first word, drop the padding
    max_k = max(ks)
    if k_local == max_k:
        off_send = offsets
    else:
        off_send = torch.cat([offsets, offsets[-1:].expand(max_k - k_local)])
    off_all = torch.empty(world * (max_k + 1), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(off_all, off_send.contiguous(), group=group)
    view = off_all.view(world, max_k + 1)
    base = torch.tensor(word_base[:world], dtype=torch.int64).to(dev, non_blocking=True)
    view += base[:, None]  # rebase by each rank's first word (one kernel)
    if all(k == max_k for k in ks):
        # the padded table is [world][k + 1]; dropping each rank's last entry except the final one makes it dense
        g_off = torch.cat([view[:, :max_k].reshape(-1), view[world - 1, max_k:]])
    else:
        g_off = torch.cat([view[r, :ks[r]] for r in range(world)] + [view[world - 1, ks[world - 1]:ks[world - 1] + 1]])
    return GatheredContainer(dense, g_off, stream_base[:-1], word_base[:-1])


def all_gather_compressed(words: torch.Tensor, offsets: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                          stream_counts: Optional[Sequence[int]] = None) -> GatheredContainer:
    """Concatenates every rank's container (words[:offsets[-1]], offsets) into one global container that
    every rank holds.  `stream_counts[r]` = number of streams of rank r if the caller knows them (checked).
    `words` may be longer than offsets[-1] (capacity-sized buffers are fine; the slack is not sent).
    The one host synchronisation of the exchange is for the 16 bytes of sizes per rank; callers that have other
    work to enqueue meanwhile use the `_begin` / `_end` pair."""
    return all_gather_compressed_end(all_gather_compressed_begin(words, offsets, group, stream_counts))


class PeerGather:
    """The same exchange without any SM: every rank *pushes* its words and offsets straight into every peer's
    dense container with copy-engine transfers over NVLink, ordered by stream memory operations
    (cuStreamWriteValue32 / cuStreamWaitValue32 on flags in peer memory), so the gather runs concurrently with
    coder kernels that occupy every SM of the GPU.  (Any kernel-based collective -- an NCCL all-gather, even an
    8-byte one -- has to wait for a free SM, and the ANS encode kernel holds all registers of all SMs for its
    whole run: with NCCL the exchange and the coders serialise.)  The receive buffers are symmetric memory
    (`torch.distributed._symmetric_memory`: every rank's buffer is mapped into every rank's address space).

    Protocol of one gather (sequence number q, all on the caller's side stream, no kernel launches):
      begin:  my total -> meta[me] of every rank (8-byte copies); flag A[me] := q on every rank;
              wait until A[r] >= q for all r; meta -> pinned host memory; event
      end:    (host waits for the event: it needs the sizes to issue copies)  my words -> words[base_me ..] and my
              offsets -> offsets[K_me ..] of every rank; flag B[me] := q on every rank; wait until B[r] >= q for all r
      finish: (on the consumer's stream, after the gather's event) offsets of rank r += base_r   (one small kernel)

    One node, one process per GPU.  Receive buffers are allocated and exchanged once (`capacity_words` per rank
    and buffer); `n_buffers` of them are used round-robin, so that a container gathered in step i stays valid
    while step i+1 is being gathered."""

    def __init__(self, capacity_words: int, stream_counts: Sequence[int], group: Optional[dist.ProcessGroup] = None,
                 n_buffers: int = 2, dtype=torch.int32):
        import torch.distributed._symmetric_memory as symm_mem

        from . import _native as N
        self._lib = N.load()
        self.group = group
        pg = group if group is not None else dist.group.WORLD
        self.world = world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.ks = [int(x) for x in stream_counts]
        self.stream_base = [0]
        for kk in self.ks:
            self.stream_base.append(self.stream_base[-1] + kk)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        cap = torch.tensor([int(capacity_words)], dtype=torch.int64, device=dev)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=group)  # symmetric buffers: same size on every rank
        self.cap = int(cap.item())
        self.n_buffers = n_buffers
        n_words, n_off = world * self.cap, self.stream_base[-1] + 1
        self.dense, self.g_off, self._handles = [], [], []
        self._dense_ptrs, self._off_ptrs = [], []
        self.peer_dense = [[None] * n_buffers for _ in range(world)]
        self.peer_off = [[None] * n_buffers for _ in range(world)]
        for b in range(n_buffers):
            d = symm_mem.empty(n_words, dtype=dtype, device=dev)
            o = symm_mem.empty(n_off, dtype=torch.int64, device=dev)
            hd, ho = symm_mem.rendezvous(d, pg), symm_mem.rendezvous(o, pg)
            self._handles += [hd, ho]
            self.dense.append(d)
            self.g_off.append(o)
            for r in range(world):
                self.peer_dense[r][b] = d if r == self.rank else hd.get_buffer(r, (n_words,), dtype, 0)
                self.peer_off[r][b] = o if r == self.rank else ho.get_buffer(r, (n_off,), torch.int64, 0)
            self._dense_ptrs.append((ctypes.c_void_p * world)(*[int(x) for x in hd.buffer_ptrs]))
            self._off_ptrs.append((ctypes.c_void_p * world)(*[int(x) for x in ho.buffer_ptrs]))
        # sizes and flags: meta int64[world]; flags int32[2 * world] (A then B), zero-initialised
        self.meta = symm_mem.empty(world, dtype=torch.int64, device=dev)
        self.flags = symm_mem.empty(2 * world, dtype=torch.int32, device=dev)
        self.meta.zero_()
        self.flags.zero_()
        torch.cuda.synchronize()
        hm, hf = symm_mem.rendezvous(self.meta, pg), symm_mem.rendezvous(self.flags, pg)
        self._handles += [hm, hf]
        self.peer_meta = [self.meta if r == self.rank else hm.get_buffer(r, (world,), torch.int64, 0) for r in range(world)]
        self.flag_ptrs = [int(x) for x in hf.buffer_ptrs]  # flags of rank r, mapped here
        self._meta_ptrs = (ctypes.c_void_p * world)(*[int(x) for x in hm.buffer_ptrs])
        # flag (slot, me) in every rank's array; flags (slot, r) in mine
        self._signal_ptrs = [(ctypes.c_void_p * world)(*[self.flag_ptrs[d] + 4 * (slot * world + self.rank) for d in range(world)])
                             for slot in range(2)]
        self._wait_ptrs = [(ctypes.c_void_p * world)(*[self.flag_ptrs[self.rank] + 4 * (slot * world + r) for r in range(world)])
                           for slot in range(2)]
        self.meta_host = torch.empty(world, dtype=torch.int64).pin_memory()
        counts = torch.tensor(self.ks, dtype=torch.int64)
        counts[-1] += 1  # the final entry of the table belongs to the last rank
        self._owner = torch.repeat_interleave(torch.arange(world), counts).to(dev)  # rank that supplies each entry
        self._seq = 0
        self._turn = 0
        # self-test of the stream memory operations on peer memory (value 0: leaves the protocol untouched); raises
        # here, where callers can still fall back to the NCCL path, rather than in the first gather
        self._signal_all(0, 0)
        self._wait_all(0, 0)
        torch.cuda.synchronize()
        dist.barrier(group=group)
        torch.cuda.synchronize()

    # -- stream memory operations on the current stream ------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc:
            raise RuntimeError("PeerGather: " + self._lib.ctr_status_string(rc).decode() + " " + self._lib.ctr_last_cuda_error().decode())

    def _signal_all(self, slot: int, value: int) -> None:
        self._check(self._lib.ctr_stream_write_value32_many(self._signal_ptrs[slot], self.world, (self.rank + 1) % self.world, value,
                                                           torch.cuda.current_stream().cuda_stream))

    def _wait_all(self, slot: int, value: int) -> None:
        self._check(self._lib.ctr_stream_wait_value32_many(self._wait_ptrs[slot], self.world, value,
                                                          torch.cuda.current_stream().cuda_stream))

    def gather_begin(self, words: torch.Tensor, offsets: torch.Tensor) -> PendingGather:
        self._seq += 1
        q = self._seq
        self._check(self._lib.ctr_peer_push(self._meta_ptrs, self.world, (self.rank + 1) % self.world, 8 * self.rank,
                                            offsets[-1:].data_ptr(), 8, torch.cuda.current_stream().cuda_stream))
        self._signal_all(0, q)
        self._wait_all(0, q)
        self.meta_host.copy_(self.meta, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record()
        return PendingGather(words, offsets, self.group, self.meta, self.meta_host, ready, self.ks)

    def gather_end(self, p: PendingGather) -> GatheredContainer:
        world, rank = self.world, self.rank
        p.ready.synchronize()
        lens = [int(x) for x in p.metas_host]
        if max(lens) > self.cap:
            raise MemoryError("PeerGather: a rank's container exceeds the receive capacity")
        word_base = [0]
        for n in lens:
            word_base.append(word_base[-1] + n)
        b = self._turn
        self._turn = (self._turn + 1) % self.n_buffers
        wb, n = word_base[rank], lens[rank]
        sb, k = self.stream_base[rank], self.ks[rank]
        n_off = k + 1 if rank == world - 1 else k  # the last rank also supplies the final entry
        stream = torch.cuda.current_stream().cuda_stream
        first = (rank + 1) % world  # start with my right neighbour: in every round the destinations form a permutation
        self._check(self._lib.ctr_peer_push(self._dense_ptrs[b], world, first, 4 * wb, p.words.data_ptr(), 4 * n, stream))
        self._check(self._lib.ctr_peer_push(self._off_ptrs[b], world, first, 8 * sb, p.offsets.data_ptr(), 8 * n_off, stream))
        self._signal_all(1, self._seq)
        self._wait_all(1, self._seq)
        total = word_base[-1]
        gc = GatheredContainer(self.dense[b][:max(total, 1)], self.g_off[b], self.stream_base[:-1], word_base[:-1])
        gc._rebase = torch.tensor(word_base[:world], dtype=torch.int64)
        return gc

    def finish(self, gc: GatheredContainer) -> GatheredContainer:
        """Rebases the gathered offset table (rank r's entries += first word of rank r) on the current stream; call
        it once, after waiting for the event recorded behind `gather_end`.  Until then `gc.offsets` holds the
        ranks' local offsets."""
        base = getattr(gc, "_rebase", None)
        if base is not None:
            gc.offsets += base.pin_memory().to(gc.offsets.device, non_blocking=True)[self._owner]
            gc._rebase = None
        return gc
