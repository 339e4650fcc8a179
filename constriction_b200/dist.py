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


class SlotGather:
    """The exchange without any SM and without a cross-rank host wait (C ABI: ctr_gather_*, csrc/gather.cu).

    Every rank owns receive buffers in symmetric memory (`torch.distributed._symmetric_memory`: mapped into every
    rank's address space over NVLink), divided into fixed slots, one per source rank: slot (turn mod n_buffers, r)
    holds rank r's container of that turn exactly as its encoder wrote it.  A rank encodes straight into its own
    slot; a worker thread of the library waits for that encode and pushes the used part of the slot into every
    peer's buffer with copy-engine transfers, ordered by stream memory operations on flags in peer memory
    (cuStreamWriteValue32 / cuStreamWaitValue32), so the gather runs while coder kernels occupy every SM and no
    Python thread ever blocks on another rank.  One node, one process per GPU.

        turn = sg.begin_turn()                                  # my slot of this turn
        comp = bc.ans_encode(syms, model, ..., out=turn.out)    # encode straight into it
        sg.push(turn, comp.n_streams)                           # returns at once
        ...
        sg.wait(turn)                                           # current stream: all peers' slots have arrived
        other = sg.shard(turn, r, n_streams, n_symbols, "ans")  # rank r's container (views, no copy)
        bc.ans_decode(other, model)
        sg.release(turn)                                        # peers may overwrite this buffer (turn + n_buffers)
    """

    @dataclass
    class Turn:
        number: int
        out: "object"  # batch.Compressed whose words / offsets are my slot (pass as `out=` to the encoders)

    def __init__(self, slot_words: int, slot_streams: int, group: Optional[dist.ProcessGroup] = None, n_buffers: int = 2):
        import torch.distributed._symmetric_memory as symm_mem

        from . import _native as N
        self._N = N
        self._lib = N.load()
        self.group = group
        pg = group if group is not None else dist.group.WORLD
        self.world = world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        sizes = torch.tensor([int(slot_words), int(slot_streams)], dtype=torch.int64, device=dev)
        dist.all_reduce(sizes, op=dist.ReduceOp.MAX, group=group)  # symmetric buffers: same geometry on every rank
        self.slot_words = (int(sizes[0].item()) + 63) // 64 * 64
        self.slot_streams = int(sizes[1].item())
        self.n_buffers = n_buffers
        n_words = n_buffers * world * self.slot_words
        n_off = n_buffers * world * (self.slot_streams + 1)
        self.words = symm_mem.empty(n_words, dtype=torch.int32, device=dev)
        self.offsets = symm_mem.empty(n_off, dtype=torch.int64, device=dev)
        self.flags = symm_mem.empty(3 * n_buffers * world, dtype=torch.int32, device=dev)
        self.flags.zero_()
        torch.cuda.synchronize()
        self._handles = [symm_mem.rendezvous(t, pg) for t in (self.words, self.offsets, self.flags)]
        arr = lambda h: (ctypes.c_void_p * world)(*[int(x) for x in h.buffer_ptrs])
        self._h = ctypes.c_void_p()
        self._views = {}
        dist.barrier(group=group)  # every rank has zeroed its flags
        N.raise_for(self._lib.ctr_gather_create(world, self.rank, n_buffers, self.slot_words, self.slot_streams,
                                                arr(self._handles[0]), arr(self._handles[1]), arr(self._handles[2]),
                                                ctypes.byref(self._h)))

    def close(self) -> None:
        if self._h:
            self._lib.ctr_gather_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _slot_views(self, turn: int, src: int):
        b = turn % self.n_buffers
        key = (b, src)
        v = self._views.get(key)
        if v is None:
            w0 = (b * self.world + src) * self.slot_words
            o0 = (b * self.world + src) * (self.slot_streams + 1)
            v = self._views[key] = (self.words[w0:w0 + self.slot_words], self.offsets[o0:o0 + self.slot_streams + 1], {})
        return v

    def _offsets_view(self, turn: int, src: int, n_streams: int):
        w, o, cache = self._slot_views(turn, src)
        ov = cache.get(n_streams)
        if ov is None:
            ov = cache[n_streams] = o[:n_streams + 1]
        return w, ov

    def begin_turn(self, n_streams: int, n_symbols: int = 0, coder: str = "ans") -> "SlotGather.Turn":
        from .batch import Compressed
        q = ctypes.c_uint32()
        self._N.raise_for(self._lib.ctr_gather_begin_turn(self._h, ctypes.byref(q), None, None, None))
        w, o = self._offsets_view(q.value, self.rank, n_streams)
        return SlotGather.Turn(q.value, Compressed(w, o, n_streams, n_symbols, coder))

    def push(self, turn: "SlotGather.Turn", n_streams: int) -> None:
        """Behind the encode on the current stream; returns at once."""
        self._N.raise_for(self._lib.ctr_gather_push(self._h, turn.number, n_streams, torch.cuda.current_stream().cuda_stream))

    def wait(self, turn: "SlotGather.Turn") -> None:
        self._N.raise_for(self._lib.ctr_gather_wait(self._h, turn.number, torch.cuda.current_stream().cuda_stream))

    def release(self, turn: "SlotGather.Turn") -> None:
        self._N.raise_for(self._lib.ctr_gather_release(self._h, turn.number, torch.cuda.current_stream().cuda_stream))

    def sync(self) -> None:
        self._N.raise_for(self._lib.ctr_gather_sync(self._h))

    def shard(self, turn: "SlotGather.Turn", src_rank: int, n_streams: int, n_symbols: int, coder: str = "ans", sym_offsets=None):
        """Rank `src_rank`'s container of this turn as a batch.Compressed (views into my receive buffer)."""
        from .batch import Compressed
        w, o = self._offsets_view(turn.number, src_rank, n_streams)
        return Compressed(w, o, n_streams, n_symbols, coder, sym_offsets)


class NcclComm:
    """A raw NCCL communicator over the ranks of a torch.distributed group (ctypes on the NCCL library PyTorch
    ships), for the C ABI's ctr_gather_compressed_nccl: what a Rust / C host passes as its own ncclComm_t."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        import glob
        import os
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        lib = None
        cands = []
        try:
            import nvidia.nccl
            for base in nvidia.nccl.__path__:  # namespace package
                cands += glob.glob(os.path.join(base, "lib", "libnccl.so*"))
        except Exception:
            pass
        for path in cands + ["libnccl.so.2"]:
            try:
                lib = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
                os.environ.setdefault("CTR_NCCL_LIB", path)
                break
            except OSError:
                continue
        if lib is None:
            raise RuntimeError("libnccl not found")
        self._nccl = lib
        uid = (ctypes.c_byte * 128)()
        if self.rank == 0:
            rc = lib.ncclGetUniqueId(ctypes.byref(uid))
            if rc:
                raise RuntimeError(f"ncclGetUniqueId failed ({rc})")
        box = [bytes(uid)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        uid = (ctypes.c_byte * 128).from_buffer_copy(box[0])

        class _Uid(ctypes.Structure):
            _fields_ = [("internal", ctypes.c_byte * 128)]

        u = _Uid()
        ctypes.memmove(ctypes.byref(u), uid, 128)
        self.comm = ctypes.c_void_p()
        lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _Uid, ctypes.c_int]
        rc = lib.ncclCommInitRank(ctypes.byref(self.comm), self.world, u, self.rank)
        if rc:
            raise RuntimeError(f"ncclCommInitRank failed ({rc})")

    def close(self):
        if self.comm:
            self._nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
            self._nccl.ncclCommDestroy(self.comm)
            self.comm = None


def gather_compressed_nccl(comm: NcclComm, words: torch.Tensor, offsets: torch.Tensor, slot_words: int, slot_streams: int):
    """ctr_gather_compressed_nccl on the current stream: returns (words int32[world][slot_words],
    offsets int64[world][slot_streams + 1], meta int64[world][2] on the host = {total words, streams} per rank)."""
    from . import _native as N
    lib = N.load()
    dev = words.device
    world = comm.world
    k = offsets.numel() - 1
    w_out = torch.empty((world, slot_words), dtype=torch.int32, device=dev)
    o_out = torch.zeros((world, slot_streams + 1), dtype=torch.int64, device=dev)
    meta_dev = torch.empty(2 * world + 2, dtype=torch.int64, device=dev)
    meta_host = torch.empty((world, 2), dtype=torch.int64).pin_memory()
    N.raise_for(lib.ctr_gather_compressed_nccl(comm.comm, world, comm.rank, words.data_ptr(), offsets.data_ptr(), k, slot_words,
                                               slot_streams, w_out.data_ptr(), o_out.data_ptr(), meta_dev.data_ptr(),
                                               meta_host.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return w_out, o_out, meta_host
