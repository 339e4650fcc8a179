"""Multi-GPU: shard independent streams over the ranks of one node, gather the compressed words.

Streams are independent coders (the reference's own many-small-streams pattern, tests/issue52.rs:38-53),
so encode and decode need no communication at all: rank r owns a contiguous block of streams.  The only
exchange step is the concatenation of the per-rank compressed containers, done with one NCCL
all-gather of the (max-padded) word buffers over NVLink plus an all-gather of the offset tables;
`torch.distributed` is the plumbing (backend "nccl" on GPUs, "gloo" in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

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


def all_gather_compressed(words: torch.Tensor, offsets: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                          total_words: Optional[int] = None) -> GatheredContainer:
    """Concatenates every rank's container (words[:offsets[-1]], offsets) into one global container that
    every rank holds.  Two collectives on the data path: lengths (tiny) and the padded words."""
    world = dist.get_world_size(group)
    dev = words.device
    n_local_streams = offsets.numel() - 1
    if total_words is None:
        total_words = int(offsets[-1].item())
    meta = torch.tensor([total_words, n_local_streams], dtype=torch.int64, device=dev)
    metas = torch.empty(world * 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(metas, meta, group=group)
    metas = metas.view(world, 2).cpu()
    lens = [int(x) for x in metas[:, 0]]
    ks = [int(x) for x in metas[:, 1]]
    max_len, max_k = max(max(lens), 1), max(ks)

    send = words[:max_len] if words.numel() >= max_len else torch.cat(
        [words, torch.zeros(max_len - words.numel(), dtype=words.dtype, device=dev)])
    gathered = torch.empty(world * max_len, dtype=words.dtype, device=dev)
    dist.all_gather_into_tensor(gathered, send.contiguous(), group=group)

    off_send = offsets if n_local_streams == max_k else torch.cat(
        [offsets, offsets[-1:].expand(max_k - n_local_streams)])
    off_all = torch.empty(world * (max_k + 1), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(off_all, off_send.contiguous(), group=group)
    off_all = off_all.view(world, max_k + 1)

    word_base, stream_base = [0], [0]
    for r in range(world):
        word_base.append(word_base[-1] + lens[r])
        stream_base.append(stream_base[-1] + ks[r])
    dense = torch.empty(max(word_base[-1], 1), dtype=words.dtype, device=dev)
    g_off = torch.empty(stream_base[-1] + 1, dtype=torch.int64, device=dev)
    for r in range(world):
        dense[word_base[r]:word_base[r + 1]] = gathered[r * max_len:r * max_len + lens[r]]
        g_off[stream_base[r]:stream_base[r] + ks[r]] = off_all[r, :ks[r]] + word_base[r]
    g_off[-1] = word_base[-1]
    return GatheredContainer(dense, g_off, stream_base[:-1], word_base[:-1])
