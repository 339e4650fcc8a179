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


def all_gather_compressed(words: torch.Tensor, offsets: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                          stream_counts: Optional[Sequence[int]] = None) -> GatheredContainer:
    """Concatenates every rank's container (words[:offsets[-1]], offsets) into one global container that
    every rank holds.  `stream_counts[r]` = number of streams of rank r if the caller knows them (e.g. from
    `shard_bounds`); otherwise they are exchanged together with the word counts.
    `words` may be longer than offsets[-1] (capacity-sized buffers are fine; the slack is not sent)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = words.device
    k_local = offsets.numel() - 1

    # 1. word counts (and stream counts if unknown): the one host synchronisation of the exchange
    if stream_counts is not None:
        totals = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(totals, offsets[-1:], group=group)
        lens = [int(x) for x in totals.cpu()]
        ks = [int(x) for x in stream_counts]
    else:
        meta = torch.stack([offsets[-1], torch.tensor(k_local, dtype=torch.int64, device=dev)])
        metas = torch.empty(world * 2, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(metas, meta, group=group)
        metas_host = metas.view(world, 2).cpu()
        lens = [int(x) for x in metas_host[:, 0]]
        ks = [int(x) for x in metas_host[:, 1]]
        totals = metas.view(world, 2)[:, 0]
    word_base, stream_base = [0], [0]
    for r in range(world):
        word_base.append(word_base[-1] + lens[r])
        stream_base.append(stream_base[-1] + ks[r])

    # 2. words: every rank's slice lands at its final place in the dense buffer
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

    # 3. offsets: gather the (equal-length padded) tables, rebase by each rank's first word, drop the padding
    max_k = max(ks)
    if k_local == max_k:
        off_send = offsets
    else:
        off_send = torch.cat([offsets, offsets[-1:].expand(max_k - k_local)])
    off_all = torch.empty(world * (max_k + 1), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(off_all, off_send.contiguous(), group=group)
    view = off_all.view(world, max_k + 1)
    for r in range(1, world):  # rebase in place by each rank's first word (host-known after step 1)
        view[r].add_(word_base[r])
    if all(k == max_k for k in ks):
        # the padded table is [world][k + 1]; dropping each rank's last entry except the final one makes it dense
        g_off = torch.cat([view[:, :max_k].reshape(-1), view[world - 1, max_k:]])
    else:
        g_off = torch.cat([view[r, :ks[r]] for r in range(world)] + [view[world - 1, ks[world - 1]:ks[world - 1] + 1]])
    return GatheredContainer(dense, g_off, stream_base[:-1], word_base[:-1])
