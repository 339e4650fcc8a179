from __future__ import annotations

import numpy as np
import torch

from .. import batch as B

_CODER = None


def coder() -> B.BatchCoder:
    global _CODER
    if _CODER is None:
        _CODER = B.BatchCoder()
    return _CODER


def is_scalar(x) -> bool:
    return isinstance(x, (int, np.integer)) and not isinstance(x, bool)


def symbols_array(symbols) -> np.ndarray:
    a = np.asarray(symbols)
    if a.ndim != 1 or a.dtype != np.int32:
        raise TypeError("symbols must be a rank-1 numpy array with dtype=np.int32")
    return np.ascontiguousarray(a)


def to_dev_i32(a: np.ndarray) -> torch.Tensor:
    """numpy int32 / uint32 -> CUDA int32 tensor (bit pattern preserved)."""
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32).copy()).to(coder().device)


def to_dev_u64(values) -> torch.Tensor:
    a = np.array([int(v) & 0xFFFFFFFFFFFFFFFF for v in values], dtype=np.uint64)
    return torch.from_numpy(a.view(np.int64).copy()).to(coder().device)


def from_dev_u64(t: torch.Tensor) -> list:
    return [int(v) for v in t.cpu().numpy().view(np.uint64)]


def first_impossible(symbols: np.ndarray, table, per_symbol: bool, reverse: bool) -> int:
    """Number of symbols, in CODING order, that precede the first symbol without probability under its model.  The
    reference codes symbol by symbol (stream/mod.rs:592-607), so when `ImpossibleSymbol` is raised the symbols before
    the offending one stay encoded; the batched kernels reject the whole call, and the mirror then encodes that
    prefix again to leave the coder in the reference's state."""
    from ..batch import GaussianParams
    order = symbols[::-1] if reverse else symbols
    n = order.size
    if isinstance(table, GaussianParams):
        bad = (order < table.min_symbol) | (order > table.max_symbol)
    else:
        cdf = table.cdf().astype(np.int64)
        idx = order.astype(np.int64) - table.min_symbol
        inside = (idx >= 0) & (idx < table.alphabet)
        safe = np.where(inside, idx, 0)
        rows = (np.arange(n)[::-1] if reverse else np.arange(n)) if per_symbol else np.zeros(n, dtype=np.int64)
        bad = ~inside | (cdf[rows, safe + 1] == cdf[rows, safe])
    hits = np.flatnonzero(bad)
    return int(hits[0]) if hits.size else n
