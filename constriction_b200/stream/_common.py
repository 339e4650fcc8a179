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
