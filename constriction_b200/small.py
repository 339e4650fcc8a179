"""The reference's "Small" preset on the GPU: Word = u16, State = u32, PRECISION = 12 (SmallAnsCoder stack.rs:153,
SmallRangeEncoder / SmallRangeDecoder queue.rs:156,747), encoded with Small contiguous categorical models and decoded
with TRUE lookup decoder models (lookup_contiguous.rs:169-333,564-607; non-contiguous alphabets:
lookup_noncontiguous.rs:167,602-645) -- one table entry per 12-bit quantile, staged in shared memory.

Batches as in `constriction_b200.batch` (K independent coders per call); containers hold u16 words (int16 tensors
carrying the bit patterns) and offsets in u16 words."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _native as N
from .batch import _ptr, _require_cuda, _stream_ptr


class SmallModel:
    """M categorical models at 12-bit precision (ctr_small_model_t)."""

    def __init__(self, handle, n_models, alphabet, device):
        self._h, self.n_models, self.alphabet, self.device = handle, n_models, alphabet, device

    def __del__(self):
        try:
            if self._h:
                N.load().ctr_small_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @classmethod
    def categorical(cls, pmf, perfect: bool = False, device=None) -> "SmallModel":
        """SmallContiguousCategoricalEntropyModel::from_floating_point_probabilities_fast / _perfect."""
        _require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        a = np.asarray(pmf)
        if a.ndim == 1:
            a = a[None, :]
        if a.dtype not in (np.float32, np.float64) or a.ndim != 2:
            raise TypeError("pmf must be a float32/float64 array of rank 1 or 2")
        a = np.ascontiguousarray(a)
        lib = N.load()
        fn = lib.ctr_small_model_categorical_f32 if a.dtype == np.float32 else lib.ctr_small_model_categorical_f64
        out = C.c_void_p()
        with torch.cuda.device(dev):
            N.raise_for(fn(a.ctypes.data, 0, a.shape[0], a.shape[1], 1 if perfect else 0, _stream_ptr(), C.byref(out)))
        return cls(out.value, a.shape[0], a.shape[1], dev)

    @classmethod
    def from_cdf(cls, cdf, min_symbol: int = 0, symbols=None, device=None) -> "SmallModel":
        """From 12-bit CDF rows u16[M][alphabet + 1] (from_nonzero_fixed_point_probabilities); `symbols` (int32[alphabet])
        makes the alphabet non-contiguous: index i stands for symbols[i]."""
        _require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        a = np.ascontiguousarray(np.asarray(cdf, dtype=np.uint16))
        if a.ndim == 1:
            a = a[None, :]
        sym = None if symbols is None else np.ascontiguousarray(np.asarray(symbols, dtype=np.int32))
        if sym is not None and sym.size != a.shape[1] - 1:
            raise ValueError("symbols must have one entry per alphabet index")
        out = C.c_void_p()
        with torch.cuda.device(dev):
            N.raise_for(N.load().ctr_small_model_from_cdf(a.ctypes.data, 0, a.shape[0], a.shape[1] - 1, int(min_symbol),
                                                          None if sym is None else sym.ctypes.data, _stream_ptr(), C.byref(out)))
        return cls(out.value, a.shape[0], a.shape[1] - 1, dev)

    def cdf(self) -> np.ndarray:
        out = np.empty((self.n_models, self.alphabet + 1), dtype=np.uint16)
        with torch.cuda.device(self.device):
            N.raise_for(N.load().ctr_small_model_copy_cdf_host(self._h, out.ctypes.data, _stream_ptr()))
        return out


@dataclass
class SmallCompressed:
    words: torch.Tensor       # int16 tensor carrying u16 words
    offsets: torch.Tensor     # int64[K + 1], in u16 words
    n_streams: int
    n_symbols: int
    coder: str
    sym_offsets: Optional[torch.Tensor] = None

    def to_host(self):
        off = self.offsets.cpu().numpy().astype(np.uint64)
        return self.words[: int(off[-1])].cpu().numpy().view(np.uint16), off

    def stream_words(self, k: int) -> np.ndarray:
        lo, hi = int(self.offsets[k].item()), int(self.offsets[k + 1].item())
        return self.words[lo:hi].cpu().numpy().view(np.uint16)


class SmallBatchCoder:
    def __init__(self, device=None):
        _require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._lib = N.load()
        self._ws = None
        self.status = torch.zeros(4, dtype=torch.int32, device=self.device)

    def check(self):
        st = self.status.cpu().numpy().view(np.uint32)
        if st[0] != 0:
            self.status.zero_()
            N.raise_for(int(st[0]), f"stream {int(st[2]) | (int(st[3]) << 32)}")

    def _layout(self, n, k, sym_offsets, model_index):
        L = N.Layout()
        L.n_streams, L.n_symbols = int(k), int(n)
        L.sym_offsets_dev = _ptr(sym_offsets)
        L.model_index_dev = _ptr(model_index)
        L.model_index_mode = N.INDEX_NONE if model_index is None else N.INDEX_PER_STREAM
        return L

    def _encode(self, kind, symbols, model, n_streams, sym_offsets, model_index):
        if symbols.dtype != torch.int32 or not symbols.is_cuda or not symbols.is_contiguous():
            raise TypeError("symbols must be a contiguous CUDA int32 tensor")
        n = symbols.numel()
        k = sym_offsets.numel() - 1 if sym_offsets is not None else n_streams
        L = self._layout(n, k, sym_offsets, model_index)
        ws_bytes = self._lib.ctr_small_encode_workspace_bytes(C.byref(L))
        cap = self._lib.ctr_small_max_compressed_words(C.byref(L))
        if self._ws is None or self._ws.numel() < ws_bytes:
            self._ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=self.device)
        words = torch.empty(cap, dtype=torch.int16, device=self.device)
        offsets = torch.empty(k + 1, dtype=torch.int64, device=self.device)
        fn = self._lib.ctr_small_ans_encode_reverse if kind == "ans" else self._lib.ctr_small_range_encode
        with torch.cuda.device(self.device):
            N.raise_for(fn(model.handle, symbols.data_ptr(), C.byref(L), self._ws.data_ptr(), self._ws.numel(), words.data_ptr(), cap,
                           offsets.data_ptr(), self.status.data_ptr(), _stream_ptr()))
        return SmallCompressed(words, offsets, k, n, kind, sym_offsets)

    def _decode(self, kind, comp, model, model_index, out):
        L = self._layout(comp.n_symbols, comp.n_streams, comp.sym_offsets, model_index)
        if out is None:
            out = torch.empty(comp.n_symbols, dtype=torch.int32, device=self.device)
        fn = self._lib.ctr_small_ans_decode if kind == "ans" else self._lib.ctr_small_range_decode
        with torch.cuda.device(self.device):
            N.raise_for(fn(model.handle, comp.words.data_ptr(), comp.offsets.data_ptr(), C.byref(L), out.data_ptr(), self.status.data_ptr(),
                           _stream_ptr()))
        return out

    def ans_encode(self, symbols, model, n_streams=None, sym_offsets=None, model_index=None):
        return self._encode("ans", symbols, model, n_streams, sym_offsets, model_index)

    def range_encode(self, symbols, model, n_streams=None, sym_offsets=None, model_index=None):
        return self._encode("range", symbols, model, n_streams, sym_offsets, model_index)

    def ans_decode(self, comp, model, model_index=None, out=None):
        return self._decode("ans", comp, model, model_index, out)

    def range_decode(self, comp, model, model_index=None, out=None):
        return self._decode("range", comp, model, model_index, out)
