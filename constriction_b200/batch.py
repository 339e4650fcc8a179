"""Batched coders: K independent ANS / range coders per call, device tensors in and out.

PyTorch is only the plumbing here (device memory, streams); every number is produced by the CUDA
kernels in libconstriction_b200.so through the C ABI (include/constriction_b200.h).

A batch is described like in the C ABI:
  * interleaved layout (default): stream k owns symbols k, k+K, k+2K, ... of a flat int32 tensor;
  * contiguous layout: `sym_offsets` (int64[K+1]) gives stream k the slice [off[k], off[k+1]).
Stream k's compressed words are bit-identical to what the reference's `AnsCoder.get_compressed()` /
`RangeEncoder.get_compressed()` returns for one coder fed stream k's symbols
(reference: src/stream/stack.rs:1014-1100, src/stream/queue.rs:612-705,968-1035).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _native as N


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("constriction_b200 needs a CUDA device: there is no CPU fallback")


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class ModelTable:
    """Device-resident 24-bit CDF tables of M models over one alphabet (ctr_model_t)."""

    def __init__(self, handle: int, device: torch.device):
        self._h = handle
        self.device = device
        lib = N.load()
        nm, al, ms = C.c_uint32(), C.c_uint32(), C.c_int32()
        N.raise_for(lib.ctr_model_info(handle, C.byref(nm), C.byref(al), C.byref(ms)))
        self.n_models, self.alphabet, self.min_symbol = nm.value, al.value, ms.value

    def __del__(self):
        try:
            if self._h:
                N.load().ctr_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self) -> int:
        return self._h

    # -- constructors --------------------------------------------------------------------------
    @staticmethod
    def _device(device) -> torch.device:
        _require_cuda()
        return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    @classmethod
    def quantized_gaussian(cls, min_symbol: int, max_symbol: int, means, stds, device=None) -> "ModelTable":
        """M leakily-quantised Gaussians (quantize.rs:284-308,525-568), one per (mean, std) pair."""
        dev = cls._device(device)
        means = np.ascontiguousarray(np.atleast_1d(np.asarray(means, dtype=np.float64)))
        stds = np.ascontiguousarray(np.atleast_1d(np.asarray(stds, dtype=np.float64)))
        if means.shape != stds.shape or means.ndim != 1:
            raise ValueError("means and stds must be 1-D arrays of equal length")
        out = C.c_void_p()
        with torch.cuda.device(dev):
            rc = N.load().ctr_model_quantized_gaussian(int(min_symbol), int(max_symbol), means.ctypes.data,
                                                       stds.ctypes.data, means.size, _stream_ptr(), C.byref(out))
        N.raise_for(rc)
        return cls(out.value, dev)

    @classmethod
    def categorical(cls, pmf, device=None) -> "ModelTable":
        """M categorical models from float pmf rows, `fast_quantized_cdf` rounding in the array's own
        float type (categorical.rs:16-54).  `pmf`: numpy or torch, shape (alphabet,) or (M, alphabet)."""
        dev = cls._device(device)
        lib = N.load()
        out = C.c_void_p()
        if isinstance(pmf, torch.Tensor):
            t = pmf.contiguous()
            if t.dim() == 1:
                t = t[None, :]
            if t.dtype not in (torch.float32, torch.float64) or t.dim() != 2:
                raise TypeError("pmf must be a float32/float64 array of rank 1 or 2")
            fn = lib.ctr_model_categorical_f32 if t.dtype == torch.float32 else lib.ctr_model_categorical_f64
            if t.is_cuda:
                dev = t.device
                with torch.cuda.device(dev):
                    rc = fn(t.data_ptr(), 1, t.shape[0], t.shape[1], _stream_ptr(), C.byref(out))
                    torch.cuda.current_stream().synchronize()
            else:
                with torch.cuda.device(dev):
                    rc = fn(t.data_ptr(), 0, t.shape[0], t.shape[1], _stream_ptr(), C.byref(out))
        else:
            a = np.asarray(pmf)
            if a.ndim == 1:
                a = a[None, :]
            if a.dtype not in (np.float32, np.float64) or a.ndim != 2:
                raise TypeError("pmf must be a float32/float64 array of rank 1 or 2")
            a = np.ascontiguousarray(a)
            fn = lib.ctr_model_categorical_f32 if a.dtype == np.float32 else lib.ctr_model_categorical_f64
            with torch.cuda.device(dev):
                rc = fn(a.ctypes.data, 0, a.shape[0], a.shape[1], _stream_ptr(), C.byref(out))
        N.raise_for(rc)
        return cls(out.value, dev)

    @classmethod
    def categorical_perfect(cls, pmf, device=None) -> "ModelTable":
        """M categorical models with the reference's `perfect` quantisation (categorical.rs:56-177), the default of the
        Python `Categorical(p)` / `Bernoulli(p)`.  `pmf`: numpy float32 / float64, shape (alphabet,) or (M, alphabet)."""
        dev = cls._device(device)
        a = np.asarray(pmf)
        if a.ndim == 1:
            a = a[None, :]
        if a.dtype not in (np.float32, np.float64) or a.ndim != 2:
            raise TypeError("pmf must be a float32/float64 array of rank 1 or 2")
        a = np.ascontiguousarray(a)
        lib = N.load()
        fn = lib.ctr_model_categorical_perfect_f32 if a.dtype == np.float32 else lib.ctr_model_categorical_perfect_f64
        out = C.c_void_p()
        with torch.cuda.device(dev):
            rc = fn(a.ctypes.data, 0, a.shape[0], a.shape[1], _stream_ptr(), C.byref(out))
        N.raise_for(rc)
        return cls(out.value, dev)

    KINDS = {"gaussian": 0, "laplace": 1, "cauchy": 2}

    @classmethod
    def quantized(cls, kind: str, min_symbol: int, max_symbol: int, p0, p1, device=None) -> "ModelTable":
        """M leakily quantised two-parameter distributions: kind "gaussian" (mean, std), "laplace" (mean, scale),
        "cauchy" (location, scale); quantize.rs:284-308,525-568."""
        dev = cls._device(device)
        p0 = np.ascontiguousarray(np.atleast_1d(np.asarray(p0, dtype=np.float64)))
        p1 = np.ascontiguousarray(np.atleast_1d(np.asarray(p1, dtype=np.float64)))
        if p0.shape != p1.shape or p0.ndim != 1:
            raise ValueError("parameters must be 1-D arrays of equal length")
        out = C.c_void_p()
        with torch.cuda.device(dev):
            rc = N.load().ctr_model_quantized(cls.KINDS[kind], int(min_symbol), int(max_symbol), p0.ctypes.data, p1.ctypes.data,
                                              p0.size, _stream_ptr(), C.byref(out))
        N.raise_for(rc)
        return cls(out.value, dev)

    @classmethod
    def binomial(cls, ns, ps, device=None) -> "ModelTable":
        """M Binomial(n, p) models over {0..n}, rows padded to the widest (pybindings/stream/model.rs:925-960)."""
        dev = cls._device(device)
        ns = np.ascontiguousarray(np.atleast_1d(np.asarray(ns, dtype=np.int32)))
        ps = np.ascontiguousarray(np.atleast_1d(np.asarray(ps, dtype=np.float64)))
        if ns.shape != ps.shape or ns.ndim != 1:
            raise ValueError("parameters must be 1-D arrays of equal length")
        out = C.c_void_p()
        with torch.cuda.device(dev):
            rc = N.load().ctr_model_binomial(ns.ctypes.data, ps.ctypes.data, ns.size, _stream_ptr(), C.byref(out))
        N.raise_for(rc)
        return cls(out.value, dev)

    @classmethod
    def from_cdf(cls, cdf, min_symbol: int = 0, device=None) -> "ModelTable":
        """M models from fixed-point CDF rows u32[M][alphabet+1] (cdf[0]=0, cdf[-1]=2^24)."""
        dev = cls._device(device)
        a = np.ascontiguousarray(np.asarray(cdf, dtype=np.uint32))
        if a.ndim == 1:
            a = a[None, :]
        out = C.c_void_p()
        with torch.cuda.device(dev):
            rc = N.load().ctr_model_from_cdf(a.ctypes.data, 0, a.shape[0], a.shape[1] - 1, int(min_symbol),
                                             _stream_ptr(), C.byref(out))
        N.raise_for(rc)
        return cls(out.value, dev)

    @classmethod
    def uniform(cls, size: int, device=None) -> "ModelTable":
        dev = cls._device(device)
        out = C.c_void_p()
        with torch.cuda.device(dev):
            rc = N.load().ctr_model_uniform(int(size), _stream_ptr(), C.byref(out))
        N.raise_for(rc)
        return cls(out.value, dev)

    def cdf(self) -> np.ndarray:
        """The CDF rows as a host array u32[M][alphabet+1]."""
        out = np.empty((self.n_models, self.alphabet + 1), dtype=np.uint32)
        with torch.cuda.device(self.device):
            N.raise_for(N.load().ctr_model_copy_cdf_host(self._h, out.ctypes.data, _stream_ptr()))
        return out


class GaussianParams:
    """Per-symbol QuantizedGaussian parameters on the device: means[i], stds[i] (float64, laid out like the symbols)
    over the support [min_symbol, max_symbol].  Passed to the BatchCoder methods in place of a ModelTable, it
    selects the table-free kernels (ctr_*_gaussian in include/constriction_b200.h): the reference's lazily
    evaluated per-symbol models (pybindings/stream/model/internals.rs:188-249, quantize.rs:525-568,580-779)."""

    def __init__(self, min_symbol: int, max_symbol: int, means, stds, device=None):
        _require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

        def col(x):
            t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64)))
            t = t.to(device=dev, dtype=torch.float64).contiguous()
            if t.dim() != 1:
                raise ValueError("means and stds must be 1-D")
            return t

        self.min_symbol, self.max_symbol = int(min_symbol), int(max_symbol)
        if not self.max_symbol > self.min_symbol:
            raise ValueError("max_symbol must be greater than min_symbol")
        self.means, self.stds = col(means), col(stds)
        if self.means.numel() != self.stds.numel():
            raise ValueError("Model parameters have unequal shape")
        self.device = dev

    def __len__(self):
        return self.means.numel()


@dataclass
class Compressed:
    """Container of a batch: dense `words` (u32 stored in an int32 tensor) + `offsets` (int64[K+1])."""
    words: torch.Tensor
    offsets: torch.Tensor
    n_streams: int
    n_symbols: int
    coder: str
    sym_offsets: Optional[torch.Tensor] = None
    states: Optional[torch.Tensor] = None
    # checkpoints (CTR_FLAG_CHECKPOINTS): records every `checkpoint_every` symbols, see include/constriction_b200.h
    checkpoint_every: int = 0
    ckpt_offsets: Optional[torch.Tensor] = None
    checkpoints: Optional[torch.Tensor] = None

    def total_words(self) -> int:
        return int(self.offsets[-1].item())

    def to_host(self):
        """(words u32[total], offsets u64[K+1]) as numpy arrays (synchronises)."""
        off = self.offsets.cpu().numpy().astype(np.uint64)
        total = int(off[-1])
        return self.words[:total].cpu().numpy().view(np.uint32), off

    def stream_words(self, k: int) -> np.ndarray:
        lo, hi = int(self.offsets[k].item()), int(self.offsets[k + 1].item())
        return self.words[lo:hi].cpu().numpy().view(np.uint32)


class BatchCoder:
    """Launches the batched coder kernels on the current CUDA stream and caches their workspace."""

    def __init__(self, device=None):
        _require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._lib = N.load()
        self._ws: Optional[torch.Tensor] = None
        self.status = torch.zeros(4, dtype=torch.int32, device=self.device)

    # -- helpers -----------------------------------------------------------------------------------
    def _layout(self, n_symbols, n_streams, sym_offsets, model_index, index_mode, raw, ckpt=None) -> N.Layout:
        if index_mode is None:
            index_mode = N.INDEX_NONE if model_index is None else (
                N.INDEX_PER_STREAM if model_index.numel() == n_streams and n_streams != n_symbols else N.INDEX_PER_SYMBOL)
        L = N.Layout()
        L.n_streams = int(n_streams)
        L.n_symbols = int(n_symbols)
        L.sym_offsets_dev = _ptr(sym_offsets)
        L.model_index_dev = _ptr(model_index)
        L.model_index_mode = int(index_mode)
        L.flags = N.FLAG_RAW if raw else 0
        if ckpt is not None:
            every, ckpt_off, records = ckpt
            L.flags |= N.FLAG_CHECKPOINTS
            L.checkpoint_every = int(every)
            L.ckpt_offsets_dev = _ptr(ckpt_off)
            L.checkpoints_dev = _ptr(records)
        return L

    def _check_inputs(self, sym_offsets, model_index, n_streams):
        if sym_offsets is not None:
            if sym_offsets.dtype != torch.int64 or not sym_offsets.is_cuda or sym_offsets.numel() != n_streams + 1:
                raise TypeError("sym_offsets must be a CUDA int64 tensor of length n_streams + 1")
        if model_index is not None:
            if model_index.dtype != torch.int32 or not model_index.is_cuda or not model_index.is_contiguous():
                raise TypeError("model_index must be a contiguous CUDA int32 tensor")

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
        return self._ws

    def check(self) -> None:
        """Synchronises and raises the exception the reference would have raised for a data error."""
        st = self.status.cpu().numpy().view(np.uint32)
        if st[0] != 0:
            self.status.zero_()
            N.raise_for(int(st[0]), f"stream {int(st[2]) | (int(st[3]) << 32)}")

    # -- encode ------------------------------------------------------------------------------------
    def _encode(self, kind, symbols, model, n_streams, sym_offsets, model_index, index_mode, states_in, raw,
                want_states, out: Optional[Compressed], checkpoint_every=0):
        if symbols.dtype != torch.int32 or not symbols.is_cuda or not symbols.is_contiguous() or symbols.dim() != 1:
            raise TypeError("symbols must be a contiguous 1-D CUDA int32 tensor")
        n = symbols.numel()
        if sym_offsets is not None:
            n_streams = sym_offsets.numel() - 1
        if n_streams is None:
            raise ValueError("n_streams is required for the interleaved layout")
        self._check_inputs(sym_offsets, model_index, n_streams)
        L = self._layout(n, n_streams, sym_offsets, model_index, index_mode, raw)
        lib = self._lib
        ckpt_off = records = None
        if checkpoint_every:
            if sym_offsets is None:
                raise ValueError("checkpoints need the contiguous layout (sym_offsets)")
            L.checkpoint_every = int(checkpoint_every)
            ckpt_off = torch.empty(n_streams + 1, dtype=torch.int64, device=self.device)
            with torch.cuda.device(self.device):
                N.raise_for(lib.ctr_checkpoint_offsets(C.byref(L), ckpt_off.data_ptr(), _stream_ptr()))
            records = torch.empty(int(lib.ctr_checkpoint_max_records(C.byref(L))) * (2 if kind == "ans" else 4),
                                  dtype=torch.int64, device=self.device)
            L = self._layout(n, n_streams, sym_offsets, model_index, index_mode, raw, (checkpoint_every, ckpt_off, records))
        ws_bytes = lib.ctr_ans_encode_workspace_bytes(C.byref(L))
        cap = lib.ctr_ans_max_compressed_words(C.byref(L))
        ws = self._workspace(ws_bytes)
        if out is not None and out.offsets.numel() == n_streams + 1:
            # the caller's container (e.g. a slot of a multi-GPU receive buffer); it may hold fewer words than the
            # worst case: a batch that does not fit is reported as MemoryError by check()
            words, offsets = out.words, out.offsets
            cap = words.numel()
        else:
            words = torch.empty(cap, dtype=torch.int32, device=self.device)
            offsets = torch.empty(n_streams + 1, dtype=torch.int64, device=self.device)
        state_words = 1 if kind == "ans" else 4
        states_out = None
        if want_states:
            states_out = torch.empty(n_streams * state_words, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            if isinstance(model, GaussianParams):
                if len(model) != n:
                    raise ValueError("`symbols` argument has wrong length.")
                fn = lib.ctr_ans_encode_reverse_gaussian if kind == "ans" else lib.ctr_range_encode_gaussian
                rc = fn(model.min_symbol, model.max_symbol, model.means.data_ptr(), model.stds.data_ptr(),
                        symbols.data_ptr(), C.byref(L), _ptr(states_in), ws.data_ptr(), ws.numel(), words.data_ptr(), cap,
                        offsets.data_ptr(), _ptr(states_out), self.status.data_ptr(), _stream_ptr())
            else:
                fn = lib.ctr_ans_encode_reverse if kind == "ans" else lib.ctr_range_encode
                rc = fn(model.handle, symbols.data_ptr(), C.byref(L), _ptr(states_in), ws.data_ptr(), ws.numel(),
                        words.data_ptr(), cap, offsets.data_ptr(), _ptr(states_out), self.status.data_ptr(), _stream_ptr())
        N.raise_for(rc)
        return Compressed(words, offsets, n_streams, n, kind, sym_offsets, states_out, int(checkpoint_every), ckpt_off, records)

    def ans_encode(self, symbols, model: ModelTable, n_streams=None, sym_offsets=None, model_index=None,
                   index_mode=None, states_in=None, raw=False, want_states=False, out=None, checkpoint_every=0) -> Compressed:
        """Every stream encodes its symbols in reverse order (AnsCoder.encode_reverse on K coders).
        checkpoint_every=C (multiple of 32, contiguous layout): also record the coder positions every C symbols, so
        that `ans_decode` can decode every chunk on its own lane."""
        return self._encode("ans", symbols, model, n_streams, sym_offsets, model_index, index_mode, states_in, raw,
                            want_states, out, checkpoint_every)

    def range_encode(self, symbols, model: ModelTable, n_streams=None, sym_offsets=None, model_index=None,
                     index_mode=None, states_in=None, raw=False, want_states=False, out=None, checkpoint_every=0) -> Compressed:
        """Every stream encodes its symbols in forward order (RangeEncoder.encode on K coders)."""
        return self._encode("range", symbols, model, n_streams, sym_offsets, model_index, index_mode, states_in, raw,
                            want_states, out, checkpoint_every)

    # -- decode ------------------------------------------------------------------------------------
    def _decode(self, kind, words, offsets, model, n_symbols, sym_offsets, model_index, index_mode, states_in, raw,
                want_states, want_pos, out, ckpt=None):
        n_streams = offsets.numel() - 1
        if offsets.dtype != torch.int64 or not offsets.is_cuda:
            raise TypeError("offsets must be a CUDA int64 tensor")
        if words.dtype != torch.int32 or not words.is_cuda or not words.is_contiguous():
            raise TypeError("words must be a contiguous CUDA int32 tensor (u32 bit patterns)")
        self._check_inputs(sym_offsets, model_index, n_streams)
        L = self._layout(n_symbols, n_streams, sym_offsets, model_index, index_mode, raw, ckpt)
        if out is None:
            out = torch.empty(n_symbols, dtype=torch.int32, device=self.device)
        state_words = 1 if kind == "ans" else 4
        states_out = torch.empty(n_streams * state_words, dtype=torch.int64, device=self.device) if want_states else None
        pos = torch.empty(n_streams, dtype=torch.int64, device=self.device) if want_pos else None
        with torch.cuda.device(self.device):
            if isinstance(model, GaussianParams):
                if len(model) != n_symbols:
                    raise ValueError("the number of model parameters differs from the number of symbols")
                fn = self._lib.ctr_ans_decode_gaussian if kind == "ans" else self._lib.ctr_range_decode_gaussian
                rc = fn(model.min_symbol, model.max_symbol, model.means.data_ptr(), model.stds.data_ptr(), words.data_ptr(),
                        offsets.data_ptr(), C.byref(L), _ptr(states_in), out.data_ptr(), _ptr(states_out), _ptr(pos),
                        self.status.data_ptr(), _stream_ptr())
            else:
                fn = self._lib.ctr_ans_decode if kind == "ans" else self._lib.ctr_range_decode
                rc = fn(model.handle, words.data_ptr(), offsets.data_ptr(), C.byref(L), _ptr(states_in), out.data_ptr(),
                        _ptr(states_out), _ptr(pos), self.status.data_ptr(), _stream_ptr())
        N.raise_for(rc)
        if want_states or want_pos:
            return out, states_out, pos
        return out

    def ans_decode(self, compressed: Compressed, model: ModelTable, n_symbols=None, model_index=None,
                   index_mode=None, states_in=None, raw=False, want_states=False, want_pos=False, out=None,
                   use_checkpoints=True):
        """Every stream decodes its symbols in forward order (AnsCoder.decode on K coders)."""
        n = compressed.n_symbols if n_symbols is None else n_symbols
        ckpt = None
        if use_checkpoints and compressed.checkpoint_every and not raw:
            ckpt = (compressed.checkpoint_every, compressed.ckpt_offsets, compressed.checkpoints)
        return self._decode("ans", compressed.words, compressed.offsets, model, n, compressed.sym_offsets,
                            model_index, index_mode, states_in, raw, want_states, want_pos, out, ckpt)

    def range_decode(self, compressed: Compressed, model: ModelTable, n_symbols=None, model_index=None,
                     index_mode=None, states_in=None, raw=False, want_states=False, want_pos=False, out=None,
                   use_checkpoints=True):
        """Every stream decodes its symbols in forward order (RangeDecoder.decode on K coders)."""
        n = compressed.n_symbols if n_symbols is None else n_symbols
        ckpt = None
        if use_checkpoints and compressed.checkpoint_every and not raw:
            ckpt = (compressed.checkpoint_every, compressed.ckpt_offsets, compressed.checkpoints)
        return self._decode("range", compressed.words, compressed.offsets, model, n, compressed.sym_offsets,
                            model_index, index_mode, states_in, raw, want_states, want_pos, out, ckpt)


def kernel_launch_count() -> int:
    return int(N.load().ctr_kernel_launch_count())
