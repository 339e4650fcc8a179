"""Wire / on-disk form of a batch container (C ABI: ctr_container_*, csrc/container.cu): a little-endian header, the
offset tables, optional Pos::pos() records and the words.  Any stream -- words[offsets[k]:offsets[k+1]] -- is one stock
constriction stream; with records, stock coders can `seek` to every chunk boundary (src/lib.rs:425-580)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N
from .batch import Compressed


def _host_u64(t) -> np.ndarray:
    return np.ascontiguousarray(t.cpu().numpy().astype(np.int64).view(np.uint64))


def pack(comp: Compressed) -> bytes:
    """Serialises a (Default preset) container produced by BatchCoder.ans_encode / range_encode."""
    lib = N.load()
    off = _host_u64(comp.offsets)
    total = int(off[-1])
    words = np.ascontiguousarray(comp.words[:total].cpu().numpy().view(np.uint32))
    v = N.ContainerView()
    v.coder = 0 if comp.coder == "ans" else 1
    v.word_bits, v.precision = 32, 24
    v.n_streams, v.n_symbols, v.total_words = comp.n_streams, comp.n_symbols, total
    keep = [off, words]
    v.offsets = off.ctypes.data
    v.words = words.ctypes.data if total else None
    if comp.sym_offsets is None:
        v.flags = 1
    else:
        so = _host_u64(comp.sym_offsets)
        keep.append(so)
        v.sym_offsets = so.ctypes.data
    if comp.checkpoint_every:
        co = _host_u64(comp.ckpt_offsets)
        n_rec = int(co[-1])
        rw = 2 if comp.coder == "ans" else 4
        rec = np.ascontiguousarray(comp.checkpoints[: n_rec * rw].cpu().numpy().view(np.uint64))
        keep += [co, rec]
        v.checkpoint_every, v.n_records = comp.checkpoint_every, n_rec
        v.ckpt_offsets, v.records = co.ctypes.data, rec.ctypes.data if n_rec else None
    size = lib.ctr_container_size(C.byref(v))
    if size == 0:
        raise ValueError("container cannot be serialised")
    out = np.zeros(size // 8, dtype=np.uint64)
    N.raise_for(lib.ctr_container_pack(C.byref(v), out.ctypes.data, size))
    return out.tobytes()


def unpack_host(data: bytes):
    """Validates `data` and returns its parts as numpy arrays (views of one aligned copy): a dict with coder, n_streams,
    n_symbols, offsets, words and, if present, sym_offsets / checkpoint_every / ckpt_offsets / records."""
    lib = N.load()
    buf = np.frombuffer(data, dtype=np.uint8)
    aligned = np.zeros((buf.size + 7) // 8, dtype=np.uint64)
    aligned.view(np.uint8)[: buf.size] = buf
    v = N.ContainerView()
    N.raise_for(lib.ctr_container_unpack(aligned.ctypes.data, buf.size, C.byref(v)))
    base = aligned.ctypes.data

    def arr(ptr, count, dtype):
        if not ptr or not count:
            return np.empty(0, dtype=dtype)
        start = (ptr - base) // np.dtype(dtype).itemsize
        return aligned.view(dtype)[start:start + count]

    k1 = v.n_streams + 1
    out = dict(coder="ans" if v.coder == 0 else "range", word_bits=v.word_bits, precision=v.precision, n_streams=int(v.n_streams),
               n_symbols=int(v.n_symbols), offsets=arr(v.offsets, k1, np.uint64),
               words=arr(v.words, v.total_words, np.uint32 if v.word_bits == 32 else np.uint16),
               sym_offsets=arr(v.sym_offsets, k1, np.uint64) if v.sym_offsets else None, checkpoint_every=int(v.checkpoint_every))
    if v.checkpoint_every:
        out["ckpt_offsets"] = arr(v.ckpt_offsets, k1, np.uint64)
        out["records"] = arr(v.records, v.n_records * (2 if v.coder == 0 else 4), np.uint64)
    return out


def unpack(data: bytes, device=None) -> Compressed:
    """A serialised Default-preset container back on the device, ready for BatchCoder.ans_decode / range_decode."""
    h = unpack_host(data)
    if h["word_bits"] != 32:
        raise ValueError("not a Default-preset container")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    to_dev = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).view(dt).copy()).to(dev)  # noqa: E731
    words = to_dev(np.concatenate([h["words"], np.zeros(4, dtype=np.uint32)]), np.int32)  # (decoders read whole 16-byte blocks)
    comp = Compressed(words, to_dev(h["offsets"], np.int64), h["n_streams"], h["n_symbols"], h["coder"],
                      None if h["sym_offsets"] is None else to_dev(h["sym_offsets"], np.int64))
    if h["checkpoint_every"]:
        comp.checkpoint_every = h["checkpoint_every"]
        comp.ckpt_offsets = to_dev(h["ckpt_offsets"], np.int64)
        comp.checkpoints = to_dev(h["records"], np.int64)
    return comp
