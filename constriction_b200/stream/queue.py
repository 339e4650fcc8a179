"""`constriction.stream.queue.RangeEncoder` / `RangeDecoder` (reference:
src/pybindings/stream/queue.rs:124-685) on the batched CUDA kernels; coder state is carried between
calls through the raw-state interface of the C ABI."""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as N
from ..batch import GaussianParams
from ..batch import Compressed
from ._common import coder, from_dev_u64, is_scalar, symbols_array, to_dev_i32, to_dev_u64

_U64_MAX = (1 << 64) - 1


class RangeEncoder:
    def __init__(self):
        self.clear()

    def clear(self):
        self._bulk = np.zeros(0, dtype=np.uint32)
        self._st = [0, _U64_MAX, 0, 0]  # lower, range, num_inverted, first_inverted_word (queue.rs:98-106)

    def pos(self):  # queue.rs:182-196
        return (int(self._bulk.size) + int(self._st[2]), (int(self._st[0]), int(self._st[1])))

    def is_empty(self):
        return self._bulk.size == 0 and self._st[1] == _U64_MAX

    def _run(self, symbols: np.ndarray, table, per_symbol: bool, raw: bool):
        bc = coder()
        n = symbols.size
        sym = to_dev_i32(symbols) if n else torch.zeros(1, dtype=torch.int32, device=bc.device)[:0]
        per_symbol = per_symbol and not isinstance(table, GaussianParams)  # table-free kernels need no index
        idx = torch.arange(n, dtype=torch.int32, device=bc.device) if per_symbol else None
        comp = bc.range_encode(sym, table, n_streams=1, model_index=idx,
                               index_mode=N.INDEX_PER_SYMBOL if per_symbol else N.INDEX_NONE,
                               states_in=to_dev_u64(self._st), raw=raw, want_states=True)
        words, _ = comp.to_host()
        st = from_dev_u64(comp.states)
        bc.check()
        return words, st

    def _seal_words(self) -> np.ndarray:
        # zero symbols, non-raw: the kernel emits exactly the seal words of the current state
        from . import model as M
        words, _ = self._run(np.zeros(0, dtype=np.int32), _dummy_table(), False, raw=False)
        return words

    def num_words(self):
        return int(self._bulk.size) + int(self._seal_words().size)

    def num_bits(self):
        return 32 * self.num_words()

    def get_compressed(self):
        return np.concatenate([self._bulk, self._seal_words()])

    def get_decoder(self):
        return RangeDecoder(self.get_compressed())

    def clone(self):
        c = RangeEncoder()
        c._bulk = self._bulk.copy()
        c._st = list(self._st)
        return c

    def _encode(self, symbols, table, per_symbol):
        try:
            words, st = self._run(symbols, table, per_symbol, raw=True)
        except KeyError:
            # reference semantics (stream/mod.rs:592-607): the symbols before the impossible one stay encoded
            from ._common import first_impossible
            done = first_impossible(symbols, table, per_symbol, reverse=False)
            if 0 < done < symbols.size:
                from ..batch import GaussianParams as G, ModelTable
                if per_symbol and isinstance(table, G):
                    table = G(table.min_symbol, table.max_symbol, table.means[:done], table.stds[:done])
                elif per_symbol:
                    table = ModelTable.from_cdf(table.cdf()[:done], table.min_symbol)
                words, st = self._run(symbols[:done], table, per_symbol, raw=True)
                self._bulk = np.concatenate([self._bulk, words])
                self._st = st
            raise
        self._bulk = np.concatenate([self._bulk, words])
        self._st = st

    def encode(self, symbols, model, *params):
        if is_scalar(symbols):
            if params:
                raise ValueError("To encode a single symbol, use a concrete model, i.e., pass the model parameters "
                                 "directly to the constructor of the model and not to the `encode` method.")
            self._encode(np.array([symbols], dtype=np.int32), model._concrete_table(), False)
            return
        symbols = symbols_array(symbols)
        if not params:
            self._encode(symbols, model._concrete_table(), False)
        else:
            if symbols.size != model._family_len(params):
                raise ValueError("`symbols` argument has wrong length.")
            self._encode(symbols, model._family_table(params), True)


_DUMMY = None


def _dummy_table():
    """Any valid table: calls that code zero symbols still need a model handle."""
    global _DUMMY
    if _DUMMY is None:
        from ..batch import ModelTable
        _DUMMY = ModelTable.uniform(2)
    return _DUMMY


class RangeDecoder:
    def __init__(self, compressed):
        w = np.asarray(compressed)
        if w.dtype != np.uint32 or w.ndim != 1:
            raise TypeError("compressed must be a rank-1 numpy array with dtype=np.uint32")
        self._words = np.ascontiguousarray(w).copy()
        self._pos = 0
        self._st = [0, _U64_MAX, 0, 0]
        self._read_point()

    def _run(self, n, table, per_symbol, raw):
        bc = coder()
        rest = self._words[self._pos:]
        words = to_dev_i32(rest) if rest.size else torch.zeros(1, dtype=torch.int32, device=bc.device)
        offsets = torch.tensor([0, rest.size], dtype=torch.int64, device=bc.device)
        per_symbol = per_symbol and not isinstance(table, GaussianParams)  # table-free kernels need no index
        idx = torch.arange(n, dtype=torch.int32, device=bc.device) if per_symbol else None
        comp = Compressed(words, offsets, 1, n, "range")
        out, st, pos = bc.range_decode(comp, table, n_symbols=n, model_index=idx,
                                       index_mode=N.INDEX_PER_SYMBOL if per_symbol else N.INDEX_NONE,
                                       states_in=to_dev_u64(self._st) if raw else None, raw=raw, want_states=True,
                                       want_pos=True)
        result = out.cpu().numpy()
        new_st = from_dev_u64(st)
        consumed = int(pos.cpu().numpy()[0])
        bc.check()
        self._st = new_st
        self._pos += consumed
        return result

    def _read_point(self):
        # zero symbols, non-raw: the kernel performs from_compressed / read_point (queue.rs:755-773,847-868)
        self._run(0, _dummy_table(), False, raw=False)

    def pos(self):
        return (self._pos, (int(self._st[0]), int(self._st[1])))

    def seek(self, position, state):  # queue.rs:911-928
        lower, rng = state
        if position > self._words.size or (int(rng) >> 32) == 0:
            N.raise_for(N.ERR_SEEK)
        self._pos = int(position)
        self._read_point()
        self._st[0], self._st[1] = int(lower), int(rng)

    def maybe_exhausted(self):  # queue.rs:872-883
        lower, rng, point = self._st[0], self._st[1], self._st[2]
        diff = (point - lower) & _U64_MAX
        return self._pos >= self._words.size and (rng == _U64_MAX or diff < (1 << 33) - 1)

    def decode(self, model, *params):
        if len(params) == 0:
            return int(self._run(1, model._concrete_table(), False, True)[0])
        if len(params) == 1 and is_scalar(params[0]):
            return self._run(int(params[0]), model._concrete_table(), False, True)
        n = model._family_len(params)
        return self._run(n, model._family_table(params), True, True)
