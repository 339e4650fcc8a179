"""`constriction.stream.stack.AnsCoder` (reference: src/pybindings/stream/stack.rs:197-763) on the
batched CUDA kernels.  The coder's bulk words and 64-bit state live on the host between calls and are
handed to the kernels through the raw-state interface of the C ABI (CTR_FLAG_RAW)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as N
from ..batch import GaussianParams
from ._common import coder, first_impossible, from_dev_u64, is_scalar, symbols_array, to_dev_i32, to_dev_u64

_MASK32 = 0xFFFFFFFF


class AnsCoder:
    def __init__(self, compressed=None, seal=False):
        self._bulk = np.zeros(0, dtype=np.uint32)
        self._state = 0
        if compressed is None:
            if seal:
                raise ValueError("Need compressed data to seal.")
            return
        w = np.asarray(compressed)
        if w.dtype != np.uint32 or w.ndim != 1:
            raise TypeError("compressed must be a rank-1 numpy array with dtype=np.uint32")
        words = [int(x) for x in w]
        if seal:  # from_binary, stack.rs:341-360
            state = 1
            while state < (1 << 32) and words:
                state = (state << 32) | words.pop()
        else:  # from_compressed, stack.rs:299-318,440-462
            state = 0
            if words:
                first = words.pop()
                if first == 0:
                    N.raise_for(N.ERR_TRAILING_ZERO)
                state = first
                if words:
                    state = (state << 32) | words.pop()
        self._bulk = np.array(words, dtype=np.uint32)
        self._state = state

    # -- introspection (stack.rs:259-353) ----------------------------------------------------------
    def pos(self):
        return (int(self._bulk.size), int(self._state))

    def seek(self, position, state):
        if position > self._bulk.size:
            N.raise_for(N.ERR_SEEK)
        self._bulk = self._bulk[:position].copy()
        self._state = int(state)

    def clear(self):
        self._bulk = np.zeros(0, dtype=np.uint32)
        self._state = 0

    def _state_words(self):
        s = self._state
        if s == 0:
            return []
        return [s & _MASK32] if s >> 32 == 0 else [s & _MASK32, s >> 32]

    def num_words(self):
        return int(self._bulk.size) + len(self._state_words())

    def num_bits(self):
        return 32 * self.num_words()

    def num_valid_bits(self):
        return 32 * int(self._bulk.size) + max(self._state.bit_length(), 1) - 1

    def is_empty(self):
        return self._bulk.size == 0 and self._state == 0

    def get_compressed(self, unseal=False):
        if unseal:  # get_binary, stack.rs:549-556,944-955,1164-1171
            s = self._state
            if s == 0 or (s.bit_length() - 1) % 32 != 0:
                N.raise_for(N.ERR_NOT_SEALED)
            tail = [s & _MASK32] if s.bit_length() - 1 == 32 else []
            return np.concatenate([self._bulk, np.array(tail, dtype=np.uint32)])
        return np.concatenate([self._bulk, np.array(self._state_words(), dtype=np.uint32)])

    def clone(self):
        c = AnsCoder()
        c._bulk = self._bulk.copy()
        c._state = self._state
        return c

    # -- coding ------------------------------------------------------------------------------------
    def _encode(self, symbols: np.ndarray, table, per_symbol: bool):
        try:
            self._encode_all(symbols, table, per_symbol)
        except KeyError:
            # reference semantics: the symbols coded before the impossible one stay on the coder (it codes the LAST
            # symbols first); the kernels rejected the whole call, so that part is encoded again
            done = first_impossible(symbols, table, per_symbol, reverse=True)
            if 0 < done < symbols.size:
                keep = slice(symbols.size - done, symbols.size)
                from ..batch import GaussianParams as G, ModelTable
                if per_symbol and isinstance(table, G):
                    table = G(table.min_symbol, table.max_symbol, table.means[keep], table.stds[keep])
                elif per_symbol:
                    table = ModelTable.from_cdf(table.cdf()[keep], table.min_symbol)
                self._encode_all(symbols[keep], table, per_symbol)
            raise

    def _encode_all(self, symbols: np.ndarray, table, per_symbol: bool):
        bc = coder()
        n = symbols.size
        per_symbol = per_symbol and not isinstance(table, GaussianParams)  # table-free kernels need no index
        idx = torch.arange(n, dtype=torch.int32, device=bc.device) if per_symbol else None
        comp = bc.ans_encode(to_dev_i32(symbols), table, n_streams=1, model_index=idx,
                             index_mode=N.INDEX_PER_SYMBOL if per_symbol else N.INDEX_NONE,
                             states_in=to_dev_u64([self._state]), raw=True, want_states=True)
        words, _ = comp.to_host()
        new_state = from_dev_u64(comp.states)[0]
        bc.check()
        self._bulk = np.concatenate([self._bulk, words])
        self._state = new_state

    def encode_reverse(self, symbols, model, *params):
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

    def _decode(self, n: int, table, per_symbol: bool) -> np.ndarray:
        bc = coder()
        words = to_dev_i32(self._bulk) if self._bulk.size else torch.zeros(1, dtype=torch.int32, device=bc.device)
        offsets = torch.tensor([0, self._bulk.size], dtype=torch.int64, device=bc.device)
        per_symbol = per_symbol and not isinstance(table, GaussianParams)  # table-free kernels need no index
        idx = torch.arange(n, dtype=torch.int32, device=bc.device) if per_symbol else None
        from ..batch import Compressed
        comp = Compressed(words, offsets, 1, n, "ans")
        out, st, pos = bc.ans_decode(comp, table, n_symbols=n, model_index=idx,
                                     index_mode=N.INDEX_PER_SYMBOL if per_symbol else N.INDEX_NONE,
                                     states_in=to_dev_u64([self._state]), raw=True, want_states=True, want_pos=True)
        result = out.cpu().numpy()
        self._state = from_dev_u64(st)[0]
        self._bulk = self._bulk[: int(pos.cpu().numpy()[0])].copy()
        bc.check()
        return result

    def decode(self, model, *params):
        if len(params) == 0:
            return int(self._decode(1, model._concrete_table(), False)[0])
        if len(params) == 1 and is_scalar(params[0]):
            return self._decode(int(params[0]), model._concrete_table(), False)
        n = model._family_len(params)
        return self._decode(n, model._family_table(params), True)
