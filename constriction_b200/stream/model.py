"""Entropy models with the reference's Python names and semantics
(reference: src/pybindings/stream/model.rs, model/internals.rs).

A model object is either *concrete* (all parameters given: one CDF table) or a *family* (some
parameters deferred to per-symbol arrays passed to `encode*` / `decode`: one CDF row per symbol,
tabulated on the device in one launch, then addressed by a per-symbol model index)."""
from __future__ import annotations

import os

import numpy as np

from .. import batch as B


class Model:
    _table = None      # ModelTable of a concrete model (built lazily)
    _nparams = 0

    def _concrete_table(self) -> B.ModelTable:
        raise ValueError("No model parameters specified.")

    def _family_table(self, params) -> B.ModelTable:
        raise ValueError("Model parameters were specified but the model is already fully parameterized.")

    def _family_len(self, params) -> int:
        raise ValueError("Model parameters were specified but the model is already fully parameterized.")


def _float_param(p) -> np.ndarray:
    p = np.asarray(p)
    if p.ndim != 1 or p.dtype not in (np.float32, np.float64):
        raise TypeError("model parameters must be rank-1 numpy arrays with dtype float32 or float64")
    return p.astype(np.float64)  # internals.rs:169-174: f32 parameters are widened to f64


class QuantizedGaussian(Model):
    """pybindings/stream/model.rs:645-708; quantize.rs:284-308,525-568."""

    def __init__(self, min_symbol_inclusive, max_symbol_inclusive, mean=None, std=None):
        self._lo, self._hi = int(min_symbol_inclusive), int(max_symbol_inclusive)
        self._mean, self._std = mean, std
        self._nparams = (mean is None) + (std is None)
        if self._nparams == 0 and not float(std) > 0.0:
            raise ValueError("Invalid model parameter: `std` must be positive.")

    def _concrete_table(self):
        if self._nparams:
            raise ValueError("No model parameters specified.")
        if self._table is None:
            self._table = B.ModelTable.quantized_gaussian(self._lo, self._hi, [float(self._mean)], [float(self._std)])
        return self._table

    def _split(self, params):
        if self._nparams == 0:
            raise ValueError("Model parameters were specified but the model is already fully parameterized.")
        if len(params) != self._nparams:
            raise ValueError(f"Wrong number of model parameters: expected {self._nparams}, got {len(params)}.")
        cols = [_float_param(p) for p in params]
        n = cols[0].size
        if any(c.size != n for c in cols):
            raise ValueError("Model parameters have unequal shape")
        it = iter(cols)
        means = next(it) if self._mean is None else np.full(n, float(self._mean))
        stds = next(it) if self._std is None else np.full(n, float(self._std))
        return means, stds

    def _family_len(self, params):
        return self._split(params)[0].size

    def _family_table(self, params):
        means, stds = self._split(params)
        if not np.all(stds > 0.0):
            raise ValueError("Invalid model parameter: `std` must be positive.")
        if os.environ.get("CTR_GAUSS_TABLES") == "1":  # one tabulated CDF row per symbol (tests compare both paths)
            return B.ModelTable.quantized_gaussian(self._lo, self._hi, means, stds)
        # default: parameters go to the device as they are and the kernels evaluate them (no table per symbol)
        return B.GaussianParams(self._lo, self._hi, means, stds)


class Categorical(Model):
    """pybindings/stream/model.rs:455-560 (fast / lazy quantisation: categorical.rs:16-54,
    lazy_contiguous.rs:131-167,228-330; both give the same table)."""

    def __init__(self, probabilities=None, lazy=None, perfect=None):
        if lazy is None and perfect is None:
            lazy, perfect = False, True
        elif lazy and perfect:
            raise ValueError("Both arguments `lazy` and `perfect` cannot be set to `True` at the same time.")
        else:
            lazy, perfect = bool(lazy), bool(perfect)
        if perfect:
            raise NotImplementedError(
                "Categorical(perfect=True) is outside the accelerated path (SURVEY.md 8f rank 4); use perfect=False")
        self._probs = None
        if probabilities is not None:
            p = np.asarray(probabilities)
            if p.ndim != 1 or p.dtype not in (np.float32, np.float64):
                raise TypeError("probabilities must be a rank-1 numpy array with dtype float32 or float64")
            self._probs = np.ascontiguousarray(p)
            self._table = self._make(self._probs)

    @staticmethod
    def _make(p):
        try:
            return B.ModelTable.categorical(p)
        except ValueError:
            raise ValueError("Probability distribution not normalizable (the array of probabilities\n"
                             "might be empty, contain negative values or NaNs, or sum to infinity).") from None

    def _concrete_table(self):
        if self._probs is None:
            raise ValueError("No model parameters specified.")
        return self._table

    def _matrix(self, params):
        if self._probs is not None:
            raise ValueError("Model parameters were specified but the model is already fully parameterized.")
        if len(params) != 1:
            raise ValueError(f"Wrong number of model parameters: expected 1, got {len(params)}.")
        p = np.asarray(params[0])
        if p.ndim != 2 or p.dtype not in (np.float32, np.float64):
            raise TypeError("probabilities must be a rank-2 numpy array with dtype float32 or float64")
        return np.ascontiguousarray(p)

    def _family_len(self, params):
        return self._matrix(params).shape[0]

    def _family_table(self, params):
        return self._make(self._matrix(params))


class Uniform(Model):
    """pybindings/stream/model.rs:570-600; uniform.rs:44-146: every bin 2^24 // size, the last one takes the
    remainder.  `size` fixed, or one int32 `size` per symbol: the rows are then padded to the largest alphabet with
    zero-probability symbols, which can neither be encoded (KeyError, like a symbol >= size in the reference) nor
    come out of the decoder."""

    def __init__(self, size=None):
        self._size = None if size is None else int(size)
        self._nparams = 1 if size is None else 0

    def _concrete_table(self):
        if self._size is None:
            raise ValueError("No model parameters specified.")
        if self._table is None:
            self._table = B.ModelTable.uniform(self._size)
        return self._table

    def _sizes(self, params):
        if self._size is not None:
            raise ValueError("Model parameters were specified but the model is already fully parameterized.")
        if len(params) != 1:
            raise ValueError(f"Wrong number of model parameters: expected 1, got {len(params)}.")
        sizes = np.asarray(params[0])
        if sizes.ndim != 1 or sizes.dtype != np.int32:
            raise TypeError("size must be a rank-1 numpy array with dtype int32")
        return sizes

    def _family_len(self, params):
        return self._sizes(params).size

    def _family_table(self, params):
        sizes = self._sizes(params).astype(np.int64)
        if sizes.size == 0:
            return B.ModelTable.uniform(2)
        if sizes.min() < 2 or sizes.max() > (1 << 24):
            raise ValueError("Invalid model parameter: `size` must be at least 2 and at most 2^24.")
        widest = int(sizes.max())
        if widest * sizes.size > (1 << 28):
            raise NotImplementedError("Uniform with per-symbol sizes: table too large (alphabet x symbols > 2^28)")
        per_bin = (1 << 24) // sizes
        i = np.arange(widest + 1, dtype=np.int64)[None, :]
        rows = np.where(i < sizes[:, None], i * per_bin[:, None], 1 << 24).astype(np.uint32)
        return B.ModelTable.from_cdf(rows)


class Bernoulli(Model):
    """pybindings/stream/model.rs:985-1060 with perfect=False: a two-symbol categorical model over [1 - p, p]
    quantised by `fast_quantized_cdf` in f64 (categorical.rs:16-54); `p` fixed, or one `p` per symbol."""

    def __init__(self, p=None, perfect=None):
        if perfect is None or perfect:
            raise NotImplementedError(
                "Bernoulli(perfect=True) is outside the accelerated path (SURVEY.md 8f rank 4); use perfect=False")
        self._p = None if p is None else float(p)
        self._nparams = 1 if p is None else 0
        if self._p is not None:
            self._table = self._make(np.array([[1.0 - self._p, self._p]], dtype=np.float64))

    @staticmethod
    def _make(pmf):
        try:
            return B.ModelTable.categorical(pmf)
        except ValueError:
            raise ValueError("`p` must be >= 0.0 and <= 1.0.") from None

    def _concrete_table(self):
        if self._p is None:
            raise ValueError("No model parameters specified.")
        return self._table

    def _params(self, params):
        if self._p is not None:
            raise ValueError("Model parameters were specified but the model is already fully parameterized.")
        if len(params) != 1:
            raise ValueError(f"Wrong number of model parameters: expected 1, got {len(params)}.")
        return _float_param(params[0])

    def _family_len(self, params):
        return self._params(params).size

    def _family_table(self, params):
        p = self._params(params)
        return self._make(np.ascontiguousarray(np.stack([1.0 - p, p], axis=1)))


def _unsupported(name):
    def ctor(*_a, **_k):
        raise NotImplementedError(f"{name} is outside the accelerated path (SURVEY.md 8f); "
                                  "QuantizedGaussian, Categorical(perfect=False), Bernoulli(perfect=False) and Uniform are provided")
    ctor.__name__ = name
    return ctor


QuantizedLaplace = _unsupported("QuantizedLaplace")
QuantizedCauchy = _unsupported("QuantizedCauchy")
Binomial = _unsupported("Binomial")
CustomModel = _unsupported("CustomModel")
ScipyModel = _unsupported("ScipyModel")
